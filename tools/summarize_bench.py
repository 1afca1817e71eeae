#!/usr/bin/env python3
"""Markdown table of the bench JSON lines under a profiles/ directory:  python tools/summarize_bench.py profiles/r01b"""
import json
import os
import sys


def fmt(v, spec="%.4g"):
    return "-" if v is None else spec % v


def main(d):
    rows = []
    for name in sorted(os.listdir(d)):
        if not name.endswith(".json"):
            continue
        try:
            line = json.loads(open(os.path.join(d, name)).read().strip().splitlines()[-1])
        except Exception:
            continue
        r = line.get("roofline") or {}
        e = line.get("e2e") or {}
        c = line.get("cpu_baseline") or {}
        ck = line.get("clocks") or {}
        rows.append("| `%s` | %s | %s %s | %s | %s | %s | %s %s | %s | %s | %s |" % (
            name[:-5], line.get("n_gpus"), fmt(line.get("value")), line.get("unit"), fmt(line.get("ms_per_step"), "%.3f"),
            fmt(r.get("frac"), "%.3f"), fmt(r.get("kernel_ms_per_step"), "%.3f"), fmt(e.get("value")), e.get("unit", ""),
            fmt(c.get("value")), line.get("gpu_launches", "-"), fmt(ck.get("sm_mhz"), "%.0f")))
    print("| run | GPUs | value | ms / step | roofline frac | round-kernel ms / step | e2e | CPU baseline | launches | SM MHz |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    print("\n".join(rows))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "profiles/r01b")
