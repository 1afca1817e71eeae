#!/bin/bash
# Round-2 session 18 (1 GPU): the driver's default bench command on the tree with the KZG extra; ncu --set full of the MSM bucket kernel
set -u
OUT=gpurun_out/r02_s18
mkdir -p $OUT
( time timeout 1500 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err ) 2> $OUT/bench_default.time ; echo "bench default rc=$?"
tail -3 $OUT/bench_default.time
python - $OUT/bench_default.json <<'PY'
import json,sys
d=[json.loads(l) for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1]
print("headline", d["value"], d["unit"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "verified", d.get("verified"), "e2e", d["e2e"]["value"])
for e in d.get("extra_workloads", []):
    print("  extra", e.get("metric"), e.get("value"), e.get("unit"), "verified", e.get("verified"), e.get("error"))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"msm_bucket_kernel" -s 1 -c 1 -o $OUT/msm_bucket_2p22 python tools/kzg_timing.py 22 > $OUT/ncu_full.log 2>&1 ; echo "ncu full rc=$?"
ls -la $OUT
