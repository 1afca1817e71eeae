"""The GPU arm of bench.py at a small size: one JSON line with every key of the contract."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_gpu_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--log2", "20", "--steps", "2", "--warmup", "3", "--cpu-log2", "12",
                          "--e2e-steps", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["vs_baseline"] is None
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] == 2 * (1 << 20) * 32 and d["e2e"]["value"] > 0 and d["e2e"]["value"] != d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["gpu_launches"] > 0
    # the proof of the last timed step is checked after the timed region: reference verifier + evaluate kernels
    assert d["verified"] is True and all(d["verify"].values()), d["verify"]
    assert len(d["proof_digest"]) == 64
    assert d["config"]["sample_log2"] == 12 and "extra_workloads" not in d      # --log2 given: no extras


def test_gpu_arm_other_workloads_verify_their_proofs():
    for args in (["--workload", "plain24", "--log2", "18"], ["--workload", "gkr_wide", "--log2", "12", "--cpu-depth", "3"],
                 ["--workload", "mle", "--log2", "18", "--cpu-log2", "18", "--sweep", "16"]):
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), *args, "--steps", "2", "--warmup", "1"]
        if "--cpu-log2" not in args:
            cmd += ["--cpu-log2", "12"]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
        assert out.returncode == 0, out.stderr[-2000:]
        d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
        assert d["verified"] is True, (args, d.get("verify"))
        assert d["roofline"]["frac"] > 0 and d["clocks"] is not None


def test_gpu_arm_kzg_and_succinct_lines_verify_with_pairings():
    """the multilinear-KZG line (commitment + opening checked by the pairing verifier and against the oracle's commitment of the
    CPU sample) and the succinct-GKR line (verify_succinct) at small sizes"""
    for args, metric in ((["--workload", "kzg", "--log2", "14"], "kzg_commit_Mpoints_per_s"),
                         (["--workload", "succinct", "--log2", "12"], "succinct_gkr_prove_ms")):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args, "--steps", "2", "--warmup", "1"],
                             capture_output=True, text=True, timeout=900, cwd=ROOT)
        assert out.returncode == 0, out.stderr[-2000:]
        d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
        assert d["metric"] == metric and d["verified"] is True and d["value"] > 0 and d["gpu_launches"] > 0
        assert d["e2e"]["value"] > 0 and d["clocks"] is not None and len(d["result_digest"]) == 64
        if metric.startswith("kzg"):
            assert d["cpu_baseline"]["kind"] == "port" and d["roofline"]["bound"] == "imad" and d["roofline"]["peak"] > 0
