#!/bin/bash
# Round-2 session 24 (1 GPU): the driver's round-end sequence on the final tree: GPU tests, smoke, reference arm, default bench
set -u
OUT=gpurun_out/r02_s24
mkdir -p $OUT
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1 ; echo "pytest gpu rc=$?"
tail -4 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1 ; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
( time timeout 900 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err ) 2> $OUT/bench_reference.time ; echo "bench reference rc=$?"; tail -3 $OUT/bench_reference.time | head -1
cut -c1-300 $OUT/bench_reference.json
( time timeout 1500 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err ) 2> $OUT/bench_default.time ; echo "bench default rc=$?"
tail -3 $OUT/bench_default.time | head -1
python - $OUT/bench_default.json <<'PY'
import json,sys
d=[json.loads(l) for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1]
print("headline", d["value"], d["unit"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "verified", d.get("verified"), "e2e", d["e2e"]["value"], "launches", d.get("gpu_launches"))
for e in d.get("extra_workloads", []):
    print("  extra", e.get("metric"), e.get("value"), e.get("unit"), "verified", e.get("verified"), "e2e", (e.get("e2e") or {}).get("value"), e.get("error"))
PY
