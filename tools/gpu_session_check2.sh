#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" ; tail -2 $OUT/pytest_gpu.log
B="timeout 600 python bench.py"
$B --workload gkr_wide --steps 3 --warmup 2 > $OUT/gkr_wide.json 2> $OUT/gkr_wide.err ; echo "gkr_wide $?"
ZKB200_TRACE=1 $B --workload gkr_wide --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2> $OUT/gkr_wide_trace.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/gkr_wide.json").read().strip().splitlines()[-1])
print("gkr_wide value=%.3f ms e2e=%s kernel_ms=%.3f launches=%s digest=%s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_step"], d["gpu_launches"], d["proof_digest"]))
PY
tail -1 $OUT/gkr_wide_trace.txt
