#!/bin/bash
# Round-2 session 32 (1 GPU): the succinct bench line with its CPU baseline, small size
set -u
OUT=gpurun_out/r02_s32
mkdir -p $OUT
timeout 100 python bench.py --workload succinct --log2 12 --steps 1 --warmup 1 > $OUT/bench_succinct_small.json 2> $OUT/bench_succinct_small.err ; echo "rc=$?"
python - $OUT/bench_succinct_small.json <<'PY'
import json,sys
d=[json.loads(l) for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1]
print(d["value"], d["verified"], json.dumps(d["cpu_baseline"])[:200])
PY
tail -2 $OUT/bench_succinct_small.err | cut -c1-200
