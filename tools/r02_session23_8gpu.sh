#!/bin/bash
# Round-2 session 23 (8 GPUs): the driver's 8-rank command on the final tree (headline + extras incl. sharded succinct GKR)
set -u
OUT=gpurun_out/r02_s23
mkdir -p $OUT
nvidia-smi -L | wc -l
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/bench_default_n8.json 2> $OUT/bench_default_n8.err ) 2> $OUT/bench_default_n8.time ; echo "bench n8 rc=$?"
tail -3 $OUT/bench_default_n8.time
python - $OUT/bench_default_n8.json <<'PY'
import json,sys
ls=[json.loads(l) for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")]
if ls:
    d=ls[-1]
    print("headline", d["value"], d["unit"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "verified", d.get("verified"))
    for e in d.get("extra_workloads", []):
        print("  extra", e.get("metric"), e.get("value"), e.get("unit"), "verified", e.get("verified"), e.get("error"))
PY
tail -5 $OUT/bench_default_n8.err | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu -k "kzg and 8" > $OUT/pytest_sharded_kzg_8gpu.log 2>&1 ; echo "pytest sharded kzg 8 rc=$?"
tail -3 $OUT/pytest_sharded_kzg_8gpu.log | cut -c1-300
