#!/bin/bash
# Round-2 session 26 (4 GPUs): the driver's 4-rank command on the final tree
set -u
OUT=gpurun_out/r02_s26
mkdir -p $OUT
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 5 --warmup 3 > $OUT/bench_default_n4.json 2> $OUT/bench_default_n4.err ) 2> $OUT/bench_default_n4.time ; echo "bench n4 rc=$?"
tail -3 $OUT/bench_default_n4.time | head -1
python - $OUT/bench_default_n4.json <<'PY'
import json,sys
ls=[json.loads(l) for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")]
if ls:
    d=ls[-1]
    print("headline", d["value"], d["unit"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "verified", d.get("verified"))
    for e in d.get("extra_workloads", []):
        print("  extra", e.get("metric"), e.get("value"), e.get("unit"), "verified", e.get("verified"), e.get("error"))
PY
tail -3 $OUT/bench_default_n4.err | cut -c1-300
