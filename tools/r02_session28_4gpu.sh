#!/bin/bash
# Round-2 session 28 (4 GPUs): sharded GKR 16 x 2^22, when to gather the shards (collapse_len sweep)
set -u
OUT=gpurun_out/r02_s28
mkdir -p $OUT
for cl in 4096 32768 131072 524288; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --workload gkr_wide --collapse-len $cl --steps 6 --warmup 2 --no-cpu --no-e2e > $OUT/gkr_wide_n4_cl$cl.json 2> $OUT/gkr_wide_n4_cl$cl.err ; echo "cl=$cl rc=$?"
  python - $OUT/gkr_wide_n4_cl$cl.json <<'PY'
import json,sys
ls=[json.loads(l) for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")]
if ls: print("   ", ls[-1]["value"], "ms verified", ls[-1].get("verified"))
PY
done
