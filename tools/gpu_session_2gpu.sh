#!/bin/bash
# Two-GPU session: the sharded provers (NCCL + shared mailboxes, collapse, device tail on every rank) against the oracle,
# then the product sumcheck bench at 2 ranks.   gpurun --gpus 2 --timeout 600 -- bash tools/gpu_session_2gpu.sh
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout=300 -p no:cacheprovider > $OUT/pytest_2gpu.log 2>&1
echo "pytest exit $?"; tail -3 $OUT/pytest_2gpu.log
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2"
$T --log2 28 --steps 3 --warmup 3 --no-e2e > $OUT/product28_n2.json 2> $OUT/product28_n2.err; echo "product28 n2 $?"
$T --log2 28 --steps 3 --warmup 3 --no-e2e --nccl-exchange > $OUT/product28_n2_nccl.json 2>> $OUT/product28_n2.err; echo "product28 n2 nccl $?"
$T --workload plain32 --log2 28 --steps 3 --warmup 3 --no-e2e > $OUT/plain28_n2.json 2> $OUT/plain28_n2.err; echo "plain28 n2 $?"
$T --workload mle --log2 28 --steps 3 --warmup 2 > $OUT/mle28_n2.json 2> $OUT/mle28_n2.err; echo "mle28 n2 $?"
timeout 300 python bench.py --log2 28 --steps 3 --warmup 3 --no-e2e --no-cpu --no-probe > $OUT/product28_n1.json 2> $OUT/product28_n1.err; echo "product28 n1 $?"
for f in product28_n1 product28_n2 product28_n2_nccl plain28_n2 mle28_n2; do
  python - "$OUT/$f.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print(sys.argv[1].split('/')[-1], "value=%.4g %s ms=%.3f frac=%.3f kernel_ms=%s launches=%s digest=%s" % (d["value"], d["unit"], d["ms_per_step"], r.get("frac",0), r.get("kernel_ms_per_step"), d.get("gpu_launches"), d.get("proof_digest", d.get("result_digest"))))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
tail -3 $OUT/product28_n2.err
