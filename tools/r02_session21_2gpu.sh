#!/bin/bash
# Round-2 session 21 (2 GPUs): sharded KZG / succinct GKR parity and the succinct bench line on 2 ranks
set -u
OUT=gpurun_out/r02_s21
mkdir -p $OUT
timeout 420 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu -k kzg > $OUT/pytest_sharded_kzg_2gpu.log 2>&1 ; echo "pytest sharded kzg rc=$?"
tail -6 $OUT/pytest_sharded_kzg_2gpu.log | cut -c1-600
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload succinct --steps 3 --warmup 1 > $OUT/bench_succinct_n2.json 2> $OUT/bench_succinct_n2.err ; echo "bench succinct n2 rc=$?"
grep '^{' $OUT/bench_succinct_n2.json | cut -c1-400; tail -3 $OUT/bench_succinct_n2.err | cut -c1-300
