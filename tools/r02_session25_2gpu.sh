#!/bin/bash
# Round-2 session 25 (2 GPUs): every sharded parity test on the final tree
set -u
OUT=gpurun_out/r02_s25
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > $OUT/pytest_sharded_2gpu.log 2>&1 ; echo "pytest sharded rc=$?"
tail -5 $OUT/pytest_sharded_2gpu.log | cut -c1-500
