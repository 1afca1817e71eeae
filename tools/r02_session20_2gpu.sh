#!/bin/bash
# Round-2 session 20 (2 GPUs): sharded provers incl. the sharded KZG / succinct GKR
set -u
OUT=gpurun_out/r02_s20
mkdir -p $OUT
nvidia-smi -L | head -4
timeout 1700 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > $OUT/pytest_sharded_2gpu.log 2>&1 ; echo "pytest sharded rc=$?"
tail -25 $OUT/pytest_sharded_2gpu.log | cut -c1-400
