#!/bin/bash
# Round-2 session 27 (1 GPU): the bench-contract tests incl. the kzg / succinct lines
set -u
OUT=gpurun_out/r02_s27
mkdir -p $OUT
timeout 800 python -m pytest tests/test_gpu_bench_contract.py -x -q -m gpu > $OUT/pytest_bench_contract.log 2>&1 ; echo "pytest rc=$?"
tail -5 $OUT/pytest_bench_contract.log | cut -c1-600
