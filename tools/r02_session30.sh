#!/bin/bash
# Round-2 session 30 (1 GPU): the KZG / succinct tests on the library with the plan / digits / host combine moved into shared headers
set -u
OUT=gpurun_out/r02_s30
mkdir -p $OUT
timeout 150 python -m pytest tests/test_gpu_kzg.py tests/test_gpu_succinct_gkr.py -x -q -m gpu > $OUT/pytest_kzg.log 2>&1 ; echo "pytest rc=$?"
tail -3 $OUT/pytest_kzg.log | cut -c1-400
