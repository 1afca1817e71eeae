#!/bin/bash
# Round-2 session 29 (1 GPU): the golden-vector test of the KZG path
set -u
OUT=gpurun_out/r02_s29
mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_kzg.py -x -q -m gpu -k "golden or reference" > $OUT/pytest_kzg_golden.log 2>&1 ; echo "pytest rc=$?"
tail -3 $OUT/pytest_kzg_golden.log | cut -c1-400
