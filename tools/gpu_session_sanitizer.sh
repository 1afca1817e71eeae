#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 200 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitizer_probe.py > $OUT/sanitizer_racecheck.txt 2>&1
echo "racecheck exit $?"; tail -4 $OUT/sanitizer_racecheck.txt
timeout 200 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitizer_probe.py > $OUT/sanitizer_memcheck.txt 2>&1
echo "memcheck exit $?"; tail -4 $OUT/sanitizer_memcheck.txt
timeout 200 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" ; tail -2 $OUT/pytest_gpu.log
