#!/bin/bash
# Round-2 session 22 (1 GPU): bucket kernel with / without the software prefetch of the next point
set -u
OUT=gpurun_out/r02_s22
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kzg.py -x -q -m gpu > $OUT/pytest_kzg.log 2>&1 ; echo "pytest kzg rc=$?"; tail -2 $OUT/pytest_kzg.log
for v in "" nopf; do
  lib=""; [ -n "$v" ] && lib="$PWD/zk_cryptography_research_implementations_b200/libzkb200_$v.so"
  for rep in 1 2; do
    ZKB200_LIB=$lib timeout 300 python tools/kzg_timing.py 20 22 > $OUT/kzg_timing_${v}_$rep.jsonl 2> $OUT/kzg_timing_${v}_$rep.err ; echo "timing [$v] rc=$?"
    cut -c1-230 $OUT/kzg_timing_${v}_$rep.jsonl
  done
done
