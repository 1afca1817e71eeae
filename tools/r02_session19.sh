#!/bin/bash
# Round-2 session 19 (1 GPU): batched small levels of the KZG opening; succinct GKR bench line
set -u
OUT=gpurun_out/r02_s19
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kzg.py tests/test_gpu_succinct_gkr.py -x -q -m gpu > $OUT/pytest_kzg.log 2>&1 ; echo "pytest kzg rc=$?"
tail -5 $OUT/pytest_kzg.log
for b in 1 0; do
  ZKB200_KZG_BATCH=$b timeout 600 python tools/kzg_timing.py 12 16 20 22 > $OUT/kzg_timing_batch$b.jsonl 2> $OUT/kzg_timing_batch$b.err ; echo "timing batch=$b rc=$?"
  cat $OUT/kzg_timing_batch$b.jsonl; tail -2 $OUT/kzg_timing_batch$b.err
done
timeout 900 python bench.py --workload succinct --steps 3 --warmup 1 > $OUT/bench_succinct.json 2> $OUT/bench_succinct.err ; echo "bench succinct rc=$?"
cut -c1-1500 $OUT/bench_succinct.json; tail -3 $OUT/bench_succinct.err
