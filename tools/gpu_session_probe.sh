#!/bin/bash
# ncu --set full captures (fine-grained PC sampling) of the GKR-shaped tail kernel and of a tiny GKR-shaped fold launch
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:sumcheck_tail_kernel -s 2 -c 1 -o $OUT/tail_nlin_full \
    python bench.py --workload gkr_wide --log2 18 --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/ncu_tail_nlin.log 2>&1
echo "tail capture $?"
timeout 300 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:fold_evals_kernel -s 4 -c 1 -o $OUT/fold_nlin_small_full \
    python bench.py --workload gkr_wide --log2 18 --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/ncu_fold_nlin.log 2>&1
echo "fold capture $?"
ls -la $OUT/*.ncu-rep
