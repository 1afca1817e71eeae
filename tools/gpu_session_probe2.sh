#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"phase1|phase2|eq_halves|eq_outer2|eval_layer|sum_slices" -c 140 --csv --log-file $OUT/launches_gkr_wide_builders.csv \
    python bench.py --workload gkr_wide --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/ncu_gkr_builders.log 2>&1
echo "exit $?"
