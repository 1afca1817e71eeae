#!/bin/bash
# Final 1-GPU session of the round: parity tests, the bench lines recorded under profiles/, ncu launch lists of the
# default command and of the GKR round kernels.   gpurun --timeout 900 -- bash tools/gpu_session_final.sh
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" ; tail -2 $OUT/pytest_gpu.log
B="timeout 600 python bench.py"
$B > $OUT/product30.json 2> $OUT/product30.err ; echo "product30 (driver's default command) $?"
$B --workload plain24 --steps 20 --warmup 5 > $OUT/plain24.json 2> $OUT/plain24.err ; echo "plain24 $?"
$B --workload gkr_wide --steps 3 --warmup 2 > $OUT/gkr_wide.json 2> $OUT/gkr_wide.err ; echo "gkr_wide $?"
$B --workload gkr --steps 3 --warmup 2 > $OUT/gkr12.json 2> $OUT/gkr12.err ; echo "gkr $?"
$B --workload gkr22 --log2 22 --steps 10 --warmup 3 --no-e2e --no-cpu --no-probe > $OUT/gkr22tables.json 2> $OUT/gkr22.err
$B --workload mle --log2 28 --steps 5 --warmup 3 > $OUT/mle28.json 2> $OUT/mle28.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_product30.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-probe > $OUT/ncu_product30.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fold_evals|round_evals|sumcheck_tail" -c 90 --csv --log-file $OUT/launches_gkr_wide_rounds.csv \
    python bench.py --workload gkr_wide --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/ncu_gkr_wide.log 2>&1
for f in product30 plain24 gkr_wide gkr12 gkr22tables mle28; do
  python - "$OUT/$f.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    e=d.get("e2e") or {}
    print(sys.argv[1].split('/')[-1], "value=%.4g %s ms=%.3f frac=%.3f kernel_ms=%s e2e=%s launches=%s clocks=%s" % (d["value"], d["unit"], d["ms_per_step"], r.get("frac",0), r.get("kernel_ms_per_step"), e.get("value"), d.get("gpu_launches"), (d.get("clocks") or {}).get("sm_mhz")))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
