#!/bin/bash
# One GPU-box session: parity tests, the bench lines of the latency-bound workloads with and without the device tail,
# a GKR timeline, and an ncu launch list.  Everything lands in gpurun_out/.
#   gpurun --timeout 1100 -- bash tools/gpu_session_a.sh
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest" ; date +%s
timeout 540 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" ; tail -5 $OUT/pytest_gpu.log
echo "== smoke" ; date +%s
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1 ; echo "smoke exit $?" ; tail -2 $OUT/smoke.log
echo "== bench" ; date +%s
B="timeout 600 python bench.py"
$B --workload plain24 --steps 20 --warmup 5 > $OUT/plain24.json 2> $OUT/plain24.err ; echo "plain24 $?"
ZKB200_TAIL_LOG=0 $B --workload plain24 --steps 20 --warmup 5 --no-e2e --no-cpu --no-probe > $OUT/plain24_tail0.json 2>> $OUT/plain24.err
for tl in 12 14; do ZKB200_TAIL_LOG=$tl $B --workload plain24 --steps 20 --warmup 5 --no-e2e --no-cpu --no-probe > $OUT/plain24_tail$tl.json 2>> $OUT/plain24.err; done
$B --workload gkr_wide --steps 3 --warmup 2 > $OUT/gkr_wide.json 2> $OUT/gkr_wide.err ; echo "gkr_wide $?"
ZKB200_TAIL_LOG=0 $B --workload gkr_wide --steps 3 --warmup 2 --no-e2e --no-cpu > $OUT/gkr_wide_tail0.json 2>> $OUT/gkr_wide.err
ZKB200_TRACE=1 $B --workload gkr_wide --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2> $OUT/gkr_wide_trace.txt
$B --workload gkr --steps 3 --warmup 2 > $OUT/gkr12.json 2> $OUT/gkr12.err ; echo "gkr $?"
$B --log2 26 --steps 5 --warmup 3 --no-e2e --no-cpu --no-probe > $OUT/product26.json 2> $OUT/product26.err
ZKB200_TAIL_LOG=0 $B --log2 26 --steps 5 --warmup 3 --no-e2e --no-cpu --no-probe > $OUT/product26_tail0.json 2>> $OUT/product26.err
echo "== default bench" ; date +%s
if [ "${ZK_SESSION_FULL:-0}" = "1" ]; then $B --steps 3 --warmup 3 --e2e-steps 1 > $OUT/product30.json 2> $OUT/product30.err ; echo "product30 $?"; fi
echo "== ncu" ; date +%s
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_plain24.csv \
    python bench.py --workload plain24 --steps 2 --warmup 1 --no-e2e --no-cpu --no-probe > $OUT/ncu_plain24.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sumcheck_tail_kernel -c 1 -o $OUT/tail_kernel_full \
    python bench.py --workload plain24 --steps 1 --warmup 1 --no-e2e --no-cpu --no-probe > $OUT/ncu_tail.log 2>&1
if [ "${ZK_SESSION_FULL:-0}" = "1" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fold_evals_kernel -s 2 -c 1 -o $OUT/fold_evals_2p27_full \
    python bench.py --log2 27 --steps 1 --warmup 1 --no-e2e --no-cpu --no-probe > $OUT/ncu_fold.log 2>&1
fi
date +%s
for f in plain24 plain24_tail0 plain24_tail12 plain24_tail14 gkr_wide gkr_wide_tail0 gkr12 product26 product26_tail0 product30; do
  python - "$OUT/$f.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    e=d.get("e2e") or {}
    print(sys.argv[1].split('/')[-1], "value=%.4g %s ms=%.3f frac=%.3f kernel_ms=%.3f e2e=%s launches=%s clocks=%s" % (d["value"], d["unit"], d["ms_per_step"], r.get("frac",0), r.get("kernel_ms_per_step",0), e.get("value"), d.get("gpu_launches"), (d.get("clocks") or {}).get("sm_mhz")))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
cat $OUT/gkr_wide_trace.txt | tail -3
