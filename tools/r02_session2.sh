#!/bin/bash
# Round-2 session 2 (1 GPU): the device-resident round loop.  Parity first, then where the hand-over should sit.
#   gpurun --timeout 1500 -- bash tools/r02_session2.sh
set -u
OUT=gpurun_out/r02_s2
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -x > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" ; tail -5 $OUT/pytest_gpu.log
B="timeout 300 python bench.py --no-cpu --no-probe --no-extras"
for tl in 13 20 22 24 26; do
  ZKB200_TAIL_LOG=$tl $B --workload plain24 --steps 20 --warmup 5 --no-e2e > $OUT/plain24_tl$tl.json 2> $OUT/plain24_tl$tl.err ; echo "plain24 tail_log=$tl rc=$?"
done
for tl in 13 22 24; do
  ZKB200_TAIL_LOG=$tl $B --workload gkr_wide --steps 5 --warmup 2 --no-e2e > $OUT/gkr_wide_tl$tl.json 2> $OUT/gkr_wide_tl$tl.err ; echo "gkr_wide tail_log=$tl rc=$?"
done
for tl in 13 24 26; do
  ZKB200_TAIL_LOG=$tl $B --steps 5 --warmup 2 --no-e2e > $OUT/product30_tl$tl.json 2> $OUT/product30_tl$tl.err ; echo "product30 tail_log=$tl rc=$?"
done
ZKB200_TRACE=1 $B --workload gkr_wide --steps 2 --warmup 1 --no-e2e > /dev/null 2> $OUT/gkr_trace.err ; grep "zk_gkr_prove_wide ms" $OUT/gkr_trace.err | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_plain24.csv \
    python bench.py --workload plain24 --steps 1 --warmup 1 --no-e2e --no-cpu --no-probe > $OUT/ncu_plain24.log 2>&1 ; echo "ncu plain24 $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sumcheck_rounds|fold_evals|round_evals|phase|eq_|eval_layer" -c 400 --csv --log-file $OUT/launches_gkr_wide.csv \
    python bench.py --workload gkr_wide --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/ncu_gkr_wide.log 2>&1 ; echo "ncu gkr $?"
python - $OUT <<'PY'
import json,sys,glob,os
for f in sorted(glob.glob(sys.argv[1]+"/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print("%-26s value=%.6g %s ms=%.4f frac=%.3f launches=%s verified=%s setup=%s" % (os.path.basename(f), d["value"], d["unit"], d["ms_per_step"], r.get("frac",0), d.get("gpu_launches"), d.get("verified"), (d.get("config") or {}).get("circuit_setup_s")))
    except Exception as ex:
        print(f, "unreadable:", ex, open(f.replace('.json','.err')).read()[-400:])
PY
