#!/bin/bash
# Short validation session: parity tests + the latency-bound bench lines + the MLE sweep.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" ; tail -2 $OUT/pytest_gpu.log
B="timeout 600 python bench.py"
$B --workload gkr_wide --steps 3 --warmup 2 > $OUT/gkr_wide.json 2> $OUT/gkr_wide.err ; echo "gkr_wide $?"
ZKB200_TRACE=1 $B --workload gkr_wide --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2> $OUT/gkr_wide_trace.txt
$B --workload gkr --steps 5 --warmup 2 > $OUT/gkr12.json 2> $OUT/gkr12.err ; echo "gkr $?"
$B --workload plain24 --steps 20 --warmup 5 > $OUT/plain24.json 2> $OUT/plain24.err ; echo "plain24 $?"
$B --workload gkr22 --log2 22 --steps 10 --warmup 3 --no-e2e --no-cpu --no-probe > $OUT/gkr22tables.json 2> $OUT/gkr22.err
$B --workload mle --log2 28 --sweep 20,22,24,26 --steps 5 --warmup 3 > $OUT/mle28.json 2> $OUT/mle28.err ; echo "mle $?"
for f in gkr_wide gkr12 plain24 gkr22tables mle28; do
  python - "$OUT/$f.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    e=d.get("e2e") or {}
    print(sys.argv[1].split('/')[-1], "value=%.4g %s ms=%.3f frac=%.3f kernel_ms=%s e2e=%s launches=%s" % (d["value"], d["unit"], d["ms_per_step"], r.get("frac",0), r.get("kernel_ms_per_step"), e.get("value"), d.get("gpu_launches")))
    for s in d.get("sweep") or []: print("   sweep 2^%d: evaluate %.3f ms (%.3g el/s, %.2f of HBM), partial_evaluate %.3f ms (%.2f of HBM)" % (s["log2_entries"], s["evaluate_ms"], s["evaluate_elements_per_s"], s["evaluate_frac_hbm"], s["partial_evaluate_ms"], s["partial_evaluate_frac_hbm"]))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
tail -2 $OUT/gkr_wide_trace.txt
