#!/bin/bash
# A/B of the column-form arithmetic (default library) against the chained even/odd-row arithmetic (libzkb200_chained.so)
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" ; tail -2 $OUT/pytest_gpu.log
B="timeout 300 python bench.py"
CH=zk_cryptography_research_implementations_b200/libzkb200_chained.so
$B --log2 28 --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/ab_product28_cols.json 2> $OUT/ab.err
ZKB200_LIB=$CH $B --log2 28 --steps 3 --warmup 3 --no-e2e --no-cpu --no-probe > $OUT/ab_product28_chained.json 2>> $OUT/ab.err
$B --workload plain24 --steps 20 --warmup 5 --no-e2e --no-cpu --no-probe > $OUT/ab_plain24_cols.json 2>> $OUT/ab.err
ZKB200_LIB=$CH $B --workload plain24 --steps 20 --warmup 5 --no-e2e --no-cpu --no-probe > $OUT/ab_plain24_chained.json 2>> $OUT/ab.err
$B --workload mle --log2 28 --steps 5 --warmup 3 --no-cpu > $OUT/ab_mle28_cols.json 2>> $OUT/ab.err
ZKB200_LIB=$CH $B --workload mle --log2 28 --steps 5 --warmup 3 --no-cpu > $OUT/ab_mle28_chained.json 2>> $OUT/ab.err
$B --workload gkr_wide --steps 3 --warmup 2 --no-e2e --no-cpu > $OUT/ab_gkr_wide_cols.json 2>> $OUT/ab.err
for f in ab_product28_cols ab_product28_chained ab_plain24_cols ab_plain24_chained ab_mle28_cols ab_mle28_chained ab_gkr_wide_cols; do
  python - "$OUT/$f.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    i=d.get("integer_roofline") or {}
    pe=d.get("partial_evaluate") or {}
    print(sys.argv[1].split('/')[-1], "value=%.5g %s ms=%.3f frac=%.3f kernel_ms=%s digest=%s %s %s" % (d["value"], d["unit"], d["ms_per_step"], r.get("frac",0), r.get("kernel_ms_per_step"), d.get("proof_digest", d.get("result_digest")), {k:round(v,1) for k,v in i.items() if k.endswith("Gops")}, ("pe_ms=%.3f"%pe["ms"]) if pe else ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
tail -3 $OUT/ab.err
