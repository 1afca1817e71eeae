#!/bin/bash
# Round-2 session 6 (1 GPU): one-level Karatsuba in mul_acc (48 instead of 64 IMAD.WIDE per unreduced product) vs the schoolbook rows
set -u
OUT=gpurun_out/r02_s6
mkdir -p $OUT
PKG=zk_cryptography_research_implementations_b200
B="timeout 400 python bench.py --no-cpu --no-extras --no-e2e"
for v in base kara; do
  if [ $v = kara ]; then export ZKB200_LIB=$PWD/$PKG/libzkb200_kara.so; else unset ZKB200_LIB; fi
  $B --steps 5 --warmup 2 > $OUT/product30_$v.json 2> $OUT/product30_$v.err ; echo "product30 $v rc=$?"
  $B --workload mle --log2 28 --steps 5 --warmup 3 > $OUT/mle28_$v.json 2> $OUT/mle28_$v.err ; echo "mle $v rc=$?"
  $B --workload gkr_wide --steps 3 --warmup 2 > $OUT/gkr_wide_$v.json 2> $OUT/gkr_wide_$v.err ; echo "gkr $v rc=$?"
  timeout 300 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_alu.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"round_evals|fold_evals" -c 6 --csv --log-file $OUT/ncu_rounds_$v.csv \
      python bench.py --log2 28 --steps 1 --warmup 0 --no-e2e --no-cpu --no-probe --no-extras > $OUT/ncu_$v.log 2>&1 ; echo "ncu $v rc=$?"
done
unset ZKB200_LIB
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -m gpu -q --timeout=600 -p no:cacheprovider -x > $OUT/pytest_base.log 2>&1; echo "pytest base $?"; tail -2 $OUT/pytest_base.log
ZKB200_LIB=$PWD/$PKG/libzkb200_kara.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -m gpu -q --timeout=600 -p no:cacheprovider -x > $OUT/pytest_kara.log 2>&1; echo "pytest kara $?"; tail -2 $OUT/pytest_kara.log
python - $OUT <<'PY'
import json,sys,glob,os
for f in sorted(glob.glob(sys.argv[1]+"/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        ir=d.get("integer_roofline") or {}
        print("%-24s value=%.6g %s ms=%.4f frac=%.3f verified=%s mul_acc=%s mont=%s fold=%s" % (os.path.basename(f), d["value"], d["unit"], d["ms_per_step"], r.get("frac",0), d.get("verified"), ir.get("mul_acc_unreduced_Gops"), ir.get("mont_mul_Gops"), ir.get("fold_by_scalar_Gops")))
    except Exception as ex:
        print(f, "unreadable:", ex, open(f.replace('.json','.err')).read()[-300:])
PY
for v in base kara; do echo "== ncu $v"; grep -v "^==" $OUT/ncu_rounds_$v.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]
for r in rows[hi+1:]:
    if len(r)>len(h)-1: print(r[h.index('ID')], r[h.index('Kernel Name')][:40], r[h.index('Metric Name')], r[h.index('Metric Value')])
" | head -40; done
