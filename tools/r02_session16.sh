#!/bin/bash
# Round-2 session 16 (1 GPU): KZG with the three-level bucket sums; window widths
set -u
OUT=gpurun_out/r02_s16
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kzg.py tests/test_gpu_succinct_gkr.py -x -q -m gpu > $OUT/pytest_kzg.log 2>&1 ; echo "pytest kzg rc=$?"
tail -5 $OUT/pytest_kzg.log
for v in ""; do
  lib=""; [ -n "$v" ] && lib="$PWD/zk_cryptography_research_implementations_b200/libzkb200_$v.so"
  ZKB200_LIB=$lib timeout 600 python tools/kzg_timing.py 16 20 22 > $OUT/kzg_timing_$v.jsonl 2> $OUT/kzg_timing_$v.err ; echo "timing [$v] rc=$?"
  cat $OUT/kzg_timing_$v.jsonl; tail -2 $OUT/kzg_timing_$v.err
done
for c in 12 13 14; do
  ZKB200_MSM_WINDOW=$c timeout 600 python tools/kzg_timing.py 20 22 > $OUT/kzg_timing_c$c.jsonl 2>&1 ; echo "timing c=$c rc=$?"; cat $OUT/kzg_timing_c$c.jsonl
done
timeout 600 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active --clock-control none -k regex:"msm_|quotient" -c 200 --csv --log-file $OUT/kzg_kernels.csv python tools/kzg_timing.py 22 > $OUT/ncu.log 2>&1 ; echo "ncu rc=$?"
python - $OUT/kzg_kernels.csv <<'PY'
import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<len(h): continue
    k=(r[h.index('Kernel Name')][:50], r[h.index('Metric Name')][:34])
    v=float(r[h.index('Metric Value')].replace(',',''))
    a=agg.setdefault(k,[0,0.0,0.0]); a[0]+=1; a[1]+=v; a[2]=max(a[2],v)
for (kn,mn),(n,t,m) in agg.items(): print("%-52s %-36s n=%3d sum=%14.1f max=%14.1f"%(kn,mn,n,t,m))
PY
