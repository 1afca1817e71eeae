#!/bin/bash
# Round-2 session 17 (1 GPU): full GPU suite on the tree with the KZG path; KZG timing (sorted / unsorted segments); bench kzg line
set -u
OUT=gpurun_out/r02_s17
mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1 ; echo "pytest gpu rc=$?"
tail -5 $OUT/pytest_gpu.log
for srt in 1 0; do
  ZKB200_MSM_SORT=$srt timeout 600 python tools/kzg_timing.py 12 16 20 22 > $OUT/kzg_timing_sort$srt.jsonl 2> $OUT/kzg_timing_sort$srt.err ; echo "timing sort=$srt rc=$?"
  cat $OUT/kzg_timing_sort$srt.jsonl; tail -2 $OUT/kzg_timing_sort$srt.err
done
timeout 900 python bench.py --workload kzg --steps 4 --warmup 2 > $OUT/bench_kzg.json 2> $OUT/bench_kzg.err ; echo "bench kzg rc=$?"
cut -c1-1800 $OUT/bench_kzg.json; tail -3 $OUT/bench_kzg.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1 ; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active --clock-control none -k regex:"msm_|quotient" -c 120 --csv --log-file $OUT/kzg_kernels.csv python tools/kzg_timing.py 22 > $OUT/ncu.log 2>&1 ; echo "ncu rc=$?"
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
for (kn,mn),(n,t,m) in agg.items():
    if mn.startswith('gpu__time'): print("%-52s n=%3d sum=%12.1f us max=%10.1f us"%(kn,n,t/1e3,m/1e3))
PY
