#!/bin/bash
# Round-2 session 11 (1 GPU): kernel durations of the GKR builders (split phase-2 build vs the one-kernel build)
set -u
OUT=gpurun_out/r02_s11
mkdir -p $OUT
for ov in 1 0; do
ZKB200_GKR_OVERLAP=$ov timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"phase|eq_outer|eval_layer_kernel" -c 40 --csv --log-file $OUT/builders_ov$ov.csv \
    python bench.py --workload gkr_wide --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/ncu_ov$ov.log 2>&1 ; echo "ncu overlap=$ov rc=$?"
python - $OUT/builders_ov$ov.csv <<'PY'
import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<len(h): continue
    k=(r[h.index('Kernel Name')].split('(')[0][-28:], r[h.index('Metric Name')])
    v=float(r[h.index('Metric Value')].replace(',','')); u=r[h.index('Metric Unit')]
    a=agg.setdefault(k,[0,0.0,u]); a[0]+=1; a[1]+=v
for (kn,mn),(n,t,u) in agg.items(): print("%-30s %-55s n=%3d avg=%12.3f %s"%(kn,mn,n,t/n,u))
PY
done
