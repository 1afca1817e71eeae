#!/bin/bash
# Round-2 session 13 (1 GPU): segmented builders at 4 blocks/SM
set -u
OUT=gpurun_out/r02_s13
mkdir -p $OUT
B="timeout 300 python bench.py --no-cpu --no-probe --no-extras --no-e2e"
for ov in 1 0; do
  ZKB200_GKR_OVERLAP=$ov $B --workload gkr_wide --steps 8 --warmup 3 > $OUT/gkr_wide_ov$ov.json 2> $OUT/gkr_wide_ov$ov.err ; echo "gkr_wide overlap=$ov rc=$?"
done
timeout 300 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"seg_bucket|phase2_pre|eval_layer_kernel|eq_outer" -c 40 --csv --log-file $OUT/builders.csv \
    python bench.py --workload gkr_wide --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/ncu.log 2>&1 ; echo "ncu rc=$?"
ZKB200_GKR_OVERLAP=0 timeout 300 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"seg_bucket" -c 16 --csv --log-file $OUT/builders_ov0.csv \
    python bench.py --workload gkr_wide --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/ncu0.log 2>&1 ; echo "ncu rc=$?"
for f in builders builders_ov0; do python - $OUT/$f.csv <<'PY'
import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<len(h): continue
    k=(r[h.index('Kernel Name')][:66], r[h.index('Metric Name')][:30])
    v=float(r[h.index('Metric Value')].replace(',','')); u=r[h.index('Metric Unit')]
    a=agg.setdefault(k,[0,0.0,u]); a[0]+=1; a[1]+=v
for (kn,mn),(n,t,u) in agg.items(): print("%-68s %-32s n=%3d avg=%12.3f %s"%(kn,mn,n,t/n,u))
PY
done
python - $OUT <<'PY'
import json,sys,glob,os
for f in sorted(glob.glob(sys.argv[1]+"/*.json")):
    d=[json.loads(l) for l in open(f).read().splitlines() if l.startswith("{")][-1]
    print("%-28s value=%.6g %s verified=%s" % (os.path.basename(f), d["value"], d["unit"], d.get("verified")))
PY
