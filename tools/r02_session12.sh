#!/bin/bash
# Round-2 session 12 (1 GPU): GKR table builders as warp-segmented gate-parallel bucket sums; overlap on/off
set -u
OUT=gpurun_out/r02_s12
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_gkr.py tests/test_gpu_tail.py -m gpu -q --timeout=600 -p no:cacheprovider -x > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" ; tail -3 $OUT/pytest_gpu.log
B="timeout 300 python bench.py --no-cpu --no-probe --no-extras --no-e2e"
for seg in 1 0; do for ov in 1 0; do
  ZKB200_GKR_SEG=$seg ZKB200_GKR_OVERLAP=$ov $B --workload gkr_wide --steps 8 --warmup 3 > $OUT/gkr_wide_seg${seg}_ov$ov.json 2> $OUT/gkr_wide_seg${seg}_ov$ov.err ; echo "gkr_wide seg=$seg overlap=$ov rc=$?"
done; done
ZKB200_TRACE=1 $B --workload gkr_wide --steps 2 --warmup 1 > /dev/null 2> $OUT/gkr_trace.err ; grep "zk_gkr_prove_wide ms" $OUT/gkr_trace.err | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"seg_bucket|phase2_pre" -c 24 --csv --log-file $OUT/builders_seg.csv \
    python bench.py --workload gkr_wide --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/ncu_seg.log 2>&1 ; echo "ncu rc=$?"
python - $OUT/builders_seg.csv <<'PY'
import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<len(h): continue
    k=(r[h.index('Kernel Name')][:70], r[h.index('Metric Name')])
    v=float(r[h.index('Metric Value')].replace(',','')); u=r[h.index('Metric Unit')]
    a=agg.setdefault(k,[0,0.0,u]); a[0]+=1; a[1]+=v
for (kn,mn),(n,t,u) in agg.items(): print("%-72s %-50s n=%3d avg=%12.3f %s"%(kn,mn,n,t/n,u))
PY
python - $OUT <<'PY'
import json,sys,glob,os
for f in sorted(glob.glob(sys.argv[1]+"/*.json")):
    try:
        d=[json.loads(l) for l in open(f).read().splitlines() if l.startswith("{")][-1]
        print("%-28s value=%.6g %s verified=%s" % (os.path.basename(f), d["value"], d["unit"], d.get("verified")))
    except Exception as ex:
        print(f, "unreadable:", ex, open(f.replace('.json','.err')).read()[-400:])
PY
