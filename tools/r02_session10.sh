#!/bin/bash
# Round-2 session 10 (1 GPU): GKR with the gate-wise phase-2 work overlapped with phase 1's latency rounds
set -u
OUT=gpurun_out/r02_s10
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_gkr.py tests/test_gpu_tail.py -m gpu -q --timeout=600 -p no:cacheprovider -x > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" ; tail -3 $OUT/pytest_gpu.log
B="timeout 300 python bench.py --no-cpu --no-probe --no-extras --no-e2e"
for ov in 1 0; do for tl in 20 13; do
  ZKB200_GKR_OVERLAP=$ov ZKB200_TAIL_LOG=$tl $B --workload gkr_wide --steps 8 --warmup 3 > $OUT/gkr_wide_ov${ov}_tl$tl.json 2> $OUT/gkr_wide_ov${ov}_tl$tl.err ; echo "gkr_wide overlap=$ov tail_log=$tl rc=$?"
done; done
$B --workload mle --log2 26 --steps 3 --warmup 2 > $OUT/mle26.json 2> $OUT/mle26.err; echo "mle rc=$?"
python - $OUT <<'PY'
import json,sys,glob,os
for f in sorted(glob.glob(sys.argv[1]+"/*.json")):
    try:
        d=[json.loads(l) for l in open(f).read().splitlines() if l.startswith("{")][-1]
        print("%-28s value=%.6g %s verified=%s e2e=%s" % (os.path.basename(f), d["value"], d["unit"], d.get("verified"), (d.get("e2e") or {}).get("value")))
    except Exception as ex:
        print(f, "unreadable:", ex, open(f.replace('.json','.err')).read()[-400:])
PY
