#!/bin/bash
# Round-2 session 7 (2 GPUs): the sharded parity test after the early-stop fix; sharded GKR with a larger collapse length
set -u
OUT=gpurun_out/r02_s7
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout=800 -p no:cacheprovider -x -s > $OUT/pytest_sharded.log 2>&1
echo "pytest sharded exit $?" ; tail -4 $OUT/pytest_sharded.log | cut -c1-600
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2"
for cl in 4096 65536; do
  $T --workload gkr_wide --steps 5 --warmup 2 --no-cpu --collapse-len $cl > $OUT/gkr_wide_n2_cl$cl.json 2> $OUT/gkr_wide_n2_cl$cl.err ; echo "gkr_wide N=2 collapse $cl rc=$?"
done
$T --workload plain24 --log2 28 --steps 10 --warmup 3 --no-cpu --no-probe > $OUT/plain28_n2.json 2> $OUT/plain28_n2.err ; echo "plain28 N=2 rc=$?"
python - $OUT <<'PY'
import json,sys,glob,os
for f in sorted(glob.glob(sys.argv[1]+"/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print("%-30s n=%s value=%.6g %s ms=%.4f frac=%.3f verified=%s" % (os.path.basename(f), d.get("n_gpus"), d["value"], d["unit"], d["ms_per_step"], r.get("frac",0), d.get("verified")))
    except Exception as ex:
        print(f, "unreadable:", ex, open(f.replace('.json','.err')).read()[-600:])
PY
