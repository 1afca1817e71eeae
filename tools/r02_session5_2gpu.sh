#!/bin/bash
# Round-2 session 5 (2 GPUs): sharded provers with the in-kernel peer exchange, multi-GPU GKR
#   gpurun --gpus 2 --timeout 1500 -- bash tools/r02_session5_2gpu.sh
set -u
OUT=gpurun_out/r02_s5
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout=800 -p no:cacheprovider -x -s > $OUT/pytest_sharded.log 2>&1
echo "pytest sharded exit $?" ; tail -6 $OUT/pytest_sharded.log
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2"
$T --steps 5 --warmup 2 > $OUT/bench_default_n2.json 2> $OUT/bench_default_n2.err ; echo "bench default N=2 rc=$?"
$T --workload gkr_wide --steps 5 --warmup 2 --no-cpu > $OUT/gkr_wide_n2.json 2> $OUT/gkr_wide_n2.err ; echo "gkr_wide N=2 rc=$?"
ZKB200_PEER_EXCHANGE=0 $T --workload gkr_wide --steps 5 --warmup 2 --no-cpu > $OUT/gkr_wide_n2_mailbox.json 2> $OUT/gkr_wide_n2_mailbox.err ; echo "gkr_wide N=2 mailboxes rc=$?"
ZKB200_TRACE=1 $T --workload gkr_wide --steps 2 --warmup 1 --no-cpu > /dev/null 2> $OUT/gkr_trace_n2.err ; grep "zk_gkr_prove_wide ms" $OUT/gkr_trace_n2.err | tail -2
$T --log2 28 --steps 10 --warmup 3 --no-cpu --no-probe --no-e2e > $OUT/product28_n2.json 2> $OUT/product28_n2.err ; echo "product28 N=2 rc=$?"
ZKB200_PEER_EXCHANGE=0 $T --log2 28 --steps 10 --warmup 3 --no-cpu --no-probe --no-e2e > $OUT/product28_n2_mailbox.json 2> $OUT/product28_n2_mailbox.err ; echo "product28 N=2 mailboxes rc=$?"
timeout 300 python bench.py --workload gkr_wide --steps 5 --warmup 2 --no-cpu --no-e2e > $OUT/gkr_wide_n1.json 2> $OUT/gkr_wide_n1.err ; echo "gkr_wide N=1 rc=$?"
python - $OUT <<'PY'
import json,sys,glob,os
def show(name,d):
    r=d.get("roofline") or {}
    print("%-30s n=%s value=%.6g %s ms=%.4f frac=%.3f verified=%s exchange=%s" % (name, d.get("n_gpus"), d["value"], d["unit"], d["ms_per_step"], r.get("frac",0), d.get("verified"), (d.get("exchange") or "")[:40]))
for f in sorted(glob.glob(sys.argv[1]+"/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        show(os.path.basename(f),d)
        for x in d.get("extra_workloads",[]):
            if "error" in x: print("   EXTRA ERROR",x)
            else: show("   extra:"+x["config"]["workload"][:14],x)
    except Exception as ex:
        print(f, "unreadable:", ex, open(f.replace('.json','.err')).read()[-600:])
PY
