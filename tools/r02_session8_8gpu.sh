#!/bin/bash
# Round-2 session 8 (8 GPUs): sharded parity at 4 and 8 ranks, the scaling bench at N = 8 / 4 with both exchange paths,
# multi-GPU GKR (row e2), the 2^32 MLE point.   gpurun --gpus 8 --timeout 1200 -- bash tools/r02_session8_8gpu.sh
set -u
OUT=gpurun_out/r02_s8
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout=800 -p no:cacheprovider -s -k "4 or 8" > $OUT/pytest_sharded.log 2>&1
echo "pytest sharded (4, 8 ranks) exit $?" ; tail -4 $OUT/pytest_sharded.log | cut -c1-600
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n "$@"; }
run 8 --steps 10 --warmup 3 > $OUT/bench_default_n8.json 2> $OUT/bench_default_n8.err ; echo "bench default N=8 rc=$?"
ZKB200_PEER_EXCHANGE=0 run 8 --steps 10 --warmup 3 --no-extras --no-e2e --no-cpu --no-probe > $OUT/product30_n8_mailbox.json 2> $OUT/product30_n8_mailbox.err ; echo "product30 N=8 mailboxes rc=$?"
ZKB200_TAIL_LOG=13 run 8 --steps 10 --warmup 3 --no-extras --no-e2e --no-cpu --no-probe > $OUT/product30_n8_tl13.json 2> $OUT/product30_n8_tl13.err ; echo "product30 N=8 tail_log 13 rc=$?"
ZKB200_PEER_EXCHANGE=0 run 8 --workload gkr_wide --steps 5 --warmup 2 --no-cpu > $OUT/gkr_wide_n8_mailbox.json 2> $OUT/gkr_wide_n8_mailbox.err ; echo "gkr_wide N=8 mailboxes rc=$?"
ZKB200_TRACE=1 run 8 --workload gkr_wide --steps 2 --warmup 1 --no-cpu > /dev/null 2> $OUT/gkr_trace_n8.err ; grep "zk_gkr_prove_wide ms" $OUT/gkr_trace_n8.err | tail -1
run 4 --steps 10 --warmup 3 --no-e2e > $OUT/bench_default_n4.json 2> $OUT/bench_default_n4.err ; echo "bench default N=4 rc=$?"
run 8 --workload plain24 --log2 30 --steps 10 --warmup 3 --no-cpu --no-probe > $OUT/plain30_n8.json 2> $OUT/plain30_n8.err ; echo "plain30 N=8 rc=$?"
python - $OUT <<'PY'
import json,sys,glob,os
def show(name,d):
    r=d.get("roofline") or {}; e=d.get("e2e") or {}
    print("%-30s n=%s value=%.6g %s ms=%.4f frac=%.3f verified=%s e2e=%s pcie=%s exch=%s" % (name, d.get("n_gpus"), d["value"], d["unit"], d["ms_per_step"], r.get("frac",0), d.get("verified"), e.get("value"), e.get("pcie_h2d_GBps_per_gpu"), (d.get("exchange") or "")[:12]))
for f in sorted(glob.glob(sys.argv[1]+"/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        show(os.path.basename(f),d)
        for x in d.get("extra_workloads",[]):
            if "error" in x: print("   EXTRA ERROR",x)
            else: show("   extra:"+x["config"]["workload"][:14],x)
    except Exception as ex:
        print(f, "unreadable:", ex, open(f.replace('.json','.err')).read()[-800:])
PY
