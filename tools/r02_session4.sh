#!/bin/bash
# Round-2 session 4 (1 GPU): transcript step after the lazy interpolation / unrolled Keccak; GKR builders with eq formed per gate
set -u
OUT=gpurun_out/r02_s4
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tail.py tests/test_gpu_gkr.py tests/test_gpu_parity_large.py -m gpu -q --timeout=600 -p no:cacheprovider -x > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" ; tail -3 $OUT/pytest_gpu.log
B="timeout 300 python bench.py --no-cpu --no-probe --no-extras --no-e2e"
PKG=zk_cryptography_research_implementations_b200
ZKB200_LIB=$PWD/$PKG/libzkb200_timing.so ZKB200_DEV_TIMING=1 ZKB200_TAIL_LOG=24 $B --workload plain24 --steps 2 --warmup 2 > $OUT/plain24_timing.json 2> $OUT/plain24_timing.err
ZKB200_LIB=$PWD/$PKG/libzkb200_timing.so ZKB200_DEV_TIMING=1 ZKB200_TAIL_LOG=24 $B --workload gkr_wide --steps 2 --warmup 1 > $OUT/gkr_wide_timing.json 2> $OUT/gkr_wide_timing.err
for f in plain24_timing gkr_wide_timing; do echo "== $f"; grep devrounds $OUT/$f.err | tail -1 | cut -c1-2600; done
for tl in 13 18 20 22; do
  ZKB200_TAIL_LOG=$tl $B --workload plain24 --steps 20 --warmup 5 > $OUT/plain24_tl$tl.json 2> $OUT/plain24_tl$tl.err
  ZKB200_TAIL_LOG=$tl $B --workload gkr_wide --steps 5 --warmup 2 > $OUT/gkr_wide_tl$tl.json 2> $OUT/gkr_wide_tl$tl.err
done
ZKB200_GKR_TABLES=1 $B --workload gkr_wide --steps 5 --warmup 2 > $OUT/gkr_wide_tables.json 2> $OUT/gkr_wide_tables.err
ZKB200_TRACE=1 $B --workload gkr_wide --steps 2 --warmup 1 > /dev/null 2> $OUT/gkr_trace.err ; grep "zk_gkr_prove_wide ms" $OUT/gkr_trace.err | tail -1
ZKB200_GKR_TABLES=1 ZKB200_TRACE=1 $B --workload gkr_wide --steps 2 --warmup 1 > /dev/null 2> $OUT/gkr_trace_tables.err ; grep "zk_gkr_prove_wide ms" $OUT/gkr_trace_tables.err | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"phase|eq_|eval_layer" -c 200 --csv --log-file $OUT/launches_gkr_builders.csv \
    python bench.py --workload gkr_wide --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/ncu_gkr.log 2>&1 ; echo "ncu gkr $?"
python - $OUT <<'PY'
import json,sys,glob,os
for f in sorted(glob.glob(sys.argv[1]+"/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print("%-26s value=%.6g %s ms=%.4f frac=%.3f launches=%s verified=%s" % (os.path.basename(f), d["value"], d["unit"], d["ms_per_step"], r.get("frac",0), d.get("gpu_launches"), d.get("verified")))
    except Exception as ex:
        print(f, "unreadable:", ex, open(f.replace('.json','.err')).read()[-300:])
PY
