#!/bin/bash
# Round-2 session 3 (1 GPU): per-round timeline of the persistent round loop (measurement builds), 1 vs 2 blocks per SM;
# the inner-product MLE evaluate.   gpurun --timeout 1500 -- bash tools/r02_session3.sh
set -u
OUT=gpurun_out/r02_s3
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tail.py tests/test_gpu_parity.py tests/test_gpu_parity_large.py tests/test_gpu_verifiers.py -m gpu -q --timeout=600 -p no:cacheprovider -x > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" ; tail -3 $OUT/pytest_gpu.log
B="timeout 300 python bench.py --no-cpu --no-probe --no-extras --no-e2e"
PKG=zk_cryptography_research_implementations_b200
for v in timing timing2; do
  ZKB200_LIB=$PWD/$PKG/libzkb200_$v.so ZKB200_DEV_TIMING=1 $B --workload plain24 --steps 3 --warmup 2 > $OUT/plain24_$v.json 2> $OUT/plain24_$v.err ; echo "plain24 $v rc=$?"
  ZKB200_LIB=$PWD/$PKG/libzkb200_$v.so ZKB200_DEV_TIMING=1 $B --workload gkr_wide --steps 2 --warmup 1 > $OUT/gkr_wide_$v.json 2> $OUT/gkr_wide_$v.err ; echo "gkr_wide $v rc=$?"
  ZKB200_LIB=$PWD/$PKG/libzkb200_$v.so ZKB200_DEV_TIMING=1 ZKB200_TAIL_LOG=26 $B --log2 26 --steps 2 --warmup 1 > $OUT/product26_$v.json 2> $OUT/product26_$v.err ; echo "product26 $v rc=$?"
done
for tl in 13 20 24; do
  ZKB200_TAIL_LOG=$tl $B --workload plain24 --steps 20 --warmup 5 > $OUT/plain24_tl$tl.json 2> $OUT/plain24_tl$tl.err
  ZKB200_TAIL_LOG=$tl $B --workload gkr_wide --steps 5 --warmup 2 > $OUT/gkr_wide_tl$tl.json 2> $OUT/gkr_wide_tl$tl.err
done
$B --workload mle --log2 28 --sweep 20,22,24,26,30 --steps 5 --warmup 3 > $OUT/mle28.json 2> $OUT/mle28.err ; echo "mle rc=$?"
ZKB200_EVAL_FOLDS=1 $B --workload mle --log2 28 --sweep 20,24 --steps 5 --warmup 3 > $OUT/mle28_folds.json 2> $OUT/mle28_folds.err
for f in plain24_timing gkr_wide_timing product26_timing plain24_timing2 gkr_wide_timing2 product26_timing2; do echo "== $f"; grep devrounds $OUT/$f.err | tail -2 | cut -c1-6000; done
python - $OUT <<'PY'
import json,sys,glob,os
for f in sorted(glob.glob(sys.argv[1]+"/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print("%-26s value=%.6g %s ms=%.4f frac=%.3f launches=%s verified=%s" % (os.path.basename(f), d["value"], d["unit"], d["ms_per_step"], r.get("frac",0), d.get("gpu_launches"), d.get("verified")))
        if "sweep" in d: print("  sweep", [(s["log2_entries"], round(s["evaluate_ms"],4), round(s["evaluate_frac_hbm"],3)) for s in d["sweep"]], "pe", d["partial_evaluate"]["ms"])
    except Exception as ex:
        print(f, "unreadable:", ex, open(f.replace('.json','.err')).read()[-300:])
PY
