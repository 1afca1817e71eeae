#!/bin/bash
# Round-2 session 9 (1 GPU): the full GPU test suite on the final tree, the driver's default command, ncu launch lists and
# `--set full` captures of the dominant kernels (summaries go to profiles/r02/).
set -u
OUT=gpurun_out/r02_s9
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --timeout=800 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" ; tail -3 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err ; echo "bench default (driver's command) rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err ; echo "bench reference rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_product30.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-probe --no-extras > $OUT/ncu_product30.log 2>&1 ; echo "ncu launches product30 $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_plain24.csv \
    python bench.py --workload plain24 --steps 1 --warmup 1 --no-e2e --no-cpu --no-probe > $OUT/ncu_plain24.log 2>&1 ; echo "ncu launches plain24 $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/launches_gkr_wide.csv \
    python bench.py --workload gkr_wide --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/ncu_gkr_wide.log 2>&1 ; echo "ncu launches gkr $?"
# --set full captures: one fold_evals launch at 2^27 (the bench line's `traffic`), the round-0 kernel, the persistent round loop, the evaluate kernel
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"fold_evals_kernel" -s 1 -c 1 -o $OUT/fold_evals_2p27 -f \
    python bench.py --log2 27 --steps 1 --warmup 0 --no-e2e --no-cpu --no-probe --no-extras > $OUT/ncu_full_fold.log 2>&1 ; echo "ncu full fold_evals $?"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"round_evals_kernel" -c 1 -o $OUT/round_evals_2p27 -f \
    python bench.py --log2 27 --steps 1 --warmup 0 --no-e2e --no-cpu --no-probe --no-extras > $OUT/ncu_full_round.log 2>&1 ; echo "ncu full round_evals $?"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"sumcheck_rounds_kernel" -c 1 -o $OUT/sumcheck_rounds_plain24 -f \
    python bench.py --workload plain24 --steps 1 --warmup 0 --no-e2e --no-cpu --no-probe > $OUT/ncu_full_rounds.log 2>&1 ; echo "ncu full sumcheck_rounds $?"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"mle_inner_kernel" -c 1 -o $OUT/mle_inner_2p28 -f \
    python bench.py --workload mle --log2 28 --steps 1 --warmup 0 --no-cpu > $OUT/ncu_full_mle.log 2>&1 ; echo "ncu full mle_inner $?"
for r in fold_evals_2p27 round_evals_2p27 sumcheck_rounds_plain24 mle_inner_2p28; do
  [ -f $OUT/$r.ncu-rep ] && python tools/ncu_summary.py $OUT/$r.ncu-rep > $OUT/${r}_ncu_full_summary.csv 2>/dev/null
done
ls -la $OUT | head -40
python - "$OUT/bench_default.json" <<'PY'
import json,sys
try:
    d=[json.loads(l) for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1]
    def show(d):
        r=d.get("roofline") or {}; e=d.get("e2e") or {}
        print((d.get("config") or {}).get("workload","?")[:40], "value=%.5g %s ms=%.4f frac=%.3f e2e=%s verified=%s" % (d["value"], d["unit"], d["ms_per_step"], r.get("frac",0), e.get("value"), d.get("verified")))
    show(d)
    for x in d.get("extra_workloads",[]):
        if "error" in x: print("EXTRA ERROR", x)
        else:
            show(x)
            if "sweep" in x: print("  sweep", [(s["log2_entries"], round(s["evaluate_ms"],4), round(s["evaluate_frac_hbm"],3), round(s["partial_evaluate_ms"],4), round(s["partial_evaluate_frac_hbm"],3)) for s in x["sweep"]])
            if "circuit_setup_s" in (x.get("config") or {}): print("  circuit_setup_s", x["config"]["circuit_setup_s"], "verify_ms", x.get("verify_ms"))
except Exception as ex:
    print("unreadable:", ex)
PY
