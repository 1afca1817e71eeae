#!/bin/bash
# Round-2 session 1 (1 GPU): parity tests after the hygiene / GPU-CSR / verifier work, the driver's default bench command
# with its extra_workloads, and the ncu launch list of the GKR workload (where the time goes before the persistent kernel).
#   gpurun --timeout 1200 -- bash tools/r02_session1.sh
set -u
OUT=gpurun_out/r02_s1
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -x > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" ; tail -5 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_default.json 2> $OUT/bench_default.err ; echo "bench default $?"
tail -3 $OUT/bench_default.err
ZKB200_TRACE=1 timeout 300 python bench.py --workload gkr_wide --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/gkr_trace.json 2> $OUT/gkr_trace.err ; echo "gkr trace $?"
grep "zk_gkr_prove_wide ms" $OUT/gkr_trace.err | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_gkr_wide.csv \
    python bench.py --workload gkr_wide --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/ncu_gkr_wide.log 2>&1 ; echo "ncu gkr $?"
python - "$OUT/bench_default.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    def show(d):
        r=d.get("roofline") or {}; e=d.get("e2e") or {}
        print((d.get("config") or {}).get("workload","?")[:40], "value=%.5g %s ms=%.3f frac=%.3f e2e=%s verified=%s" % (d["value"], d["unit"], d["ms_per_step"], r.get("frac",0), e.get("value"), d.get("verified")), d.get("verify"))
    show(d)
    for x in d.get("extra_workloads",[]):
        if "error" in x: print("EXTRA ERROR", x)
        else:
            show(x)
            if "sweep" in x: print("  sweep", [(s["log2_entries"], round(s["evaluate_ms"],3), round(s["evaluate_frac_hbm"],3), round(s["partial_evaluate_ms"],3), round(s["partial_evaluate_frac_hbm"],3)) for s in x["sweep"]])
            if "circuit_setup_s" in (x.get("config") or {}): print("  circuit_setup_s", x["config"]["circuit_setup_s"], "verify_ms", x.get("verify_ms"))
except Exception as ex:
    print("unreadable:", ex)
PY
