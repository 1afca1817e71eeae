"""Quick look on the GPU box: arithmetic probes (mont_mul / fold / mul_acc at 1, 2, 4 blocks per SM) and round-kernel
timings for a few (P, D, n).  Developer tool, not a benchmark of record: `gpurun -- python tools/gpu_first_look.py`."""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import zk_cryptography_research_implementations_b200 as zk
from zk_cryptography_research_implementations_b200.core import _ptr
from zk_cryptography_research_implementations_b200.transcripts import Transcript

out = {}
for fid in (0, 2):
    ctx = zk.Context(fid, 0)
    for kind, name in ((0, "mont_mul"), (1, "fold"), (2, "mul_acc")):
        for bps in (1, 2, 4):
            ops, ms = C.c_double(), C.c_double()
            ctx.check(ctx.lib.zk_arith_probe(ctx.h, kind, 2000, bps, C.byref(ops), C.byref(ms)))
            out["probe_f%d_%s_bps%d" % (fid, name, bps)] = {"Gops": ops.value / 1e9, "ms": ms.value}
            print(fid, name, bps, "%.1f Gop/s" % (ops.value / 1e9), flush=True)
    for (P, D, n) in ((1, 1, 24), (1, 2, 24), (1, 2, 26), (2, 2, 24)):
        N = 1 << n
        tabs = [ctx.generate(0xB200, i, N) for i in range(P * D)]
        arr = (C.c_void_p * (P * D))(*[t.release() for t in tabs])
        h = C.c_void_p()
        ctx.check(ctx.lib.zk_sumpoly_create(ctx.h, arr, P, D, C.byref(h)))
        for rep in range(3):
            for i in range(P * D):
                ctx.check(ctx.lib.zk_table_regenerate(ctx.h, ctx.lib.zk_sumpoly_table(h, i), 0xB200, i, N, 0, 1))
            ctx.synchronize()
            ctx.set_profiling(True); ctx.reset_stats()
            coeffs = np.zeros((n, D + 1, 4), dtype=np.uint64); ch = np.zeros((n, 4), dtype=np.uint64); fin = np.zeros((P * D, 4), dtype=np.uint64)
            claimed = np.zeros(4, dtype=np.uint64)
            t0 = time.perf_counter()
            if D == 1:
                rp = np.zeros((n, 2, 4), dtype=np.uint64)
                ctx.check(ctx.lib.zk_prove_basic_device(ctx.h, ctx.lib.zk_sumpoly_table(h, 0), _ptr(claimed), _ptr(rp), _ptr(ch), _ptr(fin), 2))
            else:
                tr = Transcript()
                ctx.check(ctx.lib.zk_prove_product(ctx.h, h, _ptr(claimed), tr.h, _ptr(coeffs), _ptr(ch), _ptr(fin), 0))
            dt = time.perf_counter() - t0
            st = ctx.stats()
        key = "prove_f%d_P%dD%d_n%d" % (fid, P, D, n)
        out[key] = {"wall_ms": dt * 1e3, "round_kernel_ms": st["round_ms"], "round_GBps": st["round_bytes"] / st["round_ms"] / 1e6,
                    "launches": st["launches"], "Gelem_s": N / dt / 1e9}
        print(key, out[key], flush=True)
        ctx.lib.zk_sumpoly_free(ctx.h, h)
    ctx.close()
json.dump(out, open("gpurun_out/first_look.json", "w"), indent=1)
