"""debug: locate the first round where the GPU product prover departs from the oracle at mid sizes"""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import coracle as co
import zk_cryptography_research_implementations_b200 as zk
from zk_cryptography_research_implementations_b200.core import _ptr
from zk_cryptography_research_implementations_b200.transcripts import Transcript
fid = 0
ctx = zk.Context(fid, 0)
for n in (10, 12, 14, 16, 17):
    N = 1 << n
    f = ctx.generate(1, 0, N).download(); g = ctx.generate(1, 1, N).download()
    z = np.zeros_like(f)
    tabs = np.stack([np.stack([f, g]), np.stack([z, z])])
    claimed = np.zeros(4, dtype=np.uint64)
    co.lib().zko_fe_sum(fid, co._p(co.sumpoly_reduce(fid, tabs)), N, co._p(claimed))
    coeffs, ch, fin = co.product_prove(fid, tabs, claimed, co.Transcript())
    for flags in (0, 1):
        c2 = np.zeros((n, 3, 4), dtype=np.uint64); ch2 = np.zeros((n, 4), dtype=np.uint64); fin2 = np.zeros((2, 4), dtype=np.uint64)
        host = np.ascontiguousarray(np.stack([f, g]))
        tr = Transcript()
        ctx.check(ctx.lib.zk_prove_product_host(ctx.h, _ptr(host), 1, 2, N, _ptr(claimed), tr.h, _ptr(c2), _ptr(ch2), _ptr(fin2), flags))
        bad = [k for k in range(n) if not np.array_equal(c2[k], coeffs[k])]
        print("n", n, "flags", flags, "grid_cap", os.environ.get("ZKB200_GRID_CAP"), "first bad round", bad[:3], "final ok", np.array_equal(fin2, fin[0]), flush=True)
        if bad:
            k = bad[0]
            print("   which coeff differs:", [i for i in range(3) if not np.array_equal(c2[k, i], coeffs[k, i])])
