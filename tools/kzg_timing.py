#!/usr/bin/env python3
"""Wall-clock timing of the multilinear-KZG calls (setup, commit, open) at a few sizes; prints one JSON line per size."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zk_cryptography_research_implementations_b200 as zk  # noqa: E402
from zk_cryptography_research_implementations_b200 import multilinear_kzg as kzg  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [16, 20, 22]
    ctx = zk.Context(zk.BLS12_381_FR, 0)
    for n in sizes:
        taus = zk.fe_from_ints(zk.BLS12_381_FR, [0x1234567 * (i + 3) ** 5 for i in range(n)])
        opening = zk.fe_from_ints(zk.BLS12_381_FR, [0x7654321 * (i + 5) ** 7 for i in range(n)])
        t0 = time.perf_counter()
        setup = kzg.TrustedSetup.initialize_setup(ctx, taus)
        ctx.synchronize()
        t_setup = time.perf_counter() - t0
        table = ctx.generate(0xB200, 3, 1 << n)
        ctx.synchronize()
        reps = 3
        kzg.MultilinearKZG.commit_to_polynomial(table, setup)
        t0 = time.perf_counter()
        for _ in range(reps):
            c = kzg.MultilinearKZG.commit_to_polynomial(table, setup)
        t_commit = (time.perf_counter() - t0) / reps
        kzg.MultilinearKZG.open_and_prove(table, setup, opening)
        t0 = time.perf_counter()
        for _ in range(reps):
            pr = kzg.MultilinearKZG.open_and_prove(table, setup, opening)
        t_open = (time.perf_counter() - t0) / reps
        # a constant table: every scalar in one bucket of the lowest window (the reference's own test tables look like this)
        const = ctx.upload(np.tile(zk.fe_from_int(zk.BLS12_381_FR, 3), (1 << n, 1)))
        kzg.MultilinearKZG.commit_to_polynomial(const, setup)
        t0 = time.perf_counter()
        cc = kzg.MultilinearKZG.commit_to_polynomial(const, setup)
        t_const = time.perf_counter() - t0
        print(json.dumps({"n": n, "const_commit_ms": round(t_const * 1e3, 3), "const_commit_x": hex(int(cc[0])), "setup_s": round(t_setup, 4), "commit_ms": round(t_commit * 1e3, 3), "open_ms": round(t_open * 1e3, 3),
                          "commit_Mpoints_per_s": round((1 << n) / t_commit / 1e6, 3), "commit_x": hex(int(c[0]))}), flush=True)
        setup.release()


if __name__ == "__main__":
    main()
