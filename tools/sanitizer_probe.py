#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): product and plain sumchecks with the hand-over to
the device tail at several points, multi-block grids on small tables, and a small wide-GKR prove.

    compute-sanitizer --tool racecheck python tools/sanitizer_probe.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("ZKB200_GRID_CAP", "5")     # several blocks with grid-stride loops even on small tables

import zk_cryptography_research_implementations_b200 as zk  # noqa: E402
from zk_cryptography_research_implementations_b200 import gkr, sumcheck_protocol as scp  # noqa: E402
from zk_cryptography_research_implementations_b200.polynomials import MultilinearPolynomial, ProductPolynomial, SumPolynomial  # noqa: E402
from zk_cryptography_research_implementations_b200.transcripts import Transcript  # noqa: E402


def main():
    fid = zk.BN254_FQ
    ctx = zk.Context(fid, 0)
    n = 11
    tabs = [[ctx.generate(3, 2 * p + d, 1 << n).download() for d in range(2)] for p in range(2)]
    claimed = zk.fe_from_ints(fid, [12345])[0]
    ref = None
    for tl in (0, 5, 13):
        ctx.set_tail_log(tl)
        sp = SumPolynomial([ProductPolynomial([MultilinearPolynomial.new(ctx, t) for t in prod]) for prod in tabs])
        proof = scp.prove(sp, claimed, Transcript())
        got = np.stack([p.coefficients for p in proof.round_univariate_polynomials])
        if ref is None:
            ref = got
        assert np.array_equal(got, ref), tl
    table = ctx.generate(4, 0, 1 << n).download()
    refp = None
    for tl in (0, 4, 13):
        ctx.set_tail_log(tl)
        p = scp.Prover.init(ctx, table).prove()
        if refp is None:
            refp = p.round_univariate_polynomials
        assert np.array_equal(p.round_univariate_polynomials, refp), tl
    rng = np.random.default_rng(1)
    bits = [1, 5, 5]
    layers = []
    for li in range(2):
        n_out, n_in = 1 << bits[li], 1 << bits[li + 1]
        layers.append([(int(rng.integers(0, n_in)), int(rng.integers(0, n_in)), o % n_out, int(rng.integers(0, 2))) for o in range(max(n_out, n_in))])
        layers[-1] = sorted(set(layers[-1]))
    inputs = ctx.generate(5, 0, 1 << bits[-1]).download()
    ctx.set_tail_log(13)
    wc = gkr.WideCircuit(ctx, bits, layers)
    a = gkr.prove_wide(ctx, wc, inputs)
    ctx.set_tail_log(0)
    b = gkr.prove_wide(ctx, wc, inputs)
    assert np.array_equal(a.claimed_sum, b.claimed_sum)
    wc.close()
    vals = MultilinearPolynomial.new(ctx, table).evaluate(table[:n])
    assert vals.shape == (4,)
    ctx.close()
    print("sanitizer probe ok")


if __name__ == "__main__":
    main()
