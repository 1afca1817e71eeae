"""Limb-for-limb parity with the oracle at sizes where the HBM-regime code path runs: multi-iteration grid-stride
loops, `prefetch.global.L2`, full 592-block grids, the grid-wide column reduction over hundreds of blocks -- the
path every bench number is measured on.  The oracle (oracle/zkoracle.c, the reference's pass structure) proves these
sizes in seconds once its data-parallel loops run on all host cores (`zko_set_threads`; outputs are identical to the
single-threaded run, tests/test_oracle.py).

Reference: sumcheck_protocol/src/gkr_sumcheck/sumcheck_gkr_protocol.rs:24-67 (product prover),
sumcheck_protocol/src/basic_sumcheck/prover.rs:35-71 (plain prover) and its own 2^20 test
(sumcheck_protocol/src/basic_sumcheck/protocol.rs:41-55, `vec![Fr::from(3); 1 << 20]`)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 0xB200


@pytest.fixture
def all_cores(co):
    co.set_threads(os.cpu_count() or 1)
    yield
    co.set_threads(1)


def _gpu_product_prove(zk, ctx, tabs_dev, P, D, claimed, flags=0):
    from zk_cryptography_research_implementations_b200.core import _ptr
    from zk_cryptography_research_implementations_b200.transcripts import Transcript
    n = len(tabs_dev[0]).bit_length() - 1
    arr = (C.c_void_p * (P * D))(*[t.release() for t in tabs_dev])
    h = C.c_void_p()
    ctx.check(ctx.lib.zk_sumpoly_create(ctx.h, arr, P, D, C.byref(h)))
    coeffs = np.zeros((n, D + 1, 4), dtype=np.uint64)
    ch = np.zeros((n, 4), dtype=np.uint64)
    fin = np.zeros((P * D, 4), dtype=np.uint64)
    tr = Transcript()
    try:
        ctx.check(ctx.lib.zk_prove_product(ctx.h, h, _ptr(claimed), tr.h, _ptr(coeffs), _ptr(ch), _ptr(fin), flags))
    finally:
        ctx.lib.zk_sumpoly_free(ctx.h, h)
    return coeffs, ch, fin, tr


@pytest.fixture
def tail_logs(ctx_for):
    """both regimes of the round loop: 13 = host-driven per-round kernels (592-block grids, HBM streaming) down to 2^12
    entries, then a single-block device launch; 24 = the whole prove in ONE persistent multi-block launch"""
    yield (13, 24)
    for fid in (0, 1, 2):
        ctx_for(fid).set_tail_log(20)


@pytest.mark.parametrize("n", [20, 22])
def test_f_times_g_limb_for_limb_at_hbm_sizes(zk, co, ctx_for, all_cores, tail_logs, n):
    """BASELINE configs[2] in miniature (f*g on BN254 Fq, the bench's own seeded tables): every coefficient, challenge and
    folded value of the GPU proof equals the oracle's, which proves the reference's form f*g + 0*0."""
    fid = 0
    ctx = ctx_for(fid)
    N = 1 << n
    host = np.zeros((2, 2, N, 4), dtype=np.uint64)
    dev = [ctx.generate(SEED, i, N) for i in range(2)]
    for i in range(2):
        host[0, i] = dev[i].download()
    claimed = np.zeros(4, dtype=np.uint64)
    co.lib().zko_fe_sum(fid, co._p(co.sumpoly_reduce(fid, host)), N, co._p(claimed))
    want_coeffs, want_ch, want_fin = co.product_prove(fid, host, claimed, co.Transcript())
    for tl in tail_logs:
        ctx.set_tail_log(tl)
        for flags in (0, 1):     # s(1) derived from the running claim / summed directly
            tabs = [ctx.generate(SEED, i, N) for i in range(2)]
            coeffs, ch, fin, _ = _gpu_product_prove(zk, ctx, tabs, 1, 2, claimed, flags)
            assert np.array_equal(coeffs, want_coeffs), "round polynomials differ from the oracle (tail_log=%d flags=%d)" % (tl, flags)
            assert np.array_equal(ch, want_ch)
            assert np.array_equal(fin, want_fin[0])
    ok, _, _ = co.product_verify(fid, claimed, want_coeffs, co.Transcript())
    assert ok


def test_gkr_shaped_2x2_limb_for_limb_at_2p20(zk, co, ctx_for, all_cores, tail_logs):
    """(P, D) = (2, 2) -- add*(Wb+Wc) + mul*(Wb*Wc), the GKR layer shape -- over four 2^20-entry tables, BLS12-381 Fr"""
    fid = 2
    ctx = ctx_for(fid)
    N = 1 << 20
    dev = [ctx.generate(SEED + 1, i, N) for i in range(4)]
    host = np.stack([d.download() for d in dev]).reshape(2, 2, N, 4)
    claimed = np.zeros(4, dtype=np.uint64)
    co.lib().zko_fe_sum(fid, co._p(co.sumpoly_reduce(fid, host)), N, co._p(claimed))
    tr_o = co.Transcript()
    want_coeffs, want_ch, want_fin = co.product_prove(fid, host, claimed, tr_o)
    want_next = tr_o.sample_random_challenge()
    for tl in tail_logs:
        ctx.set_tail_log(tl)
        coeffs, ch, fin, tr = _gpu_product_prove(zk, ctx, [ctx.generate(SEED + 1, i, N) for i in range(4)], 2, 2, claimed)
        assert np.array_equal(coeffs, want_coeffs) and np.array_equal(ch, want_ch), tl
        assert np.array_equal(fin.reshape(2, 2, 4), want_fin)
        assert tr.sample_random_challenge() == want_next      # transcripts end in the same state


def _plain_limb_for_limb(zk, co, ctx, fid, table, tail_logs=(24,)):
    from zk_cryptography_research_implementations_b200.sumcheck_protocol import Prover
    c, rp, ch, fin = co.basic_prove(fid, table)
    for tl in tail_logs:
        ctx.set_tail_log(tl)
        proof = Prover.init(ctx, table).prove()
        assert np.array_equal(proof.initial_claimed_sum, c)
        assert np.array_equal(proof.round_univariate_polynomials, rp), tl
        assert np.array_equal(proof.challenges, ch)
        assert np.array_equal(proof.final_evaluation, fin)
    assert co.basic_verify(fid, table, proof.initial_claimed_sum, proof.round_univariate_polynomials)


def test_plain_sumcheck_limb_for_limb_at_2p20(zk, co, ctx_for, all_cores, tail_logs):
    """the reference's own largest test input, vec![Fr::from(3); 1 << 20] (protocol.rs:41-55), and a seeded random table"""
    fid = 2
    ctx = ctx_for(fid)
    N = 1 << 20
    _plain_limb_for_limb(zk, co, ctx, fid, np.tile(zk.fe_from_ints(fid, [3]), (N, 1)), tail_logs)
    _plain_limb_for_limb(zk, co, ctx, fid, ctx.generate(SEED, 0, N).download(), tail_logs)


def test_plain_sumcheck_limb_for_limb_at_2p24(zk, co, ctx_for, all_cores, tail_logs):
    """BASELINE configs[1] at full size (2^24 entries, BLS12-381 Fr, 512 MiB), the whole Prover::prove including the
    reference-mandated Keccak absorb of the table on both sides"""
    fid = 2
    ctx = ctx_for(fid)
    _plain_limb_for_limb(zk, co, ctx, fid, ctx.generate(SEED, 0, 1 << 24).download(), tail_logs)


@pytest.mark.parametrize("n", [20, 23])
def test_mle_evaluate_and_fold_against_the_oracle_at_hbm_sizes(zk, co, ctx_for, all_cores, n):
    """MultilinearPolynomial::evaluate / partial_evaluate (evaluation_form.rs:21-33,61-106) on tables past L2-resident
    sizes: the value and the whole folded table, limb for limb"""
    from zk_cryptography_research_implementations_b200.polynomials import MultilinearPolynomial as MLE
    fid = 0
    ctx = ctx_for(fid)
    N = 1 << n
    t = ctx.generate(SEED, 3, N)
    host = t.download()
    rs = ctx.generate(SEED, 99, 64).download()[:n]
    assert np.array_equal(MLE(ctx, t).evaluate(rs), co.mle_evaluate(fid, host, rs))
    for var in (0, n // 2, n - 1):
        f = MLE(ctx, ctx.upload(host))
        f.partial_evaluate_in_place(var, rs[0])
        assert np.array_equal(f.evaluated_values, co.mle_partial_evaluate(fid, host, var, rs[0])), var
