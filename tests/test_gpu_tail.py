"""The device-resident round loop (csrc/devrounds.cuh): once tables x entries <= 2^tail_log, ONE persistent launch runs
every remaining round -- sums, fold, grid barrier, Fiat-Shamir transcript, challenge -- on the GPU.  Proofs must not depend
on where the hand-over happens: tail_log = 0 (every round host-driven), small (a single-block launch), and large (a
multi-block cooperative launch from round 0) all give the oracle's proof, limb for limb, and leave the host transcript in
the oracle's state.  Also forces small multi-block grids (ZKB200_GRID_CAP) so that the grid-wide column sums, the
arrive / release barrier and the blocks leaving the loop as the tables shrink are exercised with long grid-stride loops."""
import os

import numpy as np
import pytest

from test_gpu_parity import rand_table, sumpoly_handle

pytestmark = pytest.mark.gpu

TAIL_LOGS = [0, 2, 5, 13, 16, 24]
DEFAULT_TAIL_LOG = 20


@pytest.fixture
def restore_tail(ctx_for):
    yield
    for fid in (0, 1, 2):
        ctx_for(fid).set_tail_log(DEFAULT_TAIL_LOG)


def product_case(co, fid, P, D, n, seed):
    tabs = np.stack([np.stack([rand_table(co, fid, 1 << n, seed + p * D + d) for d in range(D)]) for p in range(P)])
    claimed = np.zeros(4, dtype=np.uint64)
    co.lib().zko_fe_sum(fid, co._p(co.sumpoly_reduce(fid, tabs)), 1 << n, co._p(claimed))
    return tabs, claimed


@pytest.mark.parametrize("fid", [0, 1, 2])
@pytest.mark.parametrize("P,D", [(2, 2), (2, 3), (3, 2), (4, 2)])
def test_product_proof_is_independent_of_the_hand_over(zk, co, ctx_for, restore_tail, fid, P, D):
    from zk_cryptography_research_implementations_b200 import sumcheck_protocol as scp
    from zk_cryptography_research_implementations_b200.transcripts import Transcript
    ctx = ctx_for(fid)
    for n in (1, 2, 3, 6, 9, 12, 15):
        tabs, claimed = product_case(co, fid, P, D, n, 5000 * P + 500 * D + 20 * n)
        tr_o = co.Transcript()
        tr_o.append(b"odd")                      # 3 bytes: the sponge position is not word aligned
        coeffs, ch, fin = co.product_prove(fid, tabs, claimed, tr_o)
        want_next = tr_o.sample_random_challenge()
        for tl in TAIL_LOGS:
            ctx.set_tail_log(tl)
            tr = Transcript()
            tr.append(b"odd")
            proof = scp.prove(sumpoly_handle(zk, ctx, tabs), claimed, tr)
            got = np.stack([p_.coefficients for p_ in proof.round_univariate_polynomials])
            assert np.array_equal(got, coeffs), (n, tl, "coefficients")
            assert np.array_equal(proof.random_challenges, ch), (n, tl, "challenges")
            assert np.array_equal(proof.final_values.reshape(P, D, 4), fin), (n, tl, "final values")
            assert tr.sample_random_challenge() == want_next, (n, tl, "transcript state after the proof")


@pytest.mark.parametrize("fid", [0, 2])
def test_f_times_g_hand_over(zk, co, ctx_for, restore_tail, fid):
    """the native P = 1 kernels (BASELINE config 3 shape) against the oracle's f*g + 0*0"""
    from zk_cryptography_research_implementations_b200.core import _ptr
    from zk_cryptography_research_implementations_b200.transcripts import Transcript
    ctx = ctx_for(fid)
    for n in (1, 2, 5, 10, 14, 17):
        f, g = rand_table(co, fid, 1 << n, 7300 + n), rand_table(co, fid, 1 << n, 7400 + n)
        z = np.zeros_like(f)
        tabs = np.stack([np.stack([f, g]), np.stack([z, z])])
        claimed = np.zeros(4, dtype=np.uint64)
        co.lib().zko_fe_sum(fid, co._p(co.sumpoly_reduce(fid, tabs)), 1 << n, co._p(claimed))
        coeffs, ch, fin = co.product_prove(fid, tabs, claimed, co.Transcript())
        host = np.ascontiguousarray(np.stack([f, g]))
        for tl in TAIL_LOGS:
            for flags in (0, 1):
                ctx.set_tail_log(tl)
                c2 = np.zeros((n, 3, 4), dtype=np.uint64); ch2 = np.zeros((n, 4), dtype=np.uint64); fin2 = np.zeros((2, 4), dtype=np.uint64)
                tr = Transcript()
                ctx.check(ctx.lib.zk_prove_product_host(ctx.h, _ptr(host), 1, 2, 1 << n, _ptr(claimed), tr.h, _ptr(c2), _ptr(ch2), _ptr(fin2), flags))
                assert np.array_equal(c2, coeffs) and np.array_equal(ch2, ch) and np.array_equal(fin2, fin[0]), (n, tl, flags)
        # ZK_FLAG_HOST_ROUNDS overrides the context setting
        ctx.set_tail_log(13)
        ctx.reset_stats()
        c2 = np.zeros((n, 3, 4), dtype=np.uint64); ch2 = np.zeros((n, 4), dtype=np.uint64); fin2 = np.zeros((2, 4), dtype=np.uint64)
        tr = Transcript()    # keep it alive across the call (a temporary would be freed before the library uses it)
        ctx.check(ctx.lib.zk_prove_product_host(ctx.h, _ptr(host), 1, 2, 1 << n, _ptr(claimed), tr.h, _ptr(c2), _ptr(ch2), _ptr(fin2), 16))
        assert np.array_equal(c2, coeffs) and np.array_equal(fin2, fin[0])
        assert ctx.stats()["round_launches"] == n            # one launch per round
        ctx.reset_stats()
        tr = Transcript()
        ctx.check(ctx.lib.zk_prove_product_host(ctx.h, _ptr(host), 1, 2, 1 << n, _ptr(claimed), tr.h, _ptr(c2), _ptr(ch2), _ptr(fin2), 0))
        assert np.array_equal(c2, coeffs) and np.array_equal(fin2, fin[0])
        # two tables: host-driven while they are longer than 2^13 / 2 (round 0 and rounds 1..n-12), then ONE launch for the rest
        assert ctx.stats()["round_launches"] == (1 if n <= 12 else n - 10)
        ctx.set_tail_log(DEFAULT_TAIL_LOG)
        ctx.reset_stats()
        tr = Transcript()
        ctx.check(ctx.lib.zk_prove_product_host(ctx.h, _ptr(host), 1, 2, 1 << n, _ptr(claimed), tr.h, _ptr(c2), _ptr(ch2), _ptr(fin2), 0))
        assert np.array_equal(c2, coeffs) and np.array_equal(fin2, fin[0])
        assert ctx.stats()["round_launches"] == 1            # the whole prove is one persistent launch


@pytest.mark.parametrize("fid", [0, 1, 2])
def test_plain_proof_is_independent_of_the_hand_over(zk, co, ctx_for, restore_tail, fid):
    from zk_cryptography_research_implementations_b200.sumcheck_protocol import Prover
    ctx = ctx_for(fid)
    for n in (0, 1, 2, 3, 7, 11, 14, 17):
        T = rand_table(co, fid, 1 << n, 8100 + n)
        claimed, rp, ch, fin = co.basic_prove(fid, T)
        for tl in TAIL_LOGS:
            ctx.set_tail_log(tl)
            proof = Prover.init(ctx, T).prove()
            assert np.array_equal(proof.initial_claimed_sum, claimed), (n, tl)
            assert np.array_equal(proof.round_univariate_polynomials, rp), (n, tl)
            assert np.array_equal(proof.challenges, ch) and np.array_equal(proof.final_evaluation, fin), (n, tl)


def test_tail_log_argument_checks(zk, ctx_for, restore_tail):
    ctx = ctx_for(0)
    assert ctx.tail_log() == DEFAULT_TAIL_LOG
    with pytest.raises(zk.ZkError):
        ctx.set_tail_log(33)
    with pytest.raises(zk.ZkError):
        ctx.set_tail_log(-1)
    ctx.set_tail_log(0)
    assert ctx.tail_log() == 0


@pytest.mark.parametrize("cap", ["2", "5", "37"])
@pytest.mark.parametrize("fid", [0, 2])
def test_multi_block_device_rounds_with_capped_grids(zk, co, cap, fid):
    """the persistent launch on 2, 5 and 37 blocks (ZKB200_GRID_CAP): long grid-stride loops, every block contributing to
    the grid accumulator, the barrier counting fewer blocks round by round -- product (2x2 and f*g + linear-free), plain"""
    from zk_cryptography_research_implementations_b200 import sumcheck_protocol as scp
    from zk_cryptography_research_implementations_b200.sumcheck_protocol import Prover
    from zk_cryptography_research_implementations_b200.transcripts import Transcript
    os.environ["ZKB200_GRID_CAP"] = cap
    try:
        ctx = zk.Context(fid, 0)
        assert ctx.tail_log() == DEFAULT_TAIL_LOG
        for (P, D, n) in [(2, 2, 13), (2, 3, 11), (3, 2, 12)]:
            tabs, claimed = product_case(co, fid, P, D, n, 9300 + 7 * fid + n)
            tr_o, tr = co.Transcript(), Transcript()
            tr_o.append(b"x" * 5)
            tr.append(b"x" * 5)
            coeffs, ch, fin = co.product_prove(fid, tabs, claimed, tr_o)
            ctx.reset_stats()
            proof = scp.prove(sumpoly_handle(zk, ctx, tabs), claimed, tr)
            assert ctx.stats()["round_launches"] == 1
            got = np.stack([p_.coefficients for p_ in proof.round_univariate_polynomials])
            assert np.array_equal(got, coeffs) and np.array_equal(proof.random_challenges, ch)
            assert np.array_equal(proof.final_values.reshape(P, D, 4), fin)
            assert tr.sample_random_challenge() == tr_o.sample_random_challenge()
        T = rand_table(co, fid, 1 << 15, 9400 + fid)
        pb = Prover.init(ctx, T).prove()
        claimed_b, rp, chb, finb = co.basic_prove(fid, T)
        assert np.array_equal(pb.round_univariate_polynomials, rp) and np.array_equal(pb.challenges, chb) and np.array_equal(pb.final_evaluation, finb)
        ctx.close()
    finally:
        del os.environ["ZKB200_GRID_CAP"]


@pytest.mark.parametrize("cap", ["3", "37"])
def test_multi_block_grids_on_small_tables(zk, co, cap):
    """a context whose grids are capped (test hook): several blocks with long grid-stride loops even on tables the
    oracle can check, host-driven rounds -- the grid-wide exact column sums against the oracle"""
    from zk_cryptography_research_implementations_b200 import sumcheck_protocol as scp
    from zk_cryptography_research_implementations_b200.sumcheck_protocol import Prover
    from zk_cryptography_research_implementations_b200.transcripts import Transcript
    os.environ["ZKB200_GRID_CAP"] = cap
    try:
        for fid, P, D in [(0, 2, 2), (2, 2, 3), (1, 4, 2)]:
            ctx = zk.Context(fid, 0)
            ctx.set_tail_log(0)
            n = 12
            tabs, claimed = product_case(co, fid, P, D, n, 9100 + 7 * fid)
            coeffs, ch, fin = co.product_prove(fid, tabs, claimed, co.Transcript())
            proof = scp.prove(sumpoly_handle(zk, ctx, tabs), claimed, Transcript())
            got = np.stack([p_.coefficients for p_ in proof.round_univariate_polynomials])
            assert np.array_equal(got, coeffs) and np.array_equal(proof.random_challenges, ch)
            assert np.array_equal(proof.final_values.reshape(P, D, 4), fin)
            T = rand_table(co, fid, 1 << 13, 9200 + fid)
            pb = Prover.init(ctx, T).prove()
            claimed_b, rp, chb, finb = co.basic_prove(fid, T)
            assert np.array_equal(pb.round_univariate_polynomials, rp) and np.array_equal(pb.final_evaluation, finb)
            ctx.close()
    finally:
        del os.environ["ZKB200_GRID_CAP"]


def test_wide_gkr_proof_is_independent_of_the_hand_over(zk, co, ctx_for, restore_tail):
    """the sparse two-phase GKR prover runs its phases through zk_prove_product (one product + one LINEAR table):
    the proof with the device tail equals the proof with host-driven rounds"""
    from zk_cryptography_research_implementations_b200 import gkr
    fid = 0
    ctx = ctx_for(fid)
    rng = np.random.default_rng(5)
    bits = [1, 4, 6, 6]
    layers = []
    for li in range(len(bits) - 1):
        n_out, n_in = 1 << bits[li], 1 << bits[li + 1]
        seen, gates = set(), []
        while len(gates) < 3 * max(n_out, n_in // 2):
            g = (int(rng.integers(0, n_in)), int(rng.integers(0, n_in)), int(rng.integers(0, n_out)), int(rng.integers(0, 2)))   # (left, right, out, op)
            if g not in seen:
                seen.add(g)
                gates.append(g)
        layers.append(gates)
    inputs = rand_table(co, fid, 1 << bits[-1], 77)
    proofs = []
    for tl in (0, 3, 13, 24):
        ctx.set_tail_log(tl)
        wc = gkr.WideCircuit(ctx, bits, layers)
        proofs.append(gkr.prove_wide(ctx, wc, inputs))
        wc.close()
    first = proofs[0]
    for p in proofs[1:]:
        assert np.array_equal(p.circuit_output, first.circuit_output) and np.array_equal(p.claimed_sum, first.claimed_sum)
        assert np.array_equal(p.wb_evaluations, first.wb_evaluations) and np.array_equal(p.wc_evaluations, first.wc_evaluations)
        for a, b in zip(p.sumcheck_proofs, first.sumcheck_proofs):
            assert np.array_equal(a.claimed_sum, b.claimed_sum) and np.array_equal(a.random_challenges, b.random_challenges)
            assert all(np.array_equal(x.coefficients, y.coefficients) for x, y in zip(a.round_univariate_polynomials, b.round_univariate_polynomials))
