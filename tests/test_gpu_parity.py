"""Parity of the CUDA path (through the C-ABI) with the oracle: bit-exact on every limb.
Run on the B200 box:  python -m pytest tests -m gpu"""
import random

import numpy as np
import pytest

from conftest import FIELDS

pytestmark = pytest.mark.gpu

ALL_FIELDS = [0, 1, 2]


def rand_table(co, fid, n, seed, edge=True):
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    out = np.zeros((n, 4), dtype=np.uint64)
    import ctypes as C
    for i in range(n):
        co.lib().zko_fe_from_le_bytes_mod_order(fid, raw[i].ctypes.data_as(C.POINTER(C.c_uint8)), 32, out[i].ctypes.data_as(co.u64p))
    if edge and n >= 4:
        import pyoracle as po
        p = po.P[{v: k for k, v in FIELDS.items()}[fid]]
        out[0] = co.from_ints(fid, [0])[0]
        out[1] = co.from_ints(fid, [p - 1])[0]
        out[n - 1] = co.from_ints(fid, [1])[0]
    return out


def sumpoly_handle(zk, ctx, tables):
    """tables (P, D, n, 4) -> python-side SumPolynomial on the device"""
    from zk_cryptography_research_implementations_b200.polynomials import MultilinearPolynomial, ProductPolynomial, SumPolynomial
    return SumPolynomial([ProductPolynomial([MultilinearPolynomial.new(ctx, t) for t in prod]) for prod in tables])


# ------------------------------------------------------------------------------------------ tables
@pytest.mark.parametrize("fid", ALL_FIELDS)
def test_generate_matches_host_model(zk, co, ctx_for, fid):
    ctx = ctx_for(fid)
    for (n, first, step, tid) in [(1, 0, 1, 0), (64, 0, 1, 3), (32, 5, 8, 1), (1024, 0, 1, 0)]:
        t = ctx.generate(0xB200, tid, n, first, step)
        want = zk.synthetic_table_ints(fid, 0xB200, tid, n, first, step)
        assert co.to_ints(fid, t.download()) == want


def test_upload_download_clone_roundtrip(zk, co, ctx_for):
    ctx = ctx_for(0)
    T = rand_table(co, 0, 256, 1)
    t = ctx.upload(T)
    assert len(t) == 256 and np.array_equal(t.download(), T)
    c = t.clone()
    assert np.array_equal(c.download(), T)
    with pytest.raises(zk.ReferencePanic, match="Evaluated values must be a power of 2"):
        ctx.upload(T[:6])


# ------------------------------------------------------------------------------------------ MLE
@pytest.mark.parametrize("fid", ALL_FIELDS)
def test_partial_evaluate_every_variable(zk, co, ctx_for, fid):
    from zk_cryptography_research_implementations_b200.polynomials import MultilinearPolynomial as MLE
    ctx = ctx_for(fid)
    for n in (1, 2, 3, 6, 10):
        T = rand_table(co, fid, 1 << n, 10 + n)
        poly = MLE.new(ctx, T)
        for var in range(n):
            for r in (rand_table(co, fid, 1, 99 + var, edge=False)[0], co.from_ints(fid, [0])[0], co.from_ints(fid, [1])[0]):
                got = MLE.partial_evaluate(poly, var, r).evaluated_values
                assert np.array_equal(got, co.mle_partial_evaluate(fid, T, var, r)), (n, var)
        assert np.array_equal(poly.evaluated_values, T)   # partial_evaluate does not modify its input


@pytest.mark.parametrize("fid", ALL_FIELDS)
def test_evaluate_and_prefix_evaluate(zk, co, ctx_for, fid):
    from zk_cryptography_research_implementations_b200.polynomials import MultilinearPolynomial as MLE
    ctx = ctx_for(fid)
    for n in (0, 1, 2, 3, 4, 5, 7, 11, 13):
        T = rand_table(co, fid, 1 << n, 20 + n)
        rs = rand_table(co, fid, max(n, 1), 30 + n, edge=False)[:n]
        poly = MLE.new(ctx, T)
        assert np.array_equal(poly.evaluate(rs), co.mle_evaluate(fid, T, rs)), n
        for k in range(0, n, 3):    # fewer values than variables: entry 0 of the partially folded table
            assert np.array_equal(poly.evaluate(rs[:k]), co.mle_evaluate(fid, T, rs[:k]))
        assert np.array_equal(poly.evaluated_values, T)


@pytest.mark.parametrize("fid", ALL_FIELDS)
def test_convert_to_bytes_and_elementwise_ops(zk, co, ctx_for, fid):
    from zk_cryptography_research_implementations_b200.polynomials import MultilinearPolynomial as MLE
    ctx = ctx_for(fid)
    A, B = rand_table(co, fid, 64, 40), rand_table(co, fid, 64, 41)
    a, b = MLE.new(ctx, A), MLE.new(ctx, B)
    assert a.convert_to_bytes() == co.mle_to_bytes(fid, A)
    s = rand_table(co, fid, 1, 42, edge=False)[0]
    want = np.stack([co.fe_op("mul", fid, x, s) for x in A])
    assert np.array_equal(a.scalar_mul(s).evaluated_values, want)
    want = np.stack([co.fe_op("add", fid, x, y) for x, y in zip(A, B)])
    assert np.array_equal(MLE.add_polynomials(a, b).evaluated_values, want)
    wb, wc = MLE.new(ctx, A[:8]), MLE.new(ctx, B[:8])
    assert np.array_equal(MLE.polynomial_tensor_add(wb, wc).evaluated_values, co.tensor(fid, "add", A[:8], B[:8]))
    assert np.array_equal(MLE.polynomial_tensor_mul(wb, wc).evaluated_values, co.tensor(fid, "mul", A[:8], B[:8]))
    with pytest.raises(zk.ReferencePanic, match="Different polynomial length"):
        MLE.polynomial_tensor_mul(wb, MLE.new(ctx, B[:4]))


# ------------------------------------------------------------------------------------------ round kernels
@pytest.mark.parametrize("fid", ALL_FIELDS)
@pytest.mark.parametrize("P,D", [(2, 2), (2, 3), (3, 2), (4, 2)])
def test_round_evals_and_fused_fold(zk, co, ctx_for, fid, P, D):
    """generate_round_univariate + partial_evaluate, round by round, against the oracle"""
    import ctypes as C
    from zk_cryptography_research_implementations_b200.core import _ptr
    ctx = ctx_for(fid)
    for n in (1, 2, 5, 9):
        tabs = np.stack([np.stack([rand_table(co, fid, 1 << n, 1000 * P + 100 * D + 10 * n + p * D + d) for d in range(D)]) for p in range(P)])
        sp = sumpoly_handle(zk, ctx, tabs)
        h = sp._device_sumpoly(clone=False)
        cur = tabs
        ev = np.zeros((D + 1, 4), dtype=np.uint64)
        ctx.check(ctx.lib.zk_sumcheck_round_evals(ctx.h, h, _ptr(ev)))
        assert np.array_equal(ev, co.generate_round_univariate(fid, cur)), ("round0", n)
        for k in range(n):
            r = rand_table(co, fid, 1, 7 * n + k, edge=False)[0]
            cur = np.stack([np.stack([co.mle_partial_evaluate(fid, cur[p, d], 0, r) for d in range(D)]) for p in range(P)])
            if cur.shape[2] >= 2:
                ctx.check(ctx.lib.zk_sumcheck_fold_and_evals(ctx.h, h, _ptr(r), _ptr(ev)))
                assert np.array_equal(ev, co.generate_round_univariate(fid, cur)), ("round", n, k)
            else:
                ctx.check(ctx.lib.zk_sumcheck_fold_and_evals(ctx.h, h, _ptr(r), None))
            for i in range(P * D):
                th = ctx.lib.zk_sumpoly_table(h, i)
                got = np.zeros((cur.shape[2], 4), dtype=np.uint64)
                ctx.check(ctx.lib.zk_table_download(ctx.h, th, _ptr(got)))
                assert np.array_equal(got, cur[i // D, i % D]), ("table", n, k, i)
        ctx.lib.zk_sumpoly_free(ctx.h, h)


# ------------------------------------------------------------------------------------------ plain sumcheck
@pytest.mark.parametrize("fid", ALL_FIELDS)
def test_basic_sumcheck_bit_exact(zk, co, ctx_for, fid):
    from zk_cryptography_research_implementations_b200.sumcheck_protocol import Prover
    ctx = ctx_for(fid)
    for n in list(range(0, 9)) + [12, 15]:
        T = rand_table(co, fid, 1 << n, 50 + n)
        proof = Prover.init(ctx, T).prove()
        claimed, rp, ch, fin = co.basic_prove(fid, T)
        assert np.array_equal(proof.initial_claimed_sum, claimed), n
        assert np.array_equal(proof.round_univariate_polynomials, rp), n
        assert np.array_equal(proof.challenges, ch) and np.array_equal(proof.final_evaluation, fin)
        assert co.basic_verify(fid, proof.initial_polynomial, proof.initial_claimed_sum, proof.round_univariate_polynomials)


def test_basic_sumcheck_golden(zk, co, ctx_for, golden):
    from zk_cryptography_research_implementations_b200.sumcheck_protocol import Prover
    cases = [golden["appendix_b"]["basic"]] + golden["generated"]["basic"]
    for e in cases:
        fid = FIELDS[e["field"]]
        proof = Prover.init(ctx_for(fid), zk.fe_from_ints(fid, e["table"])).prove()
        assert zk.fe_to_ints(fid, proof.initial_claimed_sum) == [e["claimed_sum"]]
        assert [zk.fe_to_ints(fid, r) for r in proof.round_univariate_polynomials] == e["round_polys"]
        assert zk.fe_to_ints(fid, proof.challenges) == e["challenges"]
        assert zk.fe_to_ints(fid, proof.final_evaluation) == [e["final"]]


def test_reference_basic_round_trips(zk, co, ctx_for, golden):
    """protocol.rs:28-116 -- including the reference's largest own test, vec![Fr::from(3); 1 << 20]"""
    from zk_cryptography_research_implementations_b200.sumcheck_protocol import Prover
    for e in golden["reference_kats"]["basic_round_trips"]:
        fid = FIELDS[e["field"]]
        table = e.get("table") or [e["constant"]] * (1 << e["log2"])
        if "constant" in e:
            T = np.tile(zk.fe_from_ints(fid, [e["constant"]]), (1 << e["log2"], 1))
        else:
            T = zk.fe_from_ints(fid, table)
        proof = Prover.init(ctx_for(fid), T).prove()
        assert co.basic_verify(fid, T, proof.initial_claimed_sum, proof.round_univariate_polynomials), e["src"]
    for e in golden["reference_kats"]["basic_claimed_sum"]:
        fid = FIELDS[e["field"]]
        proof = Prover.init(ctx_for(fid), zk.fe_from_ints(fid, e["table"])).prove()
        assert zk.fe_to_ints(fid, proof.initial_claimed_sum) == [e["sum"]]


# ------------------------------------------------------------------------------------------ product sumcheck
@pytest.mark.parametrize("fid", ALL_FIELDS)
@pytest.mark.parametrize("P,D", [(2, 2), (2, 3), (3, 2), (4, 2)])
@pytest.mark.parametrize("flags", [0, 1])
def test_product_sumcheck_bit_exact(zk, co, ctx_for, fid, P, D, flags):
    from zk_cryptography_research_implementations_b200 import sumcheck_protocol as scp
    from zk_cryptography_research_implementations_b200.transcripts import Transcript
    ctx = ctx_for(fid)
    for n in (1, 2, 3, 6, 10, 13):
        tabs = np.stack([np.stack([rand_table(co, fid, 1 << n, 2000 * P + 200 * D + 10 * n + p * D + d) for d in range(D)]) for p in range(P)])
        claimed = np.zeros(4, dtype=np.uint64)
        co.lib().zko_fe_sum(fid, co._p(co.sumpoly_reduce(fid, tabs)), 1 << n, co._p(claimed))
        tr_o, tr_g = co.Transcript(), Transcript()
        tr_o.append(b"context"); tr_g.append(b"context")
        coeffs, ch, fin = co.product_prove(fid, tabs, claimed, tr_o)
        proof = scp.prove(sumpoly_handle(zk, ctx, tabs), claimed, tr_g, flags)
        got = np.stack([p_.coefficients for p_ in proof.round_univariate_polynomials])
        assert np.array_equal(got, coeffs), (n, "coefficients")
        assert np.array_equal(proof.random_challenges, ch)
        assert np.array_equal(proof.final_values.reshape(P, D, 4), fin)
        # transcripts end in the same state
        assert tr_o.sample_random_challenge() == tr_g.sample_random_challenge()
        # the restated reference verifier accepts the proof when it replays the same transcript prefix
        tr_v = co.Transcript()
        tr_v.append(b"context")
        ok, _, _ = co.product_verify(fid, claimed, got, tr_v)
        assert ok


@pytest.mark.parametrize("fid", [0, 2])
def test_f_times_g_is_the_oracle_form_fg_plus_0x0(zk, co, ctx_for, fid):
    """BASELINE config 3 (P = 1): the reference panics on a single product, so the oracle is posed
    f*g + 0*0; the native P = 1 kernel must give the same round polynomials."""
    import ctypes as C
    from zk_cryptography_research_implementations_b200.core import _ptr
    from zk_cryptography_research_implementations_b200.transcripts import Transcript
    ctx = ctx_for(fid)
    for n in (1, 4, 11):
        f, g = rand_table(co, fid, 1 << n, 300 + n), rand_table(co, fid, 1 << n, 400 + n)
        z = np.zeros_like(f)
        tabs = np.stack([np.stack([f, g]), np.stack([z, z])])
        claimed = np.zeros(4, dtype=np.uint64)
        co.lib().zko_fe_sum(fid, co._p(co.sumpoly_reduce(fid, tabs)), 1 << n, co._p(claimed))
        coeffs, ch, fin = co.product_prove(fid, tabs, claimed, co.Transcript())
        c2 = np.zeros((n, 3, 4), dtype=np.uint64); ch2 = np.zeros((n, 4), dtype=np.uint64); fin2 = np.zeros((2, 4), dtype=np.uint64)
        tr = Transcript()
        host = np.ascontiguousarray(np.stack([f, g]))
        ctx.check(ctx.lib.zk_prove_product_host(ctx.h, _ptr(host), 1, 2, 1 << n, _ptr(claimed), tr.h, _ptr(c2), _ptr(ch2), _ptr(fin2), 0))
        assert np.array_equal(c2, coeffs) and np.array_equal(ch2, ch) and np.array_equal(fin2, fin[0])


def test_product_sumcheck_golden_and_reference_round_trip(zk, co, ctx_for, golden):
    from zk_cryptography_research_implementations_b200 import sumcheck_protocol as scp
    from zk_cryptography_research_implementations_b200.transcripts import Transcript
    for e in golden["generated"]["product"]:
        fid = FIELDS[e["field"]]
        tabs = np.stack([np.stack([zk.fe_from_ints(fid, t) for t in prod]) for prod in e["tables"]])
        proof = scp.prove(sumpoly_handle(zk, ctx_for(fid), tabs), zk.fe_from_ints(fid, [e["claimed_sum"]])[0], Transcript())
        assert [zk.fe_to_ints(fid, p.coefficients) for p in proof.round_univariate_polynomials] == e["coeffs"]
        assert zk.fe_to_ints(fid, proof.random_challenges) == e["challenges"]
        assert zk.fe_to_ints(fid, proof.final_values) == [v for prod in e["final_tables"] for v in prod]
    # sumcheck_gkr_protocol.rs:163-212 -- the reference's own tests, written the way they are there
    from zk_cryptography_research_implementations_b200.polynomials import MultilinearPolynomial as MLE, ProductPolynomial, SumPolynomial
    ctx = ctx_for(0)
    Fq = lambda v: zk.fe_from_ints(0, [v])[0]
    def build():
        poly1a = MLE.new(ctx, np.stack([Fq(0), Fq(0), Fq(0), Fq(2)]))
        poly2a = MLE.new(ctx, np.stack([Fq(0), Fq(0), Fq(0), Fq(3)]))
        poly1b = MLE.new(ctx, np.stack([Fq(0), Fq(0), Fq(0), Fq(2)]))
        poly2b = MLE.new(ctx, np.stack([Fq(0), Fq(0), Fq(0), Fq(3)]))
        return SumPolynomial([ProductPolynomial([poly1a, poly2a]), ProductPolynomial([poly1b, poly2b])])
    assert zk.fe_to_ints(0, scp.generate_round_univariate(build())) == [0, 12, 48]
    result = scp.prove(build(), Fq(12), Transcript())
    coeffs = np.stack([p.coefficients for p in result.round_univariate_polynomials])
    ok, chal, _ = co.product_verify(0, Fq(12), coeffs, co.Transcript())
    assert ok and np.array_equal(chal, result.random_challenges)


def test_reference_polynomial_kats_through_the_mirror_api(zk, co, ctx_for, golden):
    from zk_cryptography_research_implementations_b200.polynomials import (DenseUnivariatePolynomial, MultilinearPolynomial as MLE,
                                                                          ProductPolynomial, SumPolynomial)
    k = golden["reference_kats"]
    ctx = ctx_for(0)
    F = lambda vals: zk.fe_from_ints(0, vals)
    for e in k["partial_evaluate"]:
        got = MLE.partial_evaluate(MLE.new(ctx, F(e["table"])), e["var"], F([e["r"]])[0])
        assert zk.fe_to_ints(0, got.evaluated_values) == e["out"], e["src"]
    for e in k["evaluate"]:
        assert zk.fe_to_ints(0, MLE.new(ctx, F(e["table"])).evaluate(F(e["values"]))) == [e["out"]]
    with pytest.raises(zk.ReferencePanic, match=k["new_panics"]["message"]):
        MLE.new(ctx, F(k["new_panics"]["table"]))
    for e in k["tensor"]:
        f = MLE.polynomial_tensor_add if e["op"] == "add" else MLE.polynomial_tensor_mul
        assert zk.fe_to_ints(0, f(MLE.new(ctx, F(e["wb"])), MLE.new(ctx, F(e["wc"]))).evaluated_values) == e["out"]
    pp = k["product_polynomial"]
    prod = ProductPolynomial([MLE.new(ctx, F(t)) for t in pp["polys"]])
    assert zk.fe_to_ints(0, prod.evaluate(F(pp["evaluate_at"]))) == [pp["evaluate_out"]]
    assert [zk.fe_to_ints(0, x.evaluated_values) for x in prod.partial_evaluate(0, F([pp["fold_r"]])[0])] == pp["fold_out"]
    assert zk.fe_to_ints(0, prod.multiply_polynomials_element_wise().evaluated_values) == pp["elementwise"]
    assert prod.degree() == pp["degree"]
    with pytest.raises(zk.ReferencePanic, match=k["product_panics"]["message"]):
        ProductPolynomial([MLE.new(ctx, F(t)) for t in k["product_panics"]["polys"]])
    sp = k["sum_polynomial"]
    sump = SumPolynomial([ProductPolynomial([MLE.new(ctx, F(t)) for t in prod_]) for prod_ in sp["products"]])
    assert zk.fe_to_ints(0, sump.evaluate(F(sp["evaluate_at"]))) == [sp["evaluate_out"]]
    folded = sump.partial_evaluate(0, F([sp["fold_r"]])[0])
    assert [[zk.fe_to_ints(0, x.evaluated_values) for x in pr.polynomials] for pr in folded.product_polynomials] == sp["fold_out"]
    assert zk.fe_to_ints(0, sump.add_polynomials_element_wise().evaluated_values) == sp["elementwise"]
    assert sump.degree() == sp["degree"] and sump.number_of_variables() == sp["number_of_variables"]
    ue = k["univariate_evaluate"]
    assert zk.fe_to_ints(0, DenseUnivariatePolynomial(0, F(ue["coeffs"])).evaluate(F([ue["x"]])[0])) == [ue["out"]]


# ------------------------------------------------------------------------------------------ size-independent properties
@pytest.mark.parametrize("fid,P,D,n", [(0, 1, 2, 22), (0, 2, 2, 20), (2, 1, 2, 21)])
def test_large_product_sumcheck_properties(zk, co, ctx_for, fid, P, D, n):
    """At sizes the oracle cannot prove in seconds: (1) the reference verifier accepts the proof,
    (2) its final claim equals sum_p prod_d of the fully folded tables, (3) which equal the tables'
    MLE evaluations at the challenges (computed by the separate evaluate kernels), (4) the
    claim-derived s(1) and the directly computed s(1) give identical proofs."""
    import ctypes as C
    from zk_cryptography_research_implementations_b200.core import _ptr
    from zk_cryptography_research_implementations_b200.transcripts import Transcript
    ctx = ctx_for(fid)
    N = 1 << n

    def run(flags):
        tabs = [ctx.generate(0xB200, i, N) for i in range(P * D)]
        # claimed sum from the round-0 kernel: s(0) + s(1)
        arr = (C.c_void_p * (P * D))(*[t.release() for t in tabs])
        h = C.c_void_p()
        ctx.check(ctx.lib.zk_sumpoly_create(ctx.h, arr, P, D, C.byref(h)))
        ev = np.zeros((D + 1, 4), dtype=np.uint64)
        ctx.check(ctx.lib.zk_sumcheck_round_evals(ctx.h, h, _ptr(ev)))
        claimed = zk.fe_binop("add", fid, ev[0], ev[1])
        coeffs = np.zeros((n, D + 1, 4), dtype=np.uint64); ch = np.zeros((n, 4), dtype=np.uint64); fin = np.zeros((P * D, 4), dtype=np.uint64)
        tr = Transcript()   # keep it alive across the call
        ctx.check(ctx.lib.zk_prove_product(ctx.h, h, _ptr(claimed), tr.h, _ptr(coeffs), _ptr(ch), _ptr(fin), flags))
        ctx.lib.zk_sumpoly_free(ctx.h, h)
        return claimed, coeffs, ch, fin

    claimed, coeffs, ch, fin = run(0)
    claimed1, coeffs1, ch1, fin1 = run(1)
    assert np.array_equal(coeffs, coeffs1) and np.array_equal(ch, ch1) and np.array_equal(fin, fin1)
    ok, ch_v, last = co.product_verify(fid, claimed, coeffs, co.Transcript())
    assert ok and np.array_equal(ch_v, ch)
    total = np.zeros(4, dtype=np.uint64)
    for p in range(P):
        prod = fin[p * D]
        for d in range(1, D):
            prod = co.fe_op("mul", fid, prod, fin[p * D + d])
        total = co.fe_op("add", fid, total, prod)
    assert np.array_equal(total, last)
    from zk_cryptography_research_implementations_b200.polynomials import MultilinearPolynomial as MLE
    for i in range(P * D):
        t = MLE(ctx, ctx.generate(0xB200, i, N))
        assert np.array_equal(t.evaluate(ch), fin[i])


@pytest.mark.parametrize("fid,n", [(2, 22), (0, 20)])
def test_large_basic_sumcheck_properties(zk, co, ctx_for, fid, n):
    """plain sumcheck at 2^20 / 2^22: round sums telescope (s0 + s1 == previous p(r)), the final value is
    the MLE evaluation at the challenges, and the challenges replay from the transcript."""
    import ctypes as C
    from zk_cryptography_research_implementations_b200.core import _ptr
    from zk_cryptography_research_implementations_b200.polynomials import MultilinearPolynomial as MLE
    ctx = ctx_for(fid)
    N = 1 << n
    t = ctx.generate(7, 0, N)
    claimed = np.zeros(4, dtype=np.uint64); rp = np.zeros((n, 2, 4), dtype=np.uint64); ch = np.zeros((n, 4), dtype=np.uint64); fin = np.zeros(4, dtype=np.uint64)
    ctx.check(ctx.lib.zk_prove_basic_device(ctx.h, t.h, _ptr(claimed), _ptr(rp), _ptr(ch), _ptr(fin), 0))
    # replay with the oracle's transcript, absorbing the bytes the GPU converter produced
    orig = MLE(ctx, ctx.generate(7, 0, N))
    tr = co.Transcript()
    tr.append(orig.convert_to_bytes())
    tr.append(co.mle_to_bytes(fid, claimed.reshape(1, 4)))
    claim = claimed
    for k in range(n):
        assert np.array_equal(co.fe_op("add", fid, rp[k, 0], rp[k, 1]), claim), k
        tr.append(co.mle_to_bytes(fid, rp[k]))
        r = tr.random_challenge_as_field_element(fid)
        assert np.array_equal(r, ch[k]), k
        claim = co.mle_evaluate(fid, rp[k], r.reshape(1, 4))
    assert np.array_equal(claim, fin)
    assert np.array_equal(orig.evaluate(ch), fin)
    # spot-check convert_to_bytes against the oracle on a slice
    sl = orig.evaluated_values[:1024]
    assert orig.convert_to_bytes()[:32 * 1024] == co.mle_to_bytes(fid, sl)


def test_basic_sharded_entry_point_world_1(zk, co, ctx_for):
    """zk_prove_basic_sharded with a single rank == the reference proof when the caller absorbs the table"""
    from zk_cryptography_research_implementations_b200.core import _ptr
    from zk_cryptography_research_implementations_b200.transcripts import Transcript
    fid = 2
    ctx = ctx_for(fid)
    for n in (0, 1, 5, 12):
        T = rand_table(co, fid, 1 << n, 900 + n)
        claimed_w, rp_w, ch_w, fin_w = co.basic_prove(fid, T)
        t = ctx.upload(T)
        tr = Transcript()
        tr.append(co.mle_to_bytes(fid, T))
        claimed = np.zeros(4, dtype=np.uint64); rp = np.zeros((max(n, 1), 2, 4), dtype=np.uint64)
        ch = np.zeros((max(n, 1), 4), dtype=np.uint64); fin = np.zeros(4, dtype=np.uint64)
        ctx.check(ctx.lib.zk_prove_basic_sharded(ctx.h, t.h, tr.h, _ptr(claimed), _ptr(rp), _ptr(ch), _ptr(fin), 0, 1))
        assert np.array_equal(claimed, claimed_w) and np.array_equal(rp[:n], rp_w) and np.array_equal(ch[:n], ch_w) and np.array_equal(fin, fin_w)


def test_degenerate_sizes(zk, co, ctx_for):
    """one-entry tables: zero rounds (prove returns an empty proof and leaves the transcript with only the claim absorbed)"""
    from zk_cryptography_research_implementations_b200 import sumcheck_protocol as scp
    from zk_cryptography_research_implementations_b200.transcripts import Transcript
    fid = 0
    ctx = ctx_for(fid)
    tabs = np.stack([np.stack([rand_table(co, fid, 1, 70 + 2 * p + d, edge=False) for d in range(2)]) for p in range(2)])
    claimed = np.zeros(4, dtype=np.uint64)
    co.lib().zko_fe_sum(fid, co._p(co.sumpoly_reduce(fid, tabs)), 1, co._p(claimed))
    tr_o, tr_g = co.Transcript(), Transcript()
    coeffs, ch, fin = co.product_prove(fid, tabs, claimed, tr_o)
    proof = scp.prove(sumpoly_handle(zk, ctx, tabs), claimed, tr_g)
    assert coeffs.shape[0] == 0 and len(proof.round_univariate_polynomials) == 0 and proof.random_challenges.shape[0] == 0
    assert np.array_equal(proof.final_values.reshape(2, 2, 4), fin)
    assert tr_o.sample_random_challenge() == tr_g.sample_random_challenge()
    # evaluate with no values returns entry 0; partial_evaluate of a one-entry table is the reference's panic
    from zk_cryptography_research_implementations_b200.polynomials import MultilinearPolynomial as MLE
    one = MLE.new(ctx, tabs[0, 0])
    assert np.array_equal(one.evaluate(np.zeros((0, 4), dtype=np.uint64)), tabs[0, 0, 0])
    with pytest.raises(zk.ReferencePanic):
        MLE.partial_evaluate(one, 0, tabs[0, 1, 0])
