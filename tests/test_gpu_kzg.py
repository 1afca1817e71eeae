"""Multilinear KZG on the GPU (csrc/kzg.cu through the C-ABI) against the CPU oracle (oracle/zkoracle_kzg.c) and the
independent Python model with the pairing (oracle/pykzg.py): the reference's own tests (multilinear_kzg.rs:218-303),
point-for-point equality of the setup, the commitment and every opening proof, the edge cases of the group law inside the
multi-scalar multiplication (zero scalars, repeated points, P + (-P), infinity in the setup), every window width, and
-- at sizes the oracle cannot open in seconds -- the verification equation itself checked with the trapdoor."""
import random

import numpy as np
import pytest

import pykzg as pk

pytestmark = pytest.mark.gpu
FR = 2
R = pk.R


@pytest.fixture
def kzg(zk):
    from zk_cryptography_research_implementations_b200 import multilinear_kzg
    return multilinear_kzg


def _fe(co, ints):
    return co.from_ints(FR, [x % R for x in ints])


REFERENCE_CASES = [  # multilinear_kzg/src/multilinear_kzg.rs:223-303
    ([5, 2, 3], [0, 4, 0, 4, 0, 4, 3, 7], [6, 4, 0]),
    ([2, 3, 4], [0, 7, 0, 5, 0, 7, 4, 9], [5, 9, 6]),
    ([12, 9, 28, 40], [0, 0, 0, 2, 0, 0, 10, 12, 0, -12, 4, -6, 0, -12, 14, 4], [54, 90, 76, 160]),
]


@pytest.mark.parametrize("case", range(len(REFERENCE_CASES)))
def test_reference_kzg_tests(co, ctx_for, kzg, case):
    taus, vals, opening = REFERENCE_CASES[case]
    ctx = ctx_for(FR)
    t, v, o = _fe(co, taus), _fe(co, vals), _fe(co, opening)
    setup = kzg.TrustedSetup.initialize_setup(ctx, t)
    assert (setup.g1_powers_of_tau == co.kzg_setup_g1(t)).all()
    c = kzg.MultilinearKZG.commit_to_polynomial(v, setup)
    proof = kzg.MultilinearKZG.open_and_prove(v, setup, o)
    oc = co.kzg_commit(v, setup.g1_powers_of_tau)
    oev, oproofs = co.kzg_open(v, setup.g1_powers_of_tau, o)
    assert (c == oc).all() and (proof.evaluation == oev).all() and (proof.proofs == oproofs).all()
    # the reference's assertion: verify(...) == true, with the pairing
    psetup = pk.TrustedSetup.initialize([x % R for x in taus])
    assert pk.verify(psetup, co.g1_to_ints(c)[0], [x % R for x in opening], co.to_ints(FR, proof.evaluation)[0], co.g1_to_ints(proof.proofs))


def test_golden_vectors(co, ctx_for, kzg, kzg_golden):
    """the committed golden commitments and openings (tests/golden/kzg_golden.json, Python big-int model, pairing-verified)"""
    from conftest import golden_point
    ctx = ctx_for(FR)
    for e in kzg_golden["generated"]:
        taus, vals, opening = ([int(x) for x in e[k]] for k in ("taus", "values", "opening"))
        t, v, o = _fe(co, taus), _fe(co, vals), _fe(co, opening)
        setup = kzg.TrustedSetup.initialize_setup(ctx, t)
        pts = co.g1_to_ints(setup.g1_powers_of_tau)
        assert pts[0] == golden_point(e["g1_powers_of_tau_first"]) and pts[-1] == golden_point(e["g1_powers_of_tau_last"]), e["src"]
        assert co.g1_to_ints(kzg.MultilinearKZG.commit_to_polynomial(v, setup))[0] == golden_point(e["commitment"]), e["src"]
        proof = kzg.MultilinearKZG.open_and_prove(v, setup, o)
        assert co.to_ints(FR, proof.evaluation)[0] == int(e["evaluation"])
        assert co.g1_to_ints(proof.proofs) == [golden_point(p) for p in e["proofs"]], e["src"]
        assert kzg.MultilinearKZG.verify(setup, kzg.MultilinearKZG.commit_to_polynomial(v, setup), o, proof)


@pytest.mark.parametrize("n", [1, 2, 5, 9, 12])
def test_setup_commit_open_match_oracle(co, ctx_for, kzg, zk, n):
    ctx = ctx_for(FR)
    rnd = random.Random(100 + n)
    taus = _fe(co, [rnd.randrange(R) for _ in range(n)])
    opening = _fe(co, [rnd.randrange(R) for _ in range(n)])
    vals = co.table_generate(FR, 0xB200, 7, 1 << n)
    co.set_threads(8)
    try:
        osetup = co.kzg_setup_g1(taus)
        setup = kzg.TrustedSetup.initialize_setup(ctx, taus)
        assert (setup.g1_powers_of_tau == osetup).all()
        # the folded levels are the Lagrange bases of the remaining variables
        for k in range(1, n + 1):
            lvl = setup.level(k)
            if k < n:
                assert (lvl == co.kzg_setup_g1(taus[k:])).all()
            else:
                assert (lvl[0] == co.g1_generator()).all()     # the basis sums to one
        table = ctx.upload(vals)
        c = kzg.MultilinearKZG.commit_to_polynomial(table, setup)
        assert (c == co.kzg_commit(vals, osetup)).all()
        assert (kzg.MultilinearKZG.commit_to_polynomial(vals, setup) == c).all()
        proof = kzg.MultilinearKZG.open_and_prove(table, setup, opening)
        oev, oproofs = co.kzg_open(vals, osetup, opening)
        assert (proof.evaluation == oev).all() and (proof.proofs == oproofs).all()
        assert (table.download() == vals).all()                # the resident polynomial is not consumed
        assert co.kzg_verify_trapdoor(taus, c, opening, proof.evaluation, proof.proofs)
        # an uploaded setup behaves like a generated one
        setup2 = kzg.TrustedSetup.from_g1_powers_of_tau(ctx, osetup)
        proof2 = kzg.MultilinearKZG.open_and_prove(vals, setup2, opening)
        assert (proof2.proofs == oproofs).all() and (proof2.evaluation == oev).all()
    finally:
        co.set_threads(1)


def test_msm_group_law_edge_cases(co, ctx_for, kzg):
    """zero scalars, the same point many times (doubling inside a bucket), P and -P in one bucket, infinity among the points,
    scalars p-1 / 1 / 2^c boundaries (digit carries)"""
    ctx = ctx_for(FR)
    g = co.g1_generator()
    rnd = random.Random(5)
    pts_int = [pk.g1_mul(pk.G1_GEN, k) for k in (1, 2, 3, 7, R - 1, R - 2)] + [None]
    pts = co.g1_from_ints(pts_int)
    cases = []
    cases.append(([5] * 40, [0] * 40))                                       # one point, one bucket: doublings
    cases.append(([3, 3], [0, 4]))                                           # G and -G with equal digits: cancels inside a bucket
    cases.append(([0] * 7, list(range(7))))                                  # all scalars zero -> infinity
    cases.append(([1, R - 1, 1 << 15, (1 << 15) + 1, (1 << 16) - 1, 1 << 16, (1 << 255) % R, R - 2], [0, 1, 2, 3, 0, 1, 2, 3]))
    cases.append(([rnd.randrange(R) for _ in range(64)], [rnd.randrange(7) for _ in range(64)]))   # with infinity points
    cases.append(([9], [6]))                                                 # a lone infinity
    for scalars, idx in cases:
        want = None
        for s, i in zip(scalars, idx):
            want = pk.g1_add(want, pk.g1_mul(pts_int[i], s % R))
        got = kzg.g1_msm(ctx, _fe(co, scalars), pts[idx])
        assert co.g1_to_ints(got)[0] == want
    assert not kzg.g1_msm(ctx, np.zeros((0, 4), dtype=np.uint64), np.zeros((0, 12), dtype=np.uint64)).any()
    bad = pts[:1].copy()
    bad[0, 0] ^= 1
    with pytest.raises(Exception, match="not on the curve"):
        kzg.g1_msm(ctx, _fe(co, [1]), bad)


@pytest.mark.parametrize("c", [2, 3, 5, 8, 11, 15, 16])
def test_every_window_width_gives_the_same_sum(co, ctx_for, kzg, monkeypatch, c):
    ctx = ctx_for(FR)
    n = 300
    rnd = random.Random(c)
    ks = [rnd.randrange(R) for _ in range(8)]
    base = co.g1_from_ints([pk.g1_mul(pk.G1_GEN, k) for k in ks])
    idx = [rnd.randrange(8) for _ in range(n)]
    scalars = [rnd.randrange(R) for _ in range(n - 4)] + [0, 1, R - 1, (1 << 255) % R]
    want = pk.g1_mul(pk.G1_GEN, sum(s * ks[i] for s, i in zip(scalars, idx)) % R)
    monkeypatch.setenv("ZKB200_MSM_WINDOW", str(c))
    got = kzg.g1_msm(ctx, _fe(co, scalars), base[idx])
    assert co.g1_to_ints(got)[0] == want


def test_small_valued_and_constant_polynomials(co, ctx_for, kzg):
    """the shapes the reference's tests use: small integers (a few huge buckets) and a constant table"""
    ctx = ctx_for(FR)
    n = 10
    rnd = random.Random(77)
    taus = _fe(co, [rnd.randrange(R) for _ in range(n)])
    opening = _fe(co, [rnd.randrange(50) for _ in range(n)])
    setup = kzg.TrustedSetup.initialize_setup(ctx, taus)
    co.set_threads(8)
    try:
        for vals in ([3] * (1 << n), [rnd.randrange(10) for _ in range(1 << n)], [0] * (1 << n)):
            v = _fe(co, vals)
            c = kzg.MultilinearKZG.commit_to_polynomial(v, setup)
            assert (c == co.kzg_commit(v, setup.g1_powers_of_tau)).all()
            proof = kzg.MultilinearKZG.open_and_prove(v, setup, opening)
            assert co.kzg_verify_trapdoor(taus, c, opening, proof.evaluation, proof.proofs)
            assert co.to_ints(FR, proof.evaluation)[0] == pk.mle_evaluate([x % R for x in vals], co.to_ints(FR, opening))
    finally:
        co.set_threads(1)


def test_one_giant_bucket(co, ctx_for, kzg):
    """2^16 equal scalars: every point lands in one bucket of the lowest window (all three levels of the bucket sum in
    use); the Lagrange basis sums to one, so the commitment of the constant 3 is 3 G.  Two values = two giant buckets:
    the polynomial is 5 - 2 x_last, checked through its opening."""
    ctx = ctx_for(FR)
    n = 16
    rnd = random.Random(3)
    taus = _fe(co, [rnd.randrange(R) for _ in range(n)])
    setup = kzg.TrustedSetup.initialize_setup(ctx, taus)
    three = _fe(co, [3])[0]
    c = kzg.MultilinearKZG.commit_to_polynomial(np.tile(three, (1 << n, 1)), setup)
    assert co.g1_to_ints(c)[0] == pk.g1_mul(pk.G1_GEN, 3)
    vals = np.tile(three, (1 << n, 1))
    vals[::2] = _fe(co, [5])[0]
    opening = _fe(co, [rnd.randrange(R) for _ in range(n)])
    c = kzg.MultilinearKZG.commit_to_polynomial(vals, setup)
    tau_last = co.to_ints(FR, taus)[-1]
    assert co.g1_to_ints(c)[0] == pk.g1_mul(pk.G1_GEN, (5 - 2 * tau_last) % R)
    proof = kzg.MultilinearKZG.open_and_prove(vals, setup, opening)
    assert co.to_ints(FR, proof.evaluation)[0] == (5 - 2 * co.to_ints(FR, opening)[-1]) % R
    assert co.kzg_verify_trapdoor(taus, c, opening, proof.evaluation, proof.proofs)


def test_degenerate_taus_put_infinity_into_the_setup(co, ctx_for, kzg):
    ctx = ctx_for(FR)
    taus = _fe(co, [0, 1, 5])
    setup = kzg.TrustedSetup.initialize_setup(ctx, taus)
    osetup = co.kzg_setup_g1(taus)
    assert (setup.g1_powers_of_tau == osetup).all() and not osetup[0].any()
    vals = _fe(co, [1, 2, 3, 4, 5, 6, 7, 8])
    opening = _fe(co, [9, 8, 7])
    c = kzg.MultilinearKZG.commit_to_polynomial(vals, setup)
    proof = kzg.MultilinearKZG.open_and_prove(vals, setup, opening)
    oev, oproofs = co.kzg_open(vals, osetup, opening)
    assert (c == co.kzg_commit(vals, osetup)).all() and (proof.proofs == oproofs).all() and (proof.evaluation == oev).all()


def test_reference_asserts(co, ctx_for, kzg, zk):
    ctx = ctx_for(FR)
    setup = kzg.TrustedSetup.initialize_setup(ctx, _fe(co, [5, 2, 3]))
    with pytest.raises(zk.ReferencePanic, match="requires at least one variable"):
        kzg.TrustedSetup.initialize_setup(ctx, np.zeros((0, 4), dtype=np.uint64))
    with pytest.raises(zk.ReferencePanic, match="Polynomial evaluation must match g1 length"):
        kzg.MultilinearKZG.commit_to_polynomial(_fe(co, [1, 2, 3, 4]), setup)
    with pytest.raises(zk.ReferencePanic, match="number of polynomial variables must match length of opening values"):
        kzg.MultilinearKZG.open_and_prove(_fe(co, list(range(8))), setup, _fe(co, [1, 2]))
    with pytest.raises(zk.ReferencePanic, match="Opening values must match number of variables from trusted setup"):
        kzg.MultilinearKZG.open_and_prove(_fe(co, list(range(4))), setup, _fe(co, [1, 2]))
    with pytest.raises(zk.ZkError, match="BLS12-381 Fr"):
        kzg.TrustedSetup.initialize_setup(ctx_for(0), _fe(co, [5, 2, 3]))
    bad = setup.g1_powers_of_tau.copy()
    bad[3, 7] ^= 2
    with pytest.raises(zk.ZkError, match="not on the curve"):
        kzg.TrustedSetup.from_g1_powers_of_tau(ctx, bad)


@pytest.mark.parametrize("n", [16, 20])
def test_large_commit_and_opening_verify(co, ctx_for, kzg, n):
    """past what the oracle can open in seconds: the commitment against the oracle at 2^16, and the verification equation
    C - v G == sum (tau_i - r_i) Q_i with the trapdoor at both sizes (the evaluation against zk's own evaluate)"""
    ctx = ctx_for(FR)
    rnd = random.Random(n)
    taus = _fe(co, [rnd.randrange(R) for _ in range(n)])
    opening = _fe(co, [rnd.randrange(R) for _ in range(n)])
    setup = kzg.TrustedSetup.initialize_setup(ctx, taus)
    table = ctx.generate(0xB200, 3, 1 << n)
    c = kzg.MultilinearKZG.commit_to_polynomial(table, setup)
    proof = kzg.MultilinearKZG.open_and_prove(table, setup, opening)
    assert co.g1_is_on_curve(c) and all(co.g1_is_on_curve(p) for p in proof.proofs)
    assert co.kzg_verify_trapdoor(taus, c, opening, proof.evaluation, proof.proofs)
    from zk_cryptography_research_implementations_b200.polynomials import MultilinearPolynomial
    assert (MultilinearPolynomial(ctx, table).evaluate(opening) == proof.evaluation).all()
    wrong = proof.proofs.copy()
    wrong[n // 2] = co.g1_add(wrong[n // 2], co.g1_generator())
    assert not co.kzg_verify_trapdoor(taus, c, opening, proof.evaluation, wrong)
    if n == 16:
        co.set_threads(8)
        try:
            assert (c == co.kzg_commit(table.download(), setup.g1_powers_of_tau)).all()
            # spot-check the setup against the oracle's scalar multiplications
            basis = pk.lagrange_basis(co.to_ints(FR, taus[-6:]))
            sub = kzg.TrustedSetup.initialize_setup(ctx, taus[-6:])
            assert co.g1_to_ints(sub.g1_powers_of_tau) == [pk.g1_mul(pk.G1_GEN, e) for e in basis]
            assert (setup.level(n - 6) == sub.g1_powers_of_tau).all()
        finally:
            co.set_threads(1)
