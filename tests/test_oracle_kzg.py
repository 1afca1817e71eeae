"""The multilinear-KZG oracle (oracle/zkoracle_kzg.c) against the reference's own tests and the independent Python model
(oracle/pykzg.py: affine big-int curve arithmetic and the BLS12-381 pairing).  CPU only."""
import random

import numpy as np
import pytest

import pykzg as pk

R = pk.R
FR = 2


def test_curve_model_self_check():
    """generators on their curves and of order r, the pairing bilinear and non-degenerate"""
    pk.self_check()


def test_lagrange_basis_known_answers():
    """multilinear_kzg/src/trusted_setup.rs:101-126"""
    assert pk.lagrange_basis([5, 2, 3]) == [x % R for x in [-8, 12, 16, -24, 10, -15, -20, 30]]
    assert pk.lagrange_basis([5, 2]) == [x % R for x in [4, -8, -5, 10]]
    with pytest.raises(AssertionError, match="requires at least one variable"):
        pk.lagrange_basis([])


def test_group_law_against_python_ints(co):
    g = co.g1_generator()
    assert co.g1_to_ints(g)[0] == pk.G1_GEN and co.g1_is_on_curve(g)
    rnd = random.Random(7)
    for k in [0, 1, 2, 3, R - 1, R, R + 1] + [rnd.randrange(1 << 256) for _ in range(4)]:
        assert co.g1_to_ints(co.g1_mul(g, k))[0] == pk.g1_mul(pk.G1_GEN, k)
    a, b = co.g1_mul(g, 5), co.g1_mul(g, 7)
    inf = np.zeros(12, dtype=np.uint64)
    assert co.g1_to_ints(co.g1_add(a, b)) == co.g1_to_ints(co.g1_mul(g, 12))          # chord
    assert co.g1_to_ints(co.g1_add(a, a)) == co.g1_to_ints(co.g1_mul(g, 10))          # tangent
    assert co.g1_to_ints(co.g1_add(a, co.g1_mul(g, R - 5))) == [None]                 # P + (-P)
    assert co.g1_to_ints(co.g1_add(a, inf)) == co.g1_to_ints(a) == co.g1_to_ints(co.g1_add(inf, a))
    bad = a.copy()
    bad[0] ^= 1
    assert not co.g1_is_on_curve(bad)
    assert co.g1_to_ints(co.g1_from_ints([pk.G1_GEN, None])) == [pk.G1_GEN, None]


REFERENCE_CASES = [  # multilinear_kzg/src/multilinear_kzg.rs:223-303 (taus, evaluations, opening values)
    ([5, 2, 3], [0, 4, 0, 4, 0, 4, 3, 7], [6, 4, 0]),
    ([2, 3, 4], [0, 7, 0, 5, 0, 7, 4, 9], [5, 9, 6]),
    ([12, 9, 28, 40], [0, 0, 0, 2, 0, 0, 10, 12, 0, -12, 4, -6, 0, -12, 14, 4], [54, 90, 76, 160]),
]


def _both(co, taus, vals, opening, pairing):
    taus, vals, opening = ([x % R for x in v] for v in (taus, vals, opening))
    t, v, o = (co.from_ints(FR, x) for x in (taus, vals, opening))
    setup = co.kzg_setup_g1(t)
    psetup = pk.TrustedSetup.initialize(taus)
    assert co.g1_to_ints(setup) == psetup.g1_powers_of_tau
    c = co.kzg_commit(v, setup)
    assert co.g1_to_ints(c)[0] == pk.commit(vals, psetup)
    ev, proofs = co.kzg_open(v, setup, o)
    pev, pproofs = pk.open_and_prove(vals, psetup, opening)
    assert co.to_ints(FR, ev)[0] == pev and co.g1_to_ints(proofs) == pproofs
    assert co.kzg_verify_trapdoor(t, c, o, ev, proofs)
    assert pk.verify_with_trapdoor(taus, pk.commit(vals, psetup), opening, pev, pproofs)
    wrong = co.from_ints(FR, [(pev + 1) % R])[0]
    assert not co.kzg_verify_trapdoor(t, c, o, wrong, proofs)
    if pairing:   # the reference's verify(): e(C - vG, G2) == prod e(Q_i, tau_i G2 - r_i G2)
        assert pk.verify(psetup, co.g1_to_ints(c)[0], opening, pev, co.g1_to_ints(proofs))
        assert not pk.verify(psetup, co.g1_to_ints(c)[0], opening, (pev + 1) % R, co.g1_to_ints(proofs))


@pytest.mark.parametrize("case", range(len(REFERENCE_CASES)))
def test_reference_kzg_tests_verify(co, case):
    _both(co, *REFERENCE_CASES[case], pairing=(case != 1))


def test_random_polynomial_against_python_model(co):
    rnd = random.Random(11)
    n = 4
    _both(co, [rnd.randrange(R) for _ in range(n)], [rnd.randrange(R) for _ in range(1 << n)],
          [rnd.randrange(R) for _ in range(n)], pairing=False)


def test_golden_vectors(co, kzg_golden):
    """tests/golden/kzg_golden.json: the reference's Lagrange-basis known answers, and the commitments / openings of its three
    KZG tests as the Python model computed (and pairing-verified) them, against the C oracle and the Python model of today"""
    from conftest import golden_point
    assert golden_point(kzg_golden["curve"]["g1_generator"]) == pk.G1_GEN and int(kzg_golden["curve"]["r"], 16) == R
    for k in kzg_golden["reference_kats"]["lagrange_basis"]:
        assert pk.lagrange_basis(k["taus"]) == [x % R for x in k["out"]], k["src"]
    for e in kzg_golden["generated"]:
        taus, vals, opening = ([int(x) for x in e[k]] for k in ("taus", "values", "opening"))
        t, v, o = (co.from_ints(FR, x) for x in (taus, vals, opening))
        setup = co.kzg_setup_g1(t)
        got_setup = co.g1_to_ints(setup)
        assert got_setup[0] == golden_point(e["g1_powers_of_tau_first"]) and got_setup[-1] == golden_point(e["g1_powers_of_tau_last"]), e["src"]
        assert co.g1_to_ints(co.kzg_commit(v, setup))[0] == golden_point(e["commitment"]), e["src"]
        ev, proofs = co.kzg_open(v, setup, o)
        assert co.to_ints(FR, ev)[0] == int(e["evaluation"]) and co.g1_to_ints(proofs) == [golden_point(p) for p in e["proofs"]], e["src"]
        if len(taus) <= 3:
            psetup = pk.TrustedSetup.initialize(taus)
            assert pk.commit(vals, psetup) == golden_point(e["commitment"])


def test_threads_do_not_change_results(co):
    rnd = random.Random(12)
    n = 6
    taus = co.from_ints(FR, [rnd.randrange(R) for _ in range(n)])
    vals = co.table_generate(FR, 3, 0, 1 << n)
    opening = co.from_ints(FR, [rnd.randrange(R) for _ in range(n)])
    co.set_threads(1)
    s1 = co.kzg_setup_g1(taus)
    c1 = co.kzg_commit(vals, s1)
    e1, p1 = co.kzg_open(vals, s1, opening)
    if co.openmp_enabled():
        co.set_threads(4)
        try:
            assert (co.kzg_setup_g1(taus) == s1).all() and (co.kzg_commit(vals, s1) == c1).all()
            e4, p4 = co.kzg_open(vals, s1, opening)
            assert (e4 == e1).all() and (p4 == p1).all()
        finally:
            co.set_threads(1)
    assert co.kzg_verify_trapdoor(taus, c1, opening, e1, p1)


def test_reference_asserts(co):
    """multilinear_kzg.rs:29-33, :56-65"""
    taus = co.from_ints(FR, [5, 2, 3])
    setup = co.kzg_setup_g1(taus)
    with pytest.raises(AssertionError, match="Polynomial evaluation must match g1 length"):
        co.kzg_commit(co.from_ints(FR, [1, 2, 3, 4]), setup)
    with pytest.raises(AssertionError, match="number of polynomial variables must match length of opening values"):
        co.kzg_open(co.from_ints(FR, list(range(8))), setup, co.from_ints(FR, [1, 2]))
    with pytest.raises(AssertionError, match="Opening values must match number of variables from trusted setup"):
        co.kzg_open(co.from_ints(FR, list(range(4))), setup, co.from_ints(FR, [1, 2]))
