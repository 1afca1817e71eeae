"""The product's host-only KZG verifier (csrc/kzg_verify.cu: G2 setup + pairing check through the C-ABI, no GPU) against
proofs made by the oracle, and against the independent Python pairing model.  CPU only."""
import ctypes as C
import random

import numpy as np
import pytest

import pykzg as pk

FR = 2
R = pk.R


@pytest.fixture(scope="module")
def lib(zk):
    from zk_cryptography_research_implementations_b200 import _lib
    return _lib.load()


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _g2_to_ints(co, a):
    """(n, 24) Montgomery limbs -> [((x0, x1), (y0, y1)) | None]"""
    out = []
    for row in np.ascontiguousarray(a).reshape(-1, 24):
        if not row.any():
            out.append(None)
            continue
        c = []
        tmp = np.zeros(6, dtype=np.uint64)
        for k in range(4):
            co.lib().zko_fq_to_canonical(_p(np.ascontiguousarray(row[6 * k:6 * k + 6])), _p(tmp))
            c.append(sum(int(tmp[j]) << (64 * j) for j in range(6)))
        out.append(((c[0], c[1]), (c[2], c[3])))
    return out


def verify(lib, g2, commitment, opening, evaluation, proofs):
    ok = C.c_int(-1)
    g2, opening, proofs = (np.ascontiguousarray(x, dtype=np.uint64) for x in (g2, opening, proofs))
    rc = lib.zk_kzg_verify(_p(g2), g2.reshape(-1, 24).shape[0], _p(np.ascontiguousarray(commitment)), _p(opening),
                           opening.reshape(-1, 4).shape[0], _p(np.ascontiguousarray(evaluation)), _p(proofs),
                           proofs.reshape(-1, 12).shape[0], C.byref(ok))
    return rc, ok.value


def test_generators_and_g2_setup_match_python_model(co, lib):
    g1 = np.zeros(12, dtype=np.uint64)
    g2 = np.zeros(24, dtype=np.uint64)
    lib.zk_g1_generator(_p(g1))
    lib.zk_g2_generator(_p(g2))
    assert co.g1_to_ints(g1)[0] == pk.G1_GEN and lib.zk_g1_is_on_curve(_p(g1)) == 1
    assert _g2_to_ints(co, g2)[0] == pk.G2_GEN
    rnd = random.Random(1)
    taus = [5, 2, rnd.randrange(R), R - 1]
    out = np.zeros((len(taus), 24), dtype=np.uint64)
    assert lib.zk_kzg_g2_powers_of_tau(_p(co.from_ints(FR, taus)), len(taus), _p(out)) == 0
    assert _g2_to_ints(co, out) == [pk.g2_mul(pk.G2_GEN, t) for t in taus]
    assert lib.zk_kzg_g2_powers_of_tau(_p(out), 0, _p(out)) == -1          # "requires at least one variable"


CASES = [  # multilinear_kzg/src/multilinear_kzg.rs:223-303
    ([5, 2, 3], [0, 4, 0, 4, 0, 4, 3, 7], [6, 4, 0]),
    ([2, 3, 4], [0, 7, 0, 5, 0, 7, 4, 9], [5, 9, 6]),
    ([12, 9, 28, 40], [0, 0, 0, 2, 0, 0, 10, 12, 0, -12, 4, -6, 0, -12, 14, 4], [54, 90, 76, 160]),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_reference_kzg_proofs_verify(co, lib, case):
    taus, vals, opening = ([x % R for x in v] for v in CASES[case])
    t, v, o = (co.from_ints(FR, x) for x in (taus, vals, opening))
    g1 = co.kzg_setup_g1(t)
    g2 = np.zeros((len(taus), 24), dtype=np.uint64)
    assert lib.zk_kzg_g2_powers_of_tau(_p(t), len(taus), _p(g2)) == 0
    c = co.kzg_commit(v, g1)
    ev, proofs = co.kzg_open(v, g1, o)
    assert verify(lib, g2, c, o, ev, proofs) == (0, 1)
    # a wrong evaluation, a wrong proof, a wrong commitment, a wrong opening point: rejected
    wrong_ev = co.from_ints(FR, [(co.to_ints(FR, ev)[0] + 1) % R])[0]
    assert verify(lib, g2, c, o, wrong_ev, proofs) == (0, 0)
    bad = proofs.copy()
    bad[1] = co.g1_add(bad[1], co.g1_generator())
    assert verify(lib, g2, c, o, ev, bad) == (0, 0)
    assert verify(lib, g2, co.g1_add(c, co.g1_generator()), o, ev, proofs) == (0, 0)
    o2 = o.copy()
    o2[0] = co.from_ints(FR, [123456789])[0]
    assert verify(lib, g2, c, o2, ev, proofs) == (0, 0)
    # argument errors
    assert verify(lib, g2, c, o[:-1], ev, proofs)[0] == -1                   # "Number of opening values must match number of proofs"
    off = proofs.copy()
    off[0, 0] ^= 1
    assert verify(lib, g2, c, o, ev, off)[0] == -3                           # a point off the curve


def test_golden_proofs_verify(co, lib, kzg_golden):
    """the committed golden commitments / openings (tests/golden/kzg_golden.json) through the product's pairing verifier"""
    from conftest import golden_point
    g2gen = np.zeros(24, dtype=np.uint64)
    lib.zk_g2_generator(_p(g2gen))
    want = kzg_golden["curve"]["g2_generator"]
    assert _g2_to_ints(co, g2gen)[0] == ((int(want[0][0], 16), int(want[0][1], 16)), (int(want[1][0], 16), int(want[1][1], 16)))
    for e in kzg_golden["generated"]:
        taus, opening = ([int(x) for x in e[k]] for k in ("taus", "opening"))
        t, o = co.from_ints(FR, taus), co.from_ints(FR, opening)
        g2 = np.zeros((len(taus), 24), dtype=np.uint64)
        assert lib.zk_kzg_g2_powers_of_tau(_p(t), len(taus), _p(g2)) == 0
        c = co.g1_from_ints([golden_point(e["commitment"])])[0]
        proofs = co.g1_from_ints([golden_point(p) for p in e["proofs"]])
        ev = co.from_ints(FR, [int(e["evaluation"])])[0]
        assert verify(lib, g2, c, o, ev, proofs) == (0, 1), e["src"]
        assert verify(lib, g2, c, o, co.from_ints(FR, [(int(e["evaluation"]) + 1) % R])[0], proofs) == (0, 0)


def test_random_polynomial_and_infinity_proofs(co, lib):
    rnd = random.Random(9)
    n = 5
    taus = [rnd.randrange(R) for _ in range(n)]
    t = co.from_ints(FR, taus)
    g1 = co.kzg_setup_g1(t)
    g2 = np.zeros((n, 24), dtype=np.uint64)
    lib.zk_kzg_g2_powers_of_tau(_p(t), n, _p(g2))
    vals = co.table_generate(FR, 5, 0, 1 << n)
    o = co.from_ints(FR, [rnd.randrange(R) for _ in range(n)])
    c = co.kzg_commit(vals, g1)
    ev, proofs = co.kzg_open(vals, g1, o)
    assert verify(lib, g2, c, o, ev, proofs) == (0, 1)
    # a constant polynomial: every quotient is zero, every proof the point at infinity
    const = co.from_ints(FR, [7] * (1 << n))
    c = co.kzg_commit(const, g1)
    ev, proofs = co.kzg_open(const, g1, o)
    assert not proofs.any() and co.to_ints(FR, ev)[0] == 7
    assert verify(lib, g2, c, o, ev, proofs) == (0, 1)


def test_pairing_is_bilinear_and_non_degenerate(co, lib):
    """e(aP, bQ) e(-abP, Q) == 1, e(P, Q) != 1, infinity pairs to one -- through zk_pairing_product_is_one, with the G2 multiples
    taken from the product's own setup routine and checked against the Python model"""
    rnd = random.Random(4)
    g1 = co.g1_generator()
    g2 = np.zeros(24, dtype=np.uint64)
    lib.zk_g2_generator(_p(g2))

    def check(p1s, p2s):
        ok = C.c_int(-1)
        a = np.ascontiguousarray(np.stack(p1s), dtype=np.uint64)
        b = np.ascontiguousarray(np.stack(p2s), dtype=np.uint64)
        rc = lib.zk_pairing_product_is_one(_p(a), _p(b), len(p1s), C.byref(ok))
        return rc, ok.value

    for _ in range(3):
        a, b = rnd.randrange(1, R), rnd.randrange(1, R)
        bq = np.zeros((1, 24), dtype=np.uint64)
        assert lib.zk_kzg_g2_powers_of_tau(_p(co.from_ints(FR, [b])), 1, _p(bq)) == 0          # b * G2
        assert _g2_to_ints(co, bq)[0] == pk.g2_mul(pk.G2_GEN, b)
        ap = co.g1_mul(g1, a)
        minus_abp = co.g1_mul(g1, (R - a * b % R) % R)
        assert check([ap, minus_abp], [bq[0], g2]) == (0, 1)
        assert check([ap, co.g1_mul(g1, (R - a * b % R + 1) % R)], [bq[0], g2]) == (0, 0)
    assert check([g1], [g2]) == (0, 0)                                                             # non-degenerate
    inf1, inf2 = np.zeros(12, dtype=np.uint64), np.zeros(24, dtype=np.uint64)
    assert check([inf1, g1], [g2, inf2]) == (0, 1)
    bad = g2.copy()
    bad[0] ^= 1
    assert check([g1], [bad])[0] == -3
