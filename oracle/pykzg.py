"""pykzg -- Python big-integer model of the reference's multilinear KZG and of the BLS12-381 pairing it verifies with.

TEST INFRASTRUCTURE ONLY (second, independent oracle: plain Python ints, affine formulas, no Montgomery form,
no code shared with oracle/zkoracle.c or the CUDA library).  Small sizes only.

The reference takes its curve from a third-party crate that is not under /root/reference:
`ark-bls12-381 0.5.0` / `ark-ec 0.5.0` (multilinear_kzg/Cargo.toml; `P::G1::generator().mul_bigint(..)`,
`P::pairing(..)`).  Restated here from the published definition of BLS12-381:
    E  : y^2 = x^3 + 4           over Fq,            G1 = the order-r subgroup,
    E' : y^2 = x^3 + 4 (1 + u)   over Fq2 = Fq[u]/(u^2+1)   (M-type sextic twist),  G2 = the order-r subgroup,
    Fq12 = Fq2[w]/(w^6 - (1+u)),  untwist (x', y') -> (x'/w^2, y'/w^3),
    e(P, Q) = f_{|x|, Q}(P)^((q^12-1)/r)  with  x = -0xd201000000010000  (ate pairing; any fixed power of it is as good for
    the equality test of multilinear_kzg.rs:136-154, the only use the reference makes of `PairingOutput`).
Self-checks (self_check()): generators on their curves and of order r, bilinearity e(aP, bQ) = e(P, Q)^(ab), non-degeneracy.

Reference functions restated (file:line relative to the reference root):
    trusted_setup.rs:12-24   initialize_setup          -> TrustedSetup.initialize
    trusted_setup.rs:26-52   compute_lagrange_basis    -> lagrange_basis
    trusted_setup.rs:54-63   compute_g1_powers_of_tau  -> (in initialize)
    trusted_setup.rs:65-78   compute_g2_powers_of_tau  -> (in initialize)
    multilinear_kzg.rs:25-46   commit_to_polynomial    -> commit
    multilinear_kzg.rs:51-127  open_and_prove          -> open_and_prove
    multilinear_kzg.rs:132-159 verify                  -> verify
    multilinear_kzg.rs:166-214 compute_quotient_polynomial / blow_up / expand_vec
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

Q = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
X_ABS = 0xD201000000010000       # |x|, x < 0

G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)
G2_GEN = (
    (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
     0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
    (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
     0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE),
)

G1Point = Optional[Tuple[int, int]]          # affine, None = point at infinity
Fq2 = Tuple[int, int]                        # c0 + c1 u
G2Point = Optional[Tuple[Fq2, Fq2]]


# ---------------------------------------------------------------- G1 (affine, chord and tangent)
def g1_is_on_curve(p: G1Point) -> bool:
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - 4) % Q == 0


def g1_neg(p: G1Point) -> G1Point:
    return None if p is None else (p[0], (-p[1]) % Q)


def g1_add(a: G1Point, b: G1Point) -> G1Point:
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % Q == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, Q) % Q
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q) % Q
    x3 = (lam * lam - x1 - x2) % Q
    return (x3, (lam * (x1 - x3) - y1) % Q)


def g1_mul(p: G1Point, k: int) -> G1Point:
    """`mul_bigint`: k is the canonical integer of the scalar (not reduced further)."""
    acc: G1Point = None
    for bit in bin(k)[2:] if k else "":
        acc = g1_add(acc, acc)
        if bit == "1":
            acc = g1_add(acc, p)
    return acc


# ---------------------------------------------------------------- Fq2, G2
def f2_add(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)


def f2_sub(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)


def f2_mul(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def f2_inv(a: Fq2) -> Fq2:
    n = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
    return (a[0] * n % Q, (-a[1]) * n % Q)


def f2_scalar(a: Fq2, k: int) -> Fq2:
    return (a[0] * k % Q, a[1] * k % Q)


XI: Fq2 = (1, 1)
B2: Fq2 = (4, 4)


def g2_is_on_curve(p: G2Point) -> bool:
    if p is None:
        return True
    x, y = p
    return f2_sub(f2_mul(y, y), f2_add(f2_mul(f2_mul(x, x), x), B2)) == (0, 0)


def g2_neg(p: G2Point) -> G2Point:
    return None if p is None else (p[0], f2_sub((0, 0), p[1]))


def g2_add(a: G2Point, b: G2Point) -> G2Point:
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if f2_add(y1, y2) == (0, 0):
            return None
        lam = f2_mul(f2_scalar(f2_mul(x1, x1), 3), f2_inv(f2_scalar(y1, 2)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), x1), x2)
    return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1))


def g2_mul(p: G2Point, k: int) -> G2Point:
    acc: G2Point = None
    for bit in bin(k)[2:] if k else "":
        acc = g2_add(acc, acc)
        if bit == "1":
            acc = g2_add(acc, p)
    return acc


# ---------------------------------------------------------------- Fq12 = Fq2[w]/(w^6 - xi), six Fq2 coefficients
F12 = List[Fq2]
F12_ONE: F12 = [(1, 0)] + [(0, 0)] * 5


def f12_mul(a: F12, b: F12) -> F12:
    t = [(0, 0)] * 11
    for i, ai in enumerate(a):
        if ai == (0, 0):
            continue
        for j, bj in enumerate(b):
            if bj == (0, 0):
                continue
            t[i + j] = f2_add(t[i + j], f2_mul(ai, bj))
    return [f2_add(t[k], f2_mul(t[k + 6], XI)) if k < 5 else t[k] for k in range(6)]


def f12_pow(a: F12, e: int) -> F12:
    acc = F12_ONE
    for bit in bin(e)[2:]:
        acc = f12_mul(acc, acc)
        if bit == "1":
            acc = f12_mul(acc, a)
    return acc


def _line(t: G2Point, lam: Fq2, p: Tuple[int, int]) -> F12:
    """The line through T with slope lam (both on the twist) evaluated at P in G1, times w^3 (a factor in Fq4, which the
    final exponentiation removes): (lam x_T - y_T) - lam x_P w^2 + y_P w^3."""
    xt, yt = t
    return [f2_sub(f2_mul(lam, xt), yt), (0, 0), f2_scalar(lam, (-p[0]) % Q), (p[1] % Q, 0), (0, 0), (0, 0)]


def miller_loop(p: G1Point, q: G2Point) -> F12:
    if p is None or q is None:
        return F12_ONE
    f = F12_ONE
    t = q
    for bit in bin(X_ABS)[3:]:
        lam = f2_mul(f2_scalar(f2_mul(t[0], t[0]), 3), f2_inv(f2_scalar(t[1], 2)))
        f = f12_mul(f12_mul(f, f), _line(t, lam, p))
        t = g2_add(t, t)
        if bit == "1":
            lam = f2_mul(f2_sub(q[1], t[1]), f2_inv(f2_sub(q[0], t[0])))
            f = f12_mul(f, _line(t, lam, p))
            t = g2_add(t, q)
    return f


FINAL_EXP = (Q ** 12 - 1) // R


def final_exponentiation(f: F12) -> F12:
    return f12_pow(f, FINAL_EXP)


def pairing(p: G1Point, q: G2Point) -> F12:
    return final_exponentiation(miller_loop(p, q))


def pairing_product_is_one(pairs: Sequence[Tuple[G1Point, G2Point]]) -> bool:
    f = F12_ONE
    for p, q in pairs:
        f = f12_mul(f, miller_loop(p, q))
    return final_exponentiation(f) == F12_ONE


# ---------------------------------------------------------------- trusted setup (trusted_setup.rs)
def lagrange_basis(taus: Sequence[int]) -> List[int]:           # trusted_setup.rs:26-52
    n = len(taus)
    assert n > 0, "requires at least one variable"
    out = []
    for index in range(1 << n):
        e = 1
        for i in range(n):
            bit = (index >> (n - 1 - i)) & 1
            e = e * (taus[i] if bit else (1 - taus[i])) % R
        out.append(e)
    return out


class TrustedSetup:                                             # trusted_setup.rs:5-24
    def __init__(self, g1_powers: List[G1Point], g2_powers: List[G2Point]):
        self.g1_powers_of_tau = g1_powers
        self.g2_powers_of_tau = g2_powers

    @classmethod
    def initialize(cls, taus: Sequence[int]) -> "TrustedSetup":
        basis = lagrange_basis(taus)
        return cls([g1_mul(G1_GEN, e) for e in basis], [g2_mul(G2_GEN, t % R) for t in taus])


# ---------------------------------------------------------------- multilinear KZG (multilinear_kzg.rs)
def _partial_evaluate_first(vals: Sequence[int], r: int) -> List[int]:   # evaluation_form.rs:61-106 with var = 0
    half = len(vals) // 2
    return [(vals[j] + r * (vals[j + half] - vals[j])) % R for j in range(half)]


def mle_evaluate(vals: Sequence[int], point: Sequence[int]) -> int:      # evaluation_form.rs:21-33
    cur = list(vals)
    for r in point:
        cur = _partial_evaluate_first(cur, r)
    return cur[0]


def _dot(vals: Sequence[int], points: Sequence[G1Point]) -> G1Point:
    acc: G1Point = None
    for v, p in zip(vals, points):
        acc = g1_add(acc, g1_mul(p, v))
    return acc


def commit(vals: Sequence[int], setup: TrustedSetup) -> G1Point:         # multilinear_kzg.rs:25-46
    assert len(vals) == len(setup.g1_powers_of_tau), "Polynomial evaluation must match g1 length"
    return _dot(vals, setup.g1_powers_of_tau)


def open_and_prove(vals: Sequence[int], setup: TrustedSetup, opening: Sequence[int]) -> Tuple[int, List[G1Point]]:
    """multilinear_kzg.rs:51-127 -> (evaluation, proofs)."""
    n = len(opening)
    assert len(vals) == 1 << n, "number of polynomial variables must match length of opening values"
    assert n == len(setup.g2_powers_of_tau), "Opening values must match number of variables from trusted setup"
    v = mle_evaluate(vals, opening)
    sub = [(x - v) % R for x in vals]
    proofs = []
    for i in range(n):
        half = len(sub) // 2
        quotient = [(sub[j + half] - sub[j]) % R for j in range(half)]           # :166-181
        blown = quotient
        for _ in range(i + 1):                                                    # :183-214
            blown = blown + blown
        proofs.append(_dot(blown, setup.g1_powers_of_tau))
        sub = _partial_evaluate_first(sub, opening[i])
    return v, proofs


def verify(setup: TrustedSetup, commitment: G1Point, opening: Sequence[int], evaluation: int,
           proofs: Sequence[G1Point]) -> bool:                                    # multilinear_kzg.rs:132-159
    assert len(opening) == len(proofs), "Number of opening values must match number of proofs"
    lhs_point = g1_add(commitment, g1_neg(g1_mul(G1_GEN, evaluation % R)))
    pairs = [(lhs_point, G2_GEN)]
    for i, tau_g2 in enumerate(setup.g2_powers_of_tau):
        pairs.append((g1_neg(proofs[i]), g2_add(tau_g2, g2_neg(g2_mul(G2_GEN, opening[i] % R)))))
    return pairing_product_is_one(pairs)


def verify_with_trapdoor(taus: Sequence[int], commitment: G1Point, opening: Sequence[int], evaluation: int,
                         proofs: Sequence[G1Point]) -> bool:
    """The same equation checked in G1 with the toxic waste known (tests only; no pairing):
    C - v G == sum_i (tau_i - r_i) Q_i."""
    lhs = g1_add(commitment, g1_neg(g1_mul(G1_GEN, evaluation % R)))
    rhs: G1Point = None
    for t, r, q in zip(taus, opening, proofs):
        rhs = g1_add(rhs, g1_mul(q, (t - r) % R))
    return lhs == rhs


def self_check() -> None:
    assert g1_is_on_curve(G1_GEN) and g1_mul(G1_GEN, R) is None
    assert g2_is_on_curve(G2_GEN) and g2_mul(G2_GEN, R) is None
    e = pairing(G1_GEN, G2_GEN)
    assert e != F12_ONE and f12_pow(e, R) == F12_ONE
    a, b = 0x1234567, 0x89ABCDEF01
    assert pairing(g1_mul(G1_GEN, a), g2_mul(G2_GEN, b)) == f12_pow(e, a * b % R)
