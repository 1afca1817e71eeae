// host_pairing.h -- the verifier's side of the multilinear KZG on the host: BLS12-381 G2 and the pairing check
//   e(C - v G1, G2) == prod_i e(Q_i, tau_i G2 - r_i G2)            (multilinear_kzg/src/multilinear_kzg.rs:132-159)
// evaluated as one product of Miller loops followed by one final exponentiation.  The reference gets `P::pairing` from
// ark-ec / ark-bls12-381 0.5.0 (not under /root/reference); this is the published construction restated:
//   Fq2 = Fq[u]/(u^2 + 1),  E': y^2 = x^3 + 4 (1 + u) (M-type twist),  Fq12 = Fq2[w]/(w^6 - (1 + u)),
//   untwist (x', y') -> (x'/w^2, y'/w^3),  ate loop over |x| = 0xd201000000010000,  final exponent (q^12 - 1)/r.
// `PairingOutput` is only ever compared for equality by the reference, and any fixed power of the ate pairing is
// bilinear and non-degenerate, so the check accepts exactly the proofs the reference's accepts.
// A verifier runs this a handful of times per proof (n + 1 Miller loops, ~2 ms each, one ~30 ms exponentiation);
// nothing here is on the prover's path.  Independent of oracle/.
#pragma once
#include <vector>
#include "host_curve.h"

namespace zk {

struct HFq2 {
    HFq c0, c1;
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const HFq2& o) const { return c0 == o.c0 && c1 == o.c1; }
};
struct HostFq2 {
    typedef HostFq F;
    static HFq2 zero() { return HFq2{F::zero(), F::zero()}; }
    static HFq2 one() { return HFq2{F::one(), F::zero()}; }
    static HFq2 from_fq(const HFq& a) { return HFq2{a, F::zero()}; }
    static HFq2 add(const HFq2& a, const HFq2& b) { return HFq2{F::add(a.c0, b.c0), F::add(a.c1, b.c1)}; }
    static HFq2 sub(const HFq2& a, const HFq2& b) { return HFq2{F::sub(a.c0, b.c0), F::sub(a.c1, b.c1)}; }
    static HFq2 neg(const HFq2& a) { return HFq2{F::neg(a.c0), F::neg(a.c1)}; }
    static HFq2 dbl(const HFq2& a) { return add(a, a); }
    static HFq2 mul(const HFq2& a, const HFq2& b) {   // (a0 + a1 u)(b0 + b1 u), u^2 = -1, three products
        const HFq t0 = F::mul(a.c0, b.c0), t1 = F::mul(a.c1, b.c1);
        const HFq t2 = F::mul(F::add(a.c0, a.c1), F::add(b.c0, b.c1));
        return HFq2{F::sub(t0, t1), F::sub(F::sub(t2, t0), t1)};
    }
    static HFq2 sqr(const HFq2& a) { return mul(a, a); }
    static HFq2 mul_fq(const HFq2& a, const HFq& k) { return HFq2{F::mul(a.c0, k), F::mul(a.c1, k)}; }
    static HFq2 mul_xi(const HFq2& a) { return HFq2{F::sub(a.c0, a.c1), F::add(a.c0, a.c1)}; }   // times 1 + u
    static HFq2 inv(const HFq2& a) {
        const HFq n = F::inv(F::add(F::sqr(a.c0), F::sqr(a.c1)));
        return HFq2{F::mul(a.c0, n), F::neg(F::mul(a.c1, n))};
    }
};

// affine G2 point as arkworks lays out `G2Affine`'s coordinates: x.c0, x.c1, y.c0, y.c1; infinity = all zero
struct HG2Affine {
    HFq2 x, y;
    bool is_inf() const { return x.is_zero() && y.is_zero(); }
};
struct HG2Jac {   // x = X/Z^2, y = Y/Z^3, infinity: Z == 0
    HFq2 x, y, z;
    bool is_inf() const { return z.is_zero(); }
};

struct HostG2 {
    typedef HostFq2 F2;
    static HG2Affine generator() {
        HG2Affine g;
        memcpy(g.x.c0.l, ZKC_G2X0_MONT_64, 48);
        memcpy(g.x.c1.l, ZKC_G2X1_MONT_64, 48);
        memcpy(g.y.c0.l, ZKC_G2Y0_MONT_64, 48);
        memcpy(g.y.c1.l, ZKC_G2Y1_MONT_64, 48);
        return g;
    }
    static HFq2 b_twist() {   // 4 (1 + u)
        HFq b;
        memcpy(b.l, ZKC_B_MONT_64, 48);
        return HFq2{b, b};
    }
    static bool on_curve(const HG2Affine& p) {
        if (p.is_inf()) return true;
        if (HostFq::geq_q(p.x.c0.l) || HostFq::geq_q(p.x.c1.l) || HostFq::geq_q(p.y.c0.l) || HostFq::geq_q(p.y.c1.l)) return false;
        return F2::sqr(p.y) == F2::add(F2::mul(F2::sqr(p.x), p.x), b_twist());
    }
    static HG2Jac infinity() { return HG2Jac{F2::zero(), F2::zero(), F2::zero()}; }
    static HG2Jac from_affine(const HG2Affine& p) { return p.is_inf() ? infinity() : HG2Jac{p.x, p.y, F2::one()}; }
    static HG2Jac dbl(const HG2Jac& p) {
        if (p.is_inf()) return p;
        const HFq2 a = F2::sqr(p.x), b = F2::sqr(p.y), c = F2::sqr(b);
        const HFq2 d = F2::dbl(F2::sub(F2::sub(F2::sqr(F2::add(p.x, b)), a), c));
        const HFq2 e = F2::add(F2::dbl(a), a), f = F2::sqr(e);
        HG2Jac r;
        r.x = F2::sub(f, F2::dbl(d));
        r.y = F2::sub(F2::mul(e, F2::sub(d, r.x)), F2::dbl(F2::dbl(F2::dbl(c))));
        r.z = F2::dbl(F2::mul(p.y, p.z));
        return r;
    }
    static HG2Jac add(const HG2Jac& p, const HG2Jac& q) {
        if (p.is_inf()) return q;
        if (q.is_inf()) return p;
        const HFq2 z1z1 = F2::sqr(p.z), z2z2 = F2::sqr(q.z);
        const HFq2 u1 = F2::mul(p.x, z2z2), u2 = F2::mul(q.x, z1z1);
        const HFq2 s1 = F2::mul(F2::mul(p.y, q.z), z2z2), s2 = F2::mul(F2::mul(q.y, p.z), z1z1);
        if (u1 == u2) return s1 == s2 ? dbl(p) : infinity();
        const HFq2 h = F2::sub(u2, u1), i = F2::sqr(F2::dbl(h)), j = F2::mul(h, i);
        const HFq2 rr = F2::dbl(F2::sub(s2, s1)), v = F2::mul(u1, i);
        HG2Jac r;
        r.x = F2::sub(F2::sub(F2::sqr(rr), j), F2::dbl(v));
        r.y = F2::sub(F2::mul(rr, F2::sub(v, r.x)), F2::dbl(F2::mul(s1, j)));
        r.z = F2::mul(F2::sub(F2::sub(F2::sqr(F2::add(p.z, q.z)), z1z1), z2z2), h);
        return r;
    }
    static HG2Affine to_affine(const HG2Jac& p) {
        if (p.is_inf()) return HG2Affine{F2::zero(), F2::zero()};
        const HFq2 zi = F2::inv(p.z), zi2 = F2::sqr(zi);
        return HG2Affine{F2::mul(p.x, zi2), F2::mul(p.y, F2::mul(zi2, zi))};
    }
    static HG2Affine neg(const HG2Affine& p) { return HG2Affine{p.x, F2::neg(p.y)}; }
    // k * p for a canonical 256-bit little-endian integer k (`mul_bigint`)
    static HG2Affine mul(const HG2Affine& p, const uint64_t k[4]) {
        const HG2Jac base = from_affine(p);
        HG2Jac acc = infinity();
        for (int i = 255; i >= 0; --i) {
            acc = dbl(acc);
            if ((k[i >> 6] >> (i & 63)) & 1) acc = add(acc, base);
        }
        return to_affine(acc);
    }
    static HG2Affine add_affine(const HG2Affine& a, const HG2Affine& b) { return to_affine(add(from_affine(a), from_affine(b))); }
};

// ---- Fq12 = Fq2[w]/(w^6 - xi): six Fq2 coefficients
struct HFq12 {
    HFq2 c[6];
    bool operator==(const HFq12& o) const {
        for (int i = 0; i < 6; ++i)
            if (!(c[i] == o.c[i])) return false;
        return true;
    }
};
struct HostFq12 {
    typedef HostFq2 F2;
    static HFq12 one() {
        HFq12 r;
        for (int i = 0; i < 6; ++i) r.c[i] = F2::zero();
        r.c[0] = F2::one();
        return r;
    }
    static HFq12 mul(const HFq12& a, const HFq12& b) {
        HFq2 t[11];
        for (int i = 0; i < 11; ++i) t[i] = F2::zero();
        for (int i = 0; i < 6; ++i) {
            if (a.c[i].is_zero()) continue;
            for (int j = 0; j < 6; ++j) {
                if (b.c[j].is_zero()) continue;
                t[i + j] = F2::add(t[i + j], F2::mul(a.c[i], b.c[j]));
            }
        }
        HFq12 r;
        for (int k = 0; k < 6; ++k) r.c[k] = k < 5 ? F2::add(t[k], F2::mul_xi(t[k + 6])) : t[k];
        return r;
    }
    static HFq12 pow(const HFq12& a, const uint64_t* e, int bits) {
        HFq12 acc = one();
        for (int i = bits - 1; i >= 0; --i) {
            acc = mul(acc, acc);
            if ((e[i >> 6] >> (i & 63)) & 1) acc = mul(acc, a);
        }
        return acc;
    }
};

struct HostPairing {
    typedef HostFq2 F2;
    // the line through T with slope lam (on the twist) at P in G1, times w^3 (an Fq4 factor the final exponentiation
    // removes): (lam x_T - y_T) - lam x_P w^2 + y_P w^3
    static HFq12 line(const HG2Affine& t, const HFq2& lam, const HG1Affine& p) {
        HFq12 l;
        for (int i = 0; i < 6; ++i) l.c[i] = F2::zero();
        l.c[0] = F2::sub(F2::mul(lam, t.x), t.y);
        l.c[2] = F2::neg(F2::mul_fq(lam, p.x));
        l.c[3] = F2::from_fq(p.y);
        return l;
    }
    static HFq12 miller_loop(const HG1Affine& p, const HG2Affine& q) {
        HFq12 f = HostFq12::one();
        if (p.is_inf() || q.is_inf()) return f;
        HG2Affine t = q;
        for (int i = 62; i >= 0; --i) {   // |x| has 64 bits; the top one starts T = Q
            const HFq2 x2 = F2::sqr(t.x);
            HFq2 lam = F2::mul(F2::add(F2::dbl(x2), x2), F2::inv(F2::dbl(t.y)));
            f = HostFq12::mul(HostFq12::mul(f, f), line(t, lam, p));
            HFq2 x3 = F2::sub(F2::sqr(lam), F2::dbl(t.x));
            t = HG2Affine{x3, F2::sub(F2::mul(lam, F2::sub(t.x, x3)), t.y)};
            if ((ZKC_X_ABS >> i) & 1) {
                lam = F2::mul(F2::sub(q.y, t.y), F2::inv(F2::sub(q.x, t.x)));
                f = HostFq12::mul(f, line(t, lam, p));
                x3 = F2::sub(F2::sub(F2::sqr(lam), t.x), q.x);
                t = HG2Affine{x3, F2::sub(F2::mul(lam, F2::sub(t.x, x3)), t.y)};
            }
        }
        return f;
    }
    static HFq12 final_exponentiation(const HFq12& f) { return HostFq12::pow(f, ZKC_FINAL_EXP_64, ZKC_FINAL_EXP_BITS); }
    // prod_i e(p_i, q_i) == 1
    static bool product_is_one(const std::vector<HG1Affine>& ps, const std::vector<HG2Affine>& qs) {
        HFq12 f = HostFq12::one();
        for (size_t i = 0; i < ps.size(); ++i) f = HostFq12::mul(f, miller_loop(ps[i], qs[i]));
        return final_exponentiation(f) == HostFq12::one();
    }
};

}  // namespace zk
