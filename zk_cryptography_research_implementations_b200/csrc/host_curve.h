// host_curve.h -- host-side BLS12-381 base field and G1 group law: the few hundred group operations per multi-scalar
// multiplication that are serial by nature (combining the window sums, one inversion for the affine result) and the
// one-time tables of the trusted setup.  6 x 64-bit limbs, Montgomery form with R = 2^384 (arkworks' MontBackend<_,6>
// layout, so a `P::G1Affine` coordinate is this struct byte for byte).
//
// Independent of oracle/ (the product must not use the checker).
#pragma once
#include <stdint.h>
#include <string.h>
#include "curve_consts.h"

namespace zk {

typedef unsigned __int128 u128;

struct HFq {
    uint64_t l[6];
    bool operator==(const HFq& o) const { return memcmp(l, o.l, 48) == 0; }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3] | l[4] | l[5]) == 0; }
};

struct HostFq {
    static HFq zero() { return HFq{{0, 0, 0, 0, 0, 0}}; }
    static HFq one() { HFq r; memcpy(r.l, ZKC_R_64, 48); return r; }
    static bool geq_q(const uint64_t* a) {
        for (int i = 5; i >= 0; --i) {
            if (a[i] != ZKC_Q_64[i]) return a[i] > ZKC_Q_64[i];
        }
        return true;
    }
    static void sub_q(uint64_t* a) {
        uint64_t borrow = 0;
        for (int i = 0; i < 6; ++i) {
            u128 d = (u128)a[i] - ZKC_Q_64[i] - borrow;
            a[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
    }
    static HFq add(const HFq& a, const HFq& b) {
        HFq r;
        u128 c = 0;
        for (int i = 0; i < 6; ++i) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
        if (geq_q(r.l)) sub_q(r.l);
        return r;
    }
    static HFq sub(const HFq& a, const HFq& b) {
        HFq r;
        uint64_t borrow = 0;
        for (int i = 0; i < 6; ++i) {
            u128 d = (u128)a.l[i] - b.l[i] - borrow;
            r.l[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
        if (borrow) {
            u128 c = 0;
            for (int i = 0; i < 6; ++i) { c += (u128)r.l[i] + ZKC_Q_64[i]; r.l[i] = (uint64_t)c; c >>= 64; }
        }
        return r;
    }
    static HFq neg(const HFq& a) { return sub(zero(), a); }
    static HFq dbl(const HFq& a) { return add(a, a); }
    // Montgomery product a*b/R: full product, then word-by-word reduction
    static HFq mul(const HFq& a, const HFq& b) {
        uint64_t t[13] = {0};
        for (int i = 0; i < 6; ++i) {
            u128 c = 0;
            for (int j = 0; j < 6; ++j) {
                c += (u128)a.l[i] * b.l[j] + t[i + j];
                t[i + j] = (uint64_t)c;
                c >>= 64;
            }
            t[i + 6] = (uint64_t)c;
        }
        for (int i = 0; i < 6; ++i) {
            const uint64_t m = t[i] * ZKC_INV64;
            u128 c = 0;
            for (int j = 0; j < 6; ++j) {
                c += (u128)m * ZKC_Q_64[j] + t[i + j];
                t[i + j] = (uint64_t)c;
                c >>= 64;
            }
            for (int k = i + 6; c && k < 13; ++k) {
                c += t[k];
                t[k] = (uint64_t)c;
                c >>= 64;
            }
        }
        HFq r;
        memcpy(r.l, t + 6, 48);
        if (t[12] || geq_q(r.l)) sub_q(r.l);
        return r;
    }
    static HFq sqr(const HFq& a) { return mul(a, a); }
    static HFq from_canonical(const uint64_t c[6]) {
        HFq a, r2;
        memcpy(a.l, c, 48);
        memcpy(r2.l, ZKC_R2_64, 48);
        return mul(a, r2);
    }
    static HFq inv(const HFq& a) {   // a^(q-2); 0 -> 0
        uint64_t e[6];
        memcpy(e, ZKC_Q_64, 48);
        e[0] -= 2;
        HFq acc = one();
        for (int i = 380; i >= 0; --i) {
            acc = sqr(acc);
            if ((e[i >> 6] >> (i & 63)) & 1) acc = mul(acc, a);
        }
        return acc;
    }
};

// affine point as it crosses the C-ABI: x then y, infinity = all zero
struct HG1Affine {
    HFq x, y;
    bool is_inf() const { return x.is_zero() && y.is_zero(); }
};
// x = X/ZZ, y = Y/ZZZ, ZZ^3 == ZZZ^2; infinity: ZZ == 0 (the device's accumulator layout, g1.cuh)
struct HG1Xyzz {
    HFq x, y, zz, zzz;
    bool is_inf() const { return zz.is_zero(); }
};

struct HostG1 {
    typedef HostFq F;
    static HG1Xyzz infinity() { return HG1Xyzz{F::zero(), F::zero(), F::zero(), F::zero()}; }
    static HG1Affine generator() {
        HG1Affine g;
        memcpy(g.x.l, ZKC_GX_MONT_64, 48);
        memcpy(g.y.l, ZKC_GY_MONT_64, 48);
        return g;
    }
    static HG1Xyzz from_affine(const HG1Affine& p) {
        if (p.is_inf()) return infinity();
        return HG1Xyzz{p.x, p.y, F::one(), F::one()};
    }
    static bool on_curve(const HG1Affine& p) {
        if (p.is_inf()) return true;
        if (F::geq_q(p.x.l) || F::geq_q(p.y.l)) return false;
        HFq b;
        memcpy(b.l, ZKC_B_MONT_64, 48);
        return F::sqr(p.y) == F::add(F::mul(F::sqr(p.x), p.x), b);
    }
    static HG1Xyzz dbl(const HG1Xyzz& p) {
        if (p.is_inf()) return p;
        const HFq u = F::dbl(p.y), v = F::sqr(u), w = F::mul(u, v), s = F::mul(p.x, v);
        const HFq x2 = F::sqr(p.x), m = F::add(F::dbl(x2), x2);
        HG1Xyzz r;
        r.x = F::sub(F::sub(F::sqr(m), s), s);
        r.y = F::sub(F::mul(m, F::sub(s, r.x)), F::mul(w, p.y));
        r.zz = F::mul(v, p.zz);
        r.zzz = F::mul(w, p.zzz);
        return r;
    }
    static HG1Xyzz add(const HG1Xyzz& a, const HG1Xyzz& b) {
        if (a.is_inf()) return b;
        if (b.is_inf()) return a;
        const HFq u1 = F::mul(a.x, b.zz), u2 = F::mul(b.x, a.zz), s1 = F::mul(a.y, b.zzz), s2 = F::mul(b.y, a.zzz);
        const HFq p = F::sub(u2, u1), r = F::sub(s2, s1);
        if (p.is_zero()) return r.is_zero() ? dbl(a) : infinity();
        const HFq pp = F::sqr(p), ppp = F::mul(p, pp), q = F::mul(u1, pp);
        HG1Xyzz o;
        o.x = F::sub(F::sub(F::sub(F::sqr(r), ppp), q), q);
        o.y = F::sub(F::mul(r, F::sub(q, o.x)), F::mul(s1, ppp));
        o.zz = F::mul(F::mul(a.zz, b.zz), pp);
        o.zzz = F::mul(F::mul(a.zzz, b.zzz), ppp);
        return o;
    }
    static HG1Xyzz neg(const HG1Xyzz& a) {
        HG1Xyzz r = a;
        r.y = F::neg(a.y);
        return r;
    }
    static HG1Affine to_affine(const HG1Xyzz& p) {
        if (p.is_inf()) return HG1Affine{F::zero(), F::zero()};
        const HFq i3 = F::inv(p.zzz), izz = F::mul(F::sqr(i3), F::sqr(p.zz));
        return HG1Affine{F::mul(p.x, izz), F::mul(p.y, i3)};
    }
    // k * p for a canonical little-endian integer k of `bits` bits
    static HG1Xyzz mul(const HG1Xyzz& p, const uint64_t* k, int bits) {
        HG1Xyzz acc = infinity();
        for (int i = bits - 1; i >= 0; --i) {
            acc = dbl(acc);
            if ((k[i >> 6] >> (i & 63)) & 1) acc = add(acc, p);
        }
        return acc;
    }
};

}  // namespace zk
