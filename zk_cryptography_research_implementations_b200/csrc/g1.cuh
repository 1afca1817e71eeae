// g1.cuh -- BLS12-381 G1 (y^2 = x^3 + 4 over Fq) group law for the multi-scalar multiplications of the multilinear KZG
// (`power.mul_bigint(value).sum()`, multilinear_kzg/src/multilinear_kzg.rs:39-43 and :101-108).
//
// Accumulators are kept in extended Jacobian ("XYZZ") coordinates: x = X/ZZ, y = Y/ZZZ with ZZ^3 == ZZZ^2; the point at
// infinity is ZZ == 0.  Adding an affine point costs 8 products + 2 squares and no field inversion.  Inputs and results
// cross the C-ABI in affine form (canonical), so the coordinate system never shows in a result.
// Affine points: (x, y) Montgomery limbs, infinity = all zero (the pair (0, 0) is not on the curve).
#pragma once
#include "fq381.cuh"

namespace zk {

struct G1Affine {
    Fq x, y;
};
struct G1Xyzz {
    Fq x, y, zz, zzz;
};

struct G1 {
    typedef Fq381 F;
    ZK_DEV static bool is_inf(const G1Affine& p) { return F::is_zero(p.x) && F::is_zero(p.y); }
    ZK_DEV static bool is_inf(const G1Xyzz& p) { return F::is_zero(p.zz); }
    ZK_DEV static G1Xyzz infinity() {
        G1Xyzz r;
        r.x = F::zero();
        r.y = F::zero();
        r.zz = F::zero();
        r.zzz = F::zero();
        return r;
    }
    ZK_DEV static G1Xyzz from_affine(const G1Affine& p) {
        G1Xyzz r;
        if (is_inf(p)) return infinity();
        r.x = p.x;
        r.y = p.y;
        r.zz = F::mont_one();
        r.zzz = F::mont_one();
        return r;
    }
    // 2 * (affine p), p not infinity
    ZK_DEV static void dbl_affine(G1Xyzz& r, const G1Affine& p) {
        Fq u, v, w, s, m, t;
        F::dbl(u, p.y);
        F::sqr(v, u);
        F::mul(w, u, v);
        F::mul(s, p.x, v);
        F::sqr(t, p.x);
        F::dbl(m, t);
        F::add(m, m, t);
        F::sqr(r.x, m);
        F::sub(r.x, r.x, s);
        F::sub(r.x, r.x, s);
        F::sub(t, s, r.x);
        F::mul(t, m, t);
        F::mul(u, w, p.y);
        F::sub(r.y, t, u);
        r.zz = v;
        r.zzz = w;
    }
    ZK_DEV static void dbl(G1Xyzz& r, const G1Xyzz& p) {
        if (is_inf(p)) { r = p; return; }
        Fq u, v, w, s, m, t, x3;
        F::dbl(u, p.y);
        F::sqr(v, u);
        F::mul(w, u, v);
        F::mul(s, p.x, v);
        F::sqr(t, p.x);
        F::dbl(m, t);
        F::add(m, m, t);
        F::sqr(x3, m);
        F::sub(x3, x3, s);
        F::sub(x3, x3, s);
        F::sub(t, s, x3);
        F::mul(t, m, t);
        F::mul(u, w, p.y);
        F::sub(r.y, t, u);
        r.x = x3;
        F::mul(r.zz, v, p.zz);
        F::mul(r.zzz, w, p.zzz);
    }
    // acc += p (affine); `negate` adds -p
    ZK_DEV static void add_affine(G1Xyzz& acc, const G1Affine& p_in, bool negate = false) {
        if (is_inf(p_in)) return;
        G1Affine p = p_in;
        if (negate) F::neg(p.y, p_in.y);
        if (is_inf(acc)) { acc = from_affine(p); return; }
        Fq u2, s2, pp, ppp, q, r, t;
        F::mul(u2, p.x, acc.zz);
        F::mul(s2, p.y, acc.zzz);
        F::sub(u2, u2, acc.x);      // P
        F::sub(r, s2, acc.y);       // R
        if (F::is_zero(u2)) {
            if (F::is_zero(r)) dbl_affine(acc, p);
            else acc = infinity();
            return;
        }
        F::sqr(pp, u2);
        F::mul(ppp, u2, pp);
        F::mul(q, acc.x, pp);
        F::sqr(t, r);
        F::sub(t, t, ppp);
        F::sub(t, t, q);
        F::sub(t, t, q);            // X3
        F::sub(q, q, t);
        F::mul(q, r, q);
        F::mul(s2, acc.y, ppp);
        F::sub(acc.y, q, s2);
        acc.x = t;
        F::mul(acc.zz, acc.zz, pp);
        F::mul(acc.zzz, acc.zzz, ppp);
    }
    // acc += b
    ZK_DEV static void add(G1Xyzz& acc, const G1Xyzz& b) {
        if (is_inf(b)) return;
        if (is_inf(acc)) { acc = b; return; }
        Fq u1, u2, s1, s2, pp, ppp, q, r, t;
        F::mul(u1, acc.x, b.zz);
        F::mul(u2, b.x, acc.zz);
        F::mul(s1, acc.y, b.zzz);
        F::mul(s2, b.y, acc.zzz);
        F::sub(u2, u2, u1);         // P
        F::sub(r, s2, s1);          // R
        if (F::is_zero(u2)) {
            if (F::is_zero(r)) dbl(acc, acc);
            else acc = infinity();
            return;
        }
        F::sqr(pp, u2);
        F::mul(ppp, u2, pp);
        F::mul(q, u1, pp);
        F::sqr(t, r);
        F::sub(t, t, ppp);
        F::sub(t, t, q);
        F::sub(t, t, q);            // X3
        F::sub(q, q, t);
        F::mul(q, r, q);
        F::mul(s1, s1, ppp);
        F::sub(acc.y, q, s1);
        acc.x = t;
        F::mul(acc.zz, acc.zz, b.zz);
        F::mul(acc.zz, acc.zz, pp);
        F::mul(acc.zzz, acc.zzz, b.zzz);
        F::mul(acc.zzz, acc.zzz, ppp);
    }
    // a^(q-2): the inverse of a non-zero a (0 -> 0)
    ZK_DEV static void inv(Fq& r, const Fq& a) {
        Fq acc = F::mont_one();
#pragma unroll 1
        for (int i = 380; i >= 0; --i) {
            F::sqr(acc, acc);
            uint32_t word = F::q(0);   // exponent q - 2: q is odd and q(0) ends ...aaab, so only limb 0 changes
#pragma unroll
            for (int k = 0; k < 12; ++k)
                if (k == (i >> 5)) word = k == 0 ? F::q(0) - 2u : F::q(k);
            if ((word >> (i & 31)) & 1u) F::mul(acc, acc, a);
        }
        r = acc;
    }
    // affine coordinates of a finite XYZZ point given i3 = 1/ZZZ:  1/ZZ = i3^2 * ZZ^2
    ZK_DEV static void to_affine_with_inverse(G1Affine& r, const G1Xyzz& p, const Fq& i3) {
        Fq t, izz;
        F::sqr(t, i3);
        F::sqr(izz, p.zz);
        F::mul(izz, izz, t);
        F::mul(r.x, p.x, izz);
        F::mul(r.y, p.y, i3);
    }
};

}  // namespace zk
