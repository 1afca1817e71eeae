// fp.cuh -- 256-bit prime-field arithmetic for sm_100a, 8 x 32-bit limbs in registers.
//
// Elements are arkworks `Fp<MontBackend<_,4>,4>` values: canonical (< p), Montgomery form
// (R = 2^256), 4 x u64 little-endian limbs == 8 x u32 little-endian limbs, 32 bytes.
// (Reference call sites: polynomials/src/multilinear/evaluation_form.rs:89-91 `y1 + r*(y2-y1)`,
//  composed/product_polynomial.rs:66-70 `*=`, composed/sum_polynomial.rs:67-73 `+=`.)
//
// Everything is built from 32x32->64 multiply-adds written as mad.lo.cc/madc.hi.cc pairs on
// (even, odd) limb pairs, which ptxas turns into IMAD.WIDE.U32(.X): one fma-pipe instruction per
// limb product.  Two accumulators are kept per number -- `E` aligned to even limb positions and
// `O` aligned to odd ones -- so that every row of a schoolbook product is a chain of four
// non-overlapping 64-bit slots with the carry rippling between them.
//
//   mont_mul        136 IMAD   full Montgomery product (CIOS, E/O interleaved)
//   mul_wide         64 IMAD   unreduced 512-bit product (lazy accumulation of round sums)
//   FoldScalar::fold 83 IMAD   lo + r*(hi-lo) for a per-round constant r: the 8 multiples
//                              r*2^(32i) mod p are precomputed on the host, the <2^291 result is
//                              brought back to [0,p) by one Barrett step
//   redc_wide                  Montgomery-reduce a 17-limb lazy accumulator (once per thread)
//
// Bit-exactness with the reference is algebraic: every routine returns the canonical
// representative, so any evaluation order gives the reference's limbs.
#pragma once
#include "field_consts.h"
#include "ptx_carry.cuh"

namespace zk {

enum FieldId { BN254_FQ = 0, BN254_FR = 1, BLS12_381_FR = 2 };

template <int FID> struct FieldParams;
#define ZK_DEFINE_FIELD(ID, NAME)                                                                         \
    template <> struct FieldParams<ID> {                                                                  \
        ZK_DEV static constexpr uint32_t p(int i) {                                                       \
            constexpr uint32_t v[8] = {ZKF_##NAME##_P0, ZKF_##NAME##_P1, ZKF_##NAME##_P2, ZKF_##NAME##_P3, \
                                       ZKF_##NAME##_P4, ZKF_##NAME##_P5, ZKF_##NAME##_P6, ZKF_##NAME##_P7}; \
            return v[i];                                                                                  \
        }                                                                                                 \
        ZK_DEV static constexpr uint32_t r2(int i) {                                                      \
            constexpr uint32_t v[8] = {ZKF_##NAME##_R20, ZKF_##NAME##_R21, ZKF_##NAME##_R22, ZKF_##NAME##_R23, \
                                       ZKF_##NAME##_R24, ZKF_##NAME##_R25, ZKF_##NAME##_R26, ZKF_##NAME##_R27}; \
            return v[i];                                                                                  \
        }                                                                                                 \
        /* 2^256 - p: adding q*(2^256-p) mod 2^256 subtracts q*p */                                      \
        ZK_DEV static constexpr uint32_t negp(int i) {                                                    \
            uint32_t borrow = 0, out = 0;                                                                 \
            for (int k = 0; k <= i; ++k) {                                                                \
                uint64_t d = (uint64_t)0 - p(k) - borrow;                                                 \
                out = (uint32_t)d;                                                                        \
                borrow = (uint32_t)(d >> 63);                                                             \
            }                                                                                             \
            return out;                                                                                   \
        }                                                                                                 \
        static constexpr uint32_t inv32 = ZKF_##NAME##_INV32;                                             \
        /* floor(2^296 / p): Barrett constant for inputs < 2^291 */                                      \
        static constexpr uint32_t mu_lo = ZKF_##NAME##_MU296_LO;                                          \
        static constexpr uint32_t mu_hi = ZKF_##NAME##_MU296_HI;                                          \
    };
ZK_DEFINE_FIELD(0, BN254_FQ)
ZK_DEFINE_FIELD(1, BN254_FR)
ZK_DEFINE_FIELD(2, BLS12_381_FR)
#undef ZK_DEFINE_FIELD

struct Fe {
    uint32_t v[8];
};

template <int FID> struct Fp {
    typedef FieldParams<FID> F;

    // ---------------------------------------------------------------- add / sub
    ZK_DEV static void add(Fe& r, const Fe& a, const Fe& b) {
        uint32_t t[8], u[8];
        t[0] = ptx::add_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < 7; ++i) t[i] = ptx::addc_cc(a.v[i], b.v[i]);
        t[7] = ptx::addc(a.v[7], b.v[7]);  // p < 2^255: no carry out of 256 bits
        u[0] = ptx::sub_cc(t[0], F::p(0));
#pragma unroll
        for (int i = 1; i < 8; ++i) u[i] = ptx::subc_cc(t[i], F::p(i));
        uint32_t borrow = ptx::subc(0u, 0u);  // 0xffffffff if t < p
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = borrow ? t[i] : u[i];
    }
    ZK_DEV static void sub(Fe& r, const Fe& a, const Fe& b) {
        uint32_t t[8];
        t[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < 8; ++i) t[i] = ptx::subc_cc(a.v[i], b.v[i]);
        uint32_t mask = ptx::subc(0u, 0u);  // all ones if a < b
        r.v[0] = ptx::add_cc(t[0], F::p(0) & mask);
#pragma unroll
        for (int i = 1; i < 7; ++i) r.v[i] = ptx::addc_cc(t[i], F::p(i) & mask);
        r.v[7] = ptx::addc(t[7], F::p(7) & mask);
    }
    // r = 2a mod p
    ZK_DEV static void dbl(Fe& r, const Fe& a) { add(r, a, a); }
    // r = a - b + p, in (0, 2p) -- NOT canonical; feeds FoldScalar::fold only
    ZK_DEV static void sub_lazy(Fe& r, const Fe& a, const Fe& b) {
        uint32_t t[8];
        t[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < 7; ++i) t[i] = ptx::subc_cc(a.v[i], b.v[i]);
        t[7] = ptx::subc(a.v[7], b.v[7]);
        r.v[0] = ptx::add_cc(t[0], F::p(0));
#pragma unroll
        for (int i = 1; i < 7; ++i) r.v[i] = ptx::addc_cc(t[i], F::p(i));
        r.v[7] = ptx::addc(t[7], F::p(7));
    }
    // if (r >= p) r -= p, for r < 2p given as 8 limbs
    ZK_DEV static void cond_sub_p(uint32_t r[8]) {
        uint32_t u[8];
        u[0] = ptx::sub_cc(r[0], F::p(0));
#pragma unroll
        for (int i = 1; i < 8; ++i) u[i] = ptx::subc_cc(r[i], F::p(i));
        uint32_t borrow = ptx::subc(0u, 0u);
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = borrow ? r[i] : u[i];
    }

    // ---------------------------------------------------------------- row primitives
    // acc[0..7] (+)= x[s], x[s+2], x[s+4], x[s+6] times y in 64-bit slots (0,1),(2,3),(4,5),(6,7).
    // Returns with the carry out of slot 3 in CC.
    template <typename X> ZK_DEV static void row_mad(uint32_t* acc, X x, int s, uint32_t y) {
        ptx::mad_wide_cc(acc[0], acc[1], x(s), y);
#pragma unroll
        for (int k = 1; k < 4; ++k) ptx::madc_wide_cc(acc[2 * k], acc[2 * k + 1], x(s + 2 * k), y);
    }
    // same, but the chain starts with the carry already in CC
    template <typename X> ZK_DEV static void row_madc(uint32_t* acc, X x, int s, uint32_t y) {
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::madc_wide_cc(acc[2 * k], acc[2 * k + 1], x(s + 2 * k), y);
    }
    template <typename X> ZK_DEV static void row_mul(uint32_t* acc, X x, int s, uint32_t y) {
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::mul_wide(acc[2 * k], acc[2 * k + 1], x(s + 2 * k), y);
    }
    struct LimbsOf {
        const uint32_t* v;
        ZK_DEV uint32_t operator()(int i) const { return v[i]; }
    };
    struct LimbsOfP {
        ZK_DEV constexpr uint32_t operator()(int i) const { return F::p(i); }
    };
    struct LimbsOfNegP {
        ZK_DEV constexpr uint32_t operator()(int i) const { return F::negp(i); }
    };

    // ---------------------------------------------------------------- Montgomery product
    // r = a*b/2^256 mod p, canonical.  a, b < p.
    // X holds limb positions (0,1)..(6,7) plus a carry word at 8; Y holds (1,2)..(7,8).  After the
    // m*p row X[0] == 0; dividing by 2^32 turns Y into the next X and X[2..8] into the next Y; the
    // straggler X[1] is added into the new X[0] and its carry enters the new Y chain directly.
    ZK_DEV static void mont_mul(Fe& r, const Fe& a, const Fe& b) {
        uint32_t X[9], Y[8];
        LimbsOf A{a.v};
        LimbsOfP P;
        // row 0
        row_mul(X, A, 0, b.v[0]);
        row_mul(Y, A, 1, b.v[0]);
        uint32_t m = X[0] * F::inv32;
        row_mad(X, P, 0, m);
        X[8] = ptx::addc(0u, 0u);
        row_mad(Y, P, 1, m);  // carry out is provably zero (T < 2^288)
#pragma unroll
        for (int i = 1; i < 8; ++i) {
            uint32_t nX[9], nY[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) nX[k] = Y[k];
#pragma unroll
            for (int k = 0; k < 7; ++k) nY[k] = X[k + 2];
            nY[7] = 0;
            nX[0] = ptx::add_cc(nX[0], X[1]);
            row_madc(nY, A, 1, b.v[i]);
            row_mad(nX, A, 0, b.v[i]);
            nX[8] = ptx::addc(0u, 0u);
            m = nX[0] * F::inv32;
            row_mad(nY, P, 1, m);
            row_mad(nX, P, 0, m);
            nX[8] = ptx::addc(nX[8], 0u);
#pragma unroll
            for (int k = 0; k < 9; ++k) X[k] = nX[k];
#pragma unroll
            for (int k = 0; k < 8; ++k) Y[k] = nY[k];
        }
        // final shift + merge: result = Y + X[1..8]
        uint32_t t[8];
        t[0] = ptx::add_cc(Y[0], X[1]);
#pragma unroll
        for (int k = 1; k < 7; ++k) t[k] = ptx::addc_cc(Y[k], X[k + 1]);
        t[7] = ptx::addc(Y[7], X[8]);
        cond_sub_p(t);
#pragma unroll
        for (int k = 0; k < 8; ++k) r.v[k] = t[k];
    }

    // ---------------------------------------------------------------- unreduced product
    // t[0..15] = a*b (512 bits).
    ZK_DEV static void mul_wide(uint32_t t[16], const Fe& a, const Fe& b) {
        uint32_t E[18], O[18];
#pragma unroll
        for (int k = 0; k < 18; ++k) E[k] = O[k] = 0;
        LimbsOf A{a.v};
        row_mul(E, A, 0, b.v[0]);
        row_mul(O, A, 1, b.v[0]);
#pragma unroll
        for (int i = 1; i < 8; ++i) {
            if (i & 1) {
                // odd row: even limbs of a land on odd positions (O[i-1 ..]), odd limbs on even (E[i+1 ..])
                row_mad(O + (i - 1), A, 0, b.v[i]);
                O[i + 7] = ptx::addc(O[i + 7], 0u);
                row_mad(E + (i + 1), A, 1, b.v[i]);
                E[i + 9] = ptx::addc(E[i + 9], 0u);
            } else {
                row_mad(E + i, A, 0, b.v[i]);
                E[i + 8] = ptx::addc(E[i + 8], 0u);
                row_mad(O + i, A, 1, b.v[i]);
                O[i + 8] = ptx::addc(O[i + 8], 0u);
            }
        }
        t[0] = E[0];
        t[1] = ptx::add_cc(E[1], O[0]);
#pragma unroll
        for (int k = 2; k < 15; ++k) t[k] = ptx::addc_cc(E[k], O[k - 1]);
        t[15] = ptx::addc(E[15], O[14]);
    }
    // acc (17 limbs) += a*b
    ZK_DEV static void mul_acc(uint32_t acc[17], const Fe& a, const Fe& b) {
        uint32_t t[16];
        mul_wide(t, a, b);
        acc[0] = ptx::add_cc(acc[0], t[0]);
#pragma unroll
        for (int k = 1; k < 16; ++k) acc[k] = ptx::addc_cc(acc[k], t[k]);
        acc[16] = ptx::addc(acc[16], 0u);
    }

    // ---------------------------------------------------------------- Barrett step
    // s[0..9] < 2^291  ->  r = s mod p (canonical).
    //   x = s >> 232 (< 2^59);  q = (x * floor(2^296/p)) >> 64  in {floor(s/p)-1, floor(s/p)};
    //   r = (s + q*(2^256-p)) mod 2^256, then one conditional subtraction.
    ZK_DEV static void barrett(uint32_t r[8], const uint32_t s[10]) {
        uint32_t x_lo = (s[7] >> 8) | (s[8] << 24);
        uint32_t x_hi = (s[8] >> 8) | (s[9] << 24);
        uint64_t w = (uint64_t)x_lo * F::mu_lo;
        uint64_t t1 = (uint64_t)x_lo * F::mu_hi + (w >> 32);
        uint64_t t2 = (uint64_t)x_hi * F::mu_lo + (uint32_t)t1;
        uint64_t q = (uint64_t)x_hi * F::mu_hi + (t1 >> 32) + (t2 >> 32);
        uint32_t q_lo = (uint32_t)q, q_hi = (uint32_t)(q >> 32);
        // E covers positions 0..7 (+ scratch), O positions 1..8 (+ scratch); only positions < 8 matter.
        uint32_t E[10], O[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) E[k] = s[k];
        E[8] = E[9] = 0;
        LimbsOfNegP N;
        row_mad(E, N, 0, q_lo);           // q_lo * negp[even] -> positions 0,2,4,6
        row_mul(O, N, 1, q_lo);           // q_lo * negp[odd]  -> positions 1,3,5,7
        row_mad(O, N, 0, q_hi);           // q_hi * negp[even] -> positions 1,3,5,7
        row_mad(E + 2, N, 1, q_hi);       // q_hi * negp[odd]  -> positions 2,4,6,(8: discarded)
        r[0] = E[0];
        r[1] = ptx::add_cc(E[1], O[0]);
#pragma unroll
        for (int k = 2; k < 7; ++k) r[k] = ptx::addc_cc(E[k], O[k - 1]);
        r[7] = ptx::addc(E[7], O[6]);
        cond_sub_p(r);
    }

    // ---------------------------------------------------------------- lazy accumulators
    // t = (a + M p) / 2^256 for the M in [0, 2^256) that makes the division exact: a * 2^-256 mod p
    // up to one multiple of p (t <= p).  a is ANY 256-bit integer.  Same window scheme as mont_mul with
    // the a*b rows left out: 8 rows of two independent four-slot chains.
    ZK_DEV static void redc256(uint32_t t[8], const uint32_t a[8]) {
        uint32_t X[9], Y[8];
        LimbsOfP P;
#pragma unroll
        for (int k = 0; k < 8; ++k) { X[k] = a[k]; Y[k] = 0; }
        uint32_t m = X[0] * F::inv32;
        row_mad(X, P, 0, m);
        X[8] = ptx::addc(0u, 0u);
        row_mad(Y, P, 1, m);   // Y was zero: no carry out
#pragma unroll
        for (int i = 1; i < 8; ++i) {
            uint32_t nX[9], nY[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) nX[k] = Y[k];
#pragma unroll
            for (int k = 0; k < 7; ++k) nY[k] = X[k + 2];
            nY[7] = 0;
            nX[0] = ptx::add_cc(nX[0], X[1]);   // the straggler word; its carry enters the Y chain
            m = nX[0] * F::inv32;               // mul.lo leaves the carry flag alone
            row_madc(nY, P, 1, m);
            row_mad(nX, P, 0, m);
            nX[8] = ptx::addc(0u, 0u);
#pragma unroll
            for (int k = 0; k < 9; ++k) X[k] = nX[k];
#pragma unroll
            for (int k = 0; k < 8; ++k) Y[k] = nY[k];
        }
        t[0] = ptx::add_cc(Y[0], X[1]);
#pragma unroll
        for (int k = 1; k < 7; ++k) t[k] = ptx::addc_cc(Y[k], X[k + 1]);
        t[7] = ptx::addc(Y[7], X[8]);
    }
    // Montgomery-reduce a 17-limb accumulator A (sum of < 2^32 unreduced products of canonical
    // Montgomery elements): returns A / 2^256 mod p, canonical.
    //   A = A_lo + 2^256 A_hi  =>  A / 2^256 == redc256(A_lo) + A_hi  (mod p),  < p + 1 + 2^288 < 2^291.
    ZK_DEV static void redc_wide(Fe& r, const uint32_t acc[17]) {
        uint32_t t[8], s[10];
        redc256(t, acc);
        s[0] = ptx::add_cc(t[0], acc[8]);
#pragma unroll
        for (int k = 1; k < 8; ++k) s[k] = ptx::addc_cc(t[k], acc[k + 8]);
        s[8] = ptx::addc_cc(acc[16], 0u);
        s[9] = ptx::addc(0u, 0u);
        barrett(r.v, s);
    }
    // 9-limb sum of canonical elements (plain-sumcheck round sums) -> canonical
    ZK_DEV static void reduce9(Fe& r, const uint32_t acc[9]) {
        uint32_t s[10];
#pragma unroll
        for (int k = 0; k < 9; ++k) s[k] = acc[k];
        s[9] = 0;
        barrett(r.v, s);
    }
    ZK_DEV static void acc9_add(uint32_t acc[9], const Fe& a) {
        acc[0] = ptx::add_cc(acc[0], a.v[0]);
#pragma unroll
        for (int k = 1; k < 8; ++k) acc[k] = ptx::addc_cc(acc[k], a.v[k]);
        acc[8] = ptx::addc(acc[8], 0u);
    }
};

// ---------------------------------------------------------------- fold by a per-round scalar
// The host precomputes, from the PLAIN (non-Montgomery) challenge r, the eight multiples
// tab[i] = r * 2^(32 i) mod p.  For a Montgomery element d, sum_i d_i * tab[i] == d * r (mod p),
// which is the Montgomery form of the product.  fold() returns lo + r*(hi - lo), canonical.
struct FoldTable {
    uint32_t w[8][8];
};

template <int FID> struct FoldScalar {
    typedef FieldParams<FID> F;
    typedef Fp<FID> P;
    struct Row {
        const uint32_t* v;
        ZK_DEV uint32_t operator()(int i) const { return v[i]; }
    };
    ZK_DEV static void fold(Fe& out, const Fe& lo, const Fe& hi, const FoldTable& tab) {
        Fe d;
        P::sub_lazy(d, hi, lo);  // hi - lo + p in (0, 2p)
        uint32_t E[10], O[9];
#pragma unroll
        for (int k = 0; k < 8; ++k) E[k] = lo.v[k];
        E[8] = E[9] = 0;
        {
            Row R{tab.w[0]};
            P::row_mad(E, R, 0, d.v[0]);
            E[8] = ptx::addc(E[8], 0u);
            P::row_mul(O, R, 1, d.v[0]);
            O[8] = 0;
        }
#pragma unroll
        for (int i = 1; i < 8; ++i) {
            Row R{tab.w[i]};
            P::row_mad(E, R, 0, d.v[i]);
            E[8] = ptx::addc(E[8], 0u);
            P::row_mad(O, R, 1, d.v[i]);
            O[8] = ptx::addc(O[8], 0u);
        }
        // S = E + O*2^32 < p*(1 + 2^35) < 2^291
        uint32_t s[10];
        s[0] = E[0];
        s[1] = ptx::add_cc(E[1], O[0]);
#pragma unroll
        for (int k = 2; k < 9; ++k) s[k] = ptx::addc_cc(E[k], O[k - 1]);
        s[9] = ptx::addc(E[9], O[8]);
        P::barrett(out.v, s);
    }
};

}  // namespace zk
