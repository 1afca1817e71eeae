// fq381.cuh -- the BLS12-381 base field Fq (381 bits) for sm_100a: 12 x 32-bit limbs in registers, Montgomery form with
// R = 2^384 -- the in-memory layout of arkworks' `Fp<MontBackend<_,6>,6>` (6 x u64 little-endian == 12 x u32), which is
// what `P::G1` coordinates are made of (multilinear_kzg/src/multilinear_kzg.rs:25-46, trusted_setup.rs:54-63).
//
// Same construction as fp.cuh, two limbs wider: a Montgomery product is 12 rounds of two independent six-slot
// IMAD.WIDE.U32(.X) chains (even limbs / odd limbs of the multiplicand, then of q), 288 wide multiply-adds in all.
// Every routine returns the canonical representative.
#pragma once
#include "curve_consts.h"
#include "ptx_carry.cuh"

namespace zk {

struct Fq {
    uint32_t v[12];
};

struct Fq381 {
    ZK_DEV static constexpr uint32_t q(int i) {
        constexpr uint32_t v[12] = ZKC_Q_32;
        return v[i];
    }
    ZK_DEV static constexpr uint32_t one(int i) {   // R mod q
        constexpr uint32_t v[12] = ZKC_R_32;
        return v[i];
    }
    static constexpr uint32_t inv32 = ZKC_INV32;

    ZK_DEV static Fq zero() {
        Fq r;
#pragma unroll
        for (int i = 0; i < 12; ++i) r.v[i] = 0;
        return r;
    }
    ZK_DEV static Fq mont_one() {
        Fq r;
#pragma unroll
        for (int i = 0; i < 12; ++i) r.v[i] = one(i);
        return r;
    }
    ZK_DEV static bool is_zero(const Fq& a) {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 12; ++i) o |= a.v[i];
        return o == 0;
    }
    ZK_DEV static bool eq(const Fq& a, const Fq& b) {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 12; ++i) o |= a.v[i] ^ b.v[i];
        return o == 0;
    }
    // if (r >= q) r -= q, for r < 2q
    ZK_DEV static void cond_sub_q(uint32_t r[12]) {
        uint32_t u[12];
        u[0] = ptx::sub_cc(r[0], q(0));
#pragma unroll
        for (int i = 1; i < 12; ++i) u[i] = ptx::subc_cc(r[i], q(i));
        uint32_t borrow = ptx::subc(0u, 0u);
#pragma unroll
        for (int i = 0; i < 12; ++i) r[i] = borrow ? r[i] : u[i];
    }
    ZK_DEV static void add(Fq& r, const Fq& a, const Fq& b) {
        uint32_t t[12];
        t[0] = ptx::add_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < 11; ++i) t[i] = ptx::addc_cc(a.v[i], b.v[i]);
        t[11] = ptx::addc(a.v[11], b.v[11]);   // 2q < 2^382: no carry out
        cond_sub_q(t);
#pragma unroll
        for (int i = 0; i < 12; ++i) r.v[i] = t[i];
    }
    ZK_DEV static void sub(Fq& r, const Fq& a, const Fq& b) {
        uint32_t t[12];
        t[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < 12; ++i) t[i] = ptx::subc_cc(a.v[i], b.v[i]);
        uint32_t mask = ptx::subc(0u, 0u);      // all ones if a < b
        r.v[0] = ptx::add_cc(t[0], q(0) & mask);
#pragma unroll
        for (int i = 1; i < 11; ++i) r.v[i] = ptx::addc_cc(t[i], q(i) & mask);
        r.v[11] = ptx::addc(t[11], q(11) & mask);
    }
    ZK_DEV static void dbl(Fq& r, const Fq& a) { add(r, a, a); }
    ZK_DEV static void neg(Fq& r, const Fq& a) {
        Fq z = zero();
        sub(r, z, a);
    }

    // ---- row primitives: six 64-bit slots (0,1)..(10,11) of acc (+)= x(s), x(s+2), .., x(s+10) times y
    template <typename X> ZK_DEV static void row_mul(uint32_t* acc, X x, int s, uint32_t y) {
#pragma unroll
        for (int k = 0; k < 6; ++k) ptx::mul_wide(acc[2 * k], acc[2 * k + 1], x(s + 2 * k), y);
    }
    template <typename X> ZK_DEV static void row_mad(uint32_t* acc, X x, int s, uint32_t y) {
        ptx::mad_wide_cc(acc[0], acc[1], x(s), y);
#pragma unroll
        for (int k = 1; k < 6; ++k) ptx::madc_wide_cc(acc[2 * k], acc[2 * k + 1], x(s + 2 * k), y);
    }
    template <typename X> ZK_DEV static void row_madc(uint32_t* acc, X x, int s, uint32_t y) {
#pragma unroll
        for (int k = 0; k < 6; ++k) ptx::madc_wide_cc(acc[2 * k], acc[2 * k + 1], x(s + 2 * k), y);
    }
    struct LimbsOf {
        const uint32_t* v;
        ZK_DEV uint32_t operator()(int i) const { return v[i]; }
    };
    struct LimbsOfQ {
        ZK_DEV constexpr uint32_t operator()(int i) const { return q(i); }
    };

    // r = a*b / 2^384 mod q, canonical; a, b < q.  X holds limb positions (0,1)..(10,11) plus a carry word at 12, Y holds
    // (1,2)..(11,12); after the m*q row X[0] == 0, the division by 2^32 turns Y into the next X and X[2..12] into the
    // next Y, and the straggler X[1] enters the new X[0] with its carry going straight into the new Y chain.
    ZK_DEV static void mul(Fq& r, const Fq& a, const Fq& b) {
        uint32_t X[13], Y[12];
        LimbsOf A{a.v};
        LimbsOfQ Qm;
        row_mul(X, A, 0, b.v[0]);
        row_mul(Y, A, 1, b.v[0]);
        uint32_t m = X[0] * inv32;
        row_mad(X, Qm, 0, m);
        X[12] = ptx::addc(0u, 0u);
        row_mad(Y, Qm, 1, m);   // top slot: a[11]*b0 + q[11]*m < 2^62: no carry out
#pragma unroll
        for (int i = 1; i < 12; ++i) {
            uint32_t nX[13], nY[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) nX[k] = Y[k];
#pragma unroll
            for (int k = 0; k < 11; ++k) nY[k] = X[k + 2];
            nY[11] = 0;
            nX[0] = ptx::add_cc(nX[0], X[1]);
            row_madc(nY, A, 1, b.v[i]);
            row_mad(nX, A, 0, b.v[i]);
            nX[12] = ptx::addc(0u, 0u);
            m = nX[0] * inv32;
            row_mad(nY, Qm, 1, m);
            row_mad(nX, Qm, 0, m);
            nX[12] = ptx::addc(nX[12], 0u);
#pragma unroll
            for (int k = 0; k < 13; ++k) X[k] = nX[k];
#pragma unroll
            for (int k = 0; k < 12; ++k) Y[k] = nY[k];
        }
        uint32_t t[12];
        t[0] = ptx::add_cc(Y[0], X[1]);
#pragma unroll
        for (int k = 1; k < 11; ++k) t[k] = ptx::addc_cc(Y[k], X[k + 1]);
        t[11] = ptx::addc(Y[11], X[12]);
        cond_sub_q(t);
#pragma unroll
        for (int k = 0; k < 12; ++k) r.v[k] = t[k];
    }
    ZK_DEV static void sqr(Fq& r, const Fq& a) { mul(r, a, a); }
};

}  // namespace zk
