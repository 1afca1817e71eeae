// experiments/fp_karatsuba.cuh -- one-level Karatsuba for the unreduced 256 x 256-bit product: three
// 4 x 4-limb products (48 IMAD.WIDE) instead of the 64 of the schoolbook rows.  Bit-identical to Fp::mul_wide (emulated
// against the oracle, all GPU parity tests green with it) and SLOWER on B200: record in profiles/r02/ab_karatsuba.md.
// Not part of libzkb200.  To repeat the A/B: paste the three routines back into Fp<FID> (csrc/fp.cuh) in front of mul_acc and
// call mul_wide_karatsuba there.
#pragma once
#include "fp.cuh"

namespace zk {

template <int FID> struct FpKaratsuba {
    // 4 x 4-limb product, 8 limbs out: 4 plain + 12 carry-chained IMAD.WIDE in the same even/odd slot scheme as mul_wide.
    // E holds positions 0..7, O positions 1..8 (O[k] = position k + 1); neither partial sum can carry out of its top limb.
    template <typename X, typename Y> ZK_DEV static void mul4(uint32_t t[8], X x, Y y) {
        uint32_t E[8], O[8];
#pragma unroll
        for (int k = 4; k < 8; ++k) E[k] = O[k] = 0;
        ptx::mul_wide(E[0], E[1], x(0), y(0));
        ptx::mul_wide(E[2], E[3], x(2), y(0));
        ptx::mul_wide(O[0], O[1], x(1), y(0));
        ptx::mul_wide(O[2], O[3], x(3), y(0));
        // row 1: even limbs of x land on odd positions (O[0..3]), odd limbs on even ones (E[2..5])
        ptx::mad_wide_cc(O[0], O[1], x(0), y(1));
        ptx::madc_wide_cc(O[2], O[3], x(2), y(1));
        O[4] = ptx::addc(O[4], 0u);
        ptx::mad_wide_cc(E[2], E[3], x(1), y(1));
        ptx::madc_wide_cc(E[4], E[5], x(3), y(1));
        E[6] = ptx::addc(E[6], 0u);
        // row 2
        ptx::mad_wide_cc(E[2], E[3], x(0), y(2));
        ptx::madc_wide_cc(E[4], E[5], x(2), y(2));
        E[6] = ptx::addc(E[6], 0u);
        ptx::mad_wide_cc(O[2], O[3], x(1), y(2));
        ptx::madc_wide_cc(O[4], O[5], x(3), y(2));
        O[6] = ptx::addc(O[6], 0u);
        // row 3
        ptx::mad_wide_cc(O[2], O[3], x(0), y(3));
        ptx::madc_wide_cc(O[4], O[5], x(2), y(3));
        O[6] = ptx::addc(O[6], 0u);
        ptx::mad_wide_cc(E[4], E[5], x(1), y(3));
        ptx::madc_wide_cc(E[6], E[7], x(3), y(3));
        t[0] = E[0];
        t[1] = ptx::add_cc(E[1], O[0]);
#pragma unroll
        for (int k = 2; k < 7; ++k) t[k] = ptx::addc_cc(E[k], O[k - 1]);
        t[7] = ptx::addc(E[7], O[6]);
    }
    struct Limbs4 {
        const uint32_t* v;
        ZK_DEV uint32_t operator()(int i) const { return v[i]; }
    };
    // |hi - lo| of two 4-limb numbers; returns the sign mask (all ones iff hi < lo)
    ZK_DEV static uint32_t abs_diff4(uint32_t d[4], const uint32_t* hi, const uint32_t* lo) {
        d[0] = ptx::sub_cc(hi[0], lo[0]);
        d[1] = ptx::subc_cc(hi[1], lo[1]);
        d[2] = ptx::subc_cc(hi[2], lo[2]);
        d[3] = ptx::subc_cc(hi[3], lo[3]);
        const uint32_t m = ptx::subc(0u, 0u);
        d[0] = ptx::sub_cc(d[0] ^ m, m);       // two's complement negate where the difference was negative
        d[1] = ptx::subc_cc(d[1] ^ m, m);
        d[2] = ptx::subc_cc(d[2] ^ m, m);
        d[3] = ptx::subc(d[3] ^ m, m);
        return m;
    }
    // t[0..15] = a*b with three 4x4 products (48 IMAD.WIDE instead of 64):
    //   a = a0 + 2^128 a1, b = b0 + 2^128 b1, z0 = a0 b0, z2 = a1 b1, zd = |a1 - a0| |b1 - b0|,
    //   a b = z0 + 2^256 z2 + 2^128 (z0 + z2 -+ zd)      (- when the two differences have the same sign)
    ZK_DEV static void mul_wide_karatsuba(uint32_t t[16], const Fe& a, const Fe& b) {
        uint32_t zd[8], da[4], db[4];
        mul4(t, Limbs4{a.v}, Limbs4{b.v});
        mul4(t + 8, Limbs4{a.v + 4}, Limbs4{b.v + 4});
        const uint32_t sa = abs_diff4(da, a.v + 4, a.v), sb = abs_diff4(db, b.v + 4, b.v);
        mul4(zd, Limbs4{da}, Limbs4{db});
        const uint32_t n = ~(sa ^ sb);          // all ones: subtract zd; zero: add it
        uint32_t m[9];                          // z0 + z2, then -+ zd: the middle term, 0 <= m < 2^257
        m[0] = ptx::add_cc(t[0], t[8]);
#pragma unroll
        for (int k = 1; k < 8; ++k) m[k] = ptx::addc_cc(t[k], t[k + 8]);
        m[8] = ptx::addc(0u, 0u);
        (void)ptx::add_cc(n, 1u);               // carry flag := (n == all ones): the +1 of the two's complement
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = ptx::addc_cc(m[k], zd[k] ^ n);
        m[8] = ptx::addc(m[8], n);
        t[4] = ptx::add_cc(t[4], m[0]);
#pragma unroll
        for (int k = 1; k < 9; ++k) t[4 + k] = ptx::addc_cc(t[4 + k], m[k]);
        t[13] = ptx::addc_cc(t[13], 0u);
        t[14] = ptx::addc_cc(t[14], 0u);
        t[15] = ptx::addc(t[15], 0u);
    }
};

}  // namespace zk
