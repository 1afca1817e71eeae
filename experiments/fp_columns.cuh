// experiments/fp_columns.cuh -- carry-chain-free product formulations that were measured in round 1 and LOST to the
// chained even/odd rows of csrc/fp.cuh (A/B record: profiles/r01b/ab_column_arithmetic/, profiles/r01b/probe_products.json,
// DESIGN.md section 3).  Not part of libzkb200: nothing under zk_cryptography_research_implementations_b200/ includes this
// file.  Kept so the measurements can be repeated; `experiments/README.md` says how.
#pragma once
#include "fp.cuh"   // -I zk_cryptography_research_implementations_b200/csrc

namespace zk {
namespace ptx {
#if defined(ZK_HOST_EMU)
// 96-bit slot: (lo, hi) += a * b, the carry out of the slot lands in `top` -- no carry enters, no chain leaves
inline void mad_wide_top(uint32_t& lo, uint32_t& hi, uint32_t& top, uint32_t a, uint32_t b) { mad_wide_cc(lo, hi, a, b); top = addc(top, 0u); }
// c += a * b as a 64-bit integer, no flags (the caller guarantees no overflow)
inline void mad_wide_noflags(uint64_t& c, uint32_t a, uint32_t b) { c += (uint64_t)a * b; }
// bits [s, s + 32) of the 64-bit value hi:lo, s in [0, 31]
inline uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) { return s ? (lo >> s) | (hi << (32 - s)) : lo; }
#else
// 96-bit slot: (lo, hi) += a * b with the carry out of the slot caught in `top`: IMAD.WIDE.U32 with a carry-OUT
// predicate + an IADD3.X on the alu pipe (measured: the carry-out form issues at about the rate of the carry-in form
// IMAD.WIDE.U32.X, half the rate of a plain IMAD.WIDE.U32).
// (volatile like every statement that touches the condition code: the front end must not move it into another chain;
// ptxas turns the flag into predicates and then schedules the independent slots freely)
// c += a * b as a 64-bit integer: plain IMAD.WIDE.U32, no carry flag in or out -- the only full-rate form of the
// instruction on B200 (59 lanes / clk / SM against 27 with a flag).  The caller guarantees the sum fits 64 bits.
// (volatile: keeps the front end from splitting it into a multiply and 64-bit adds or hoisting it; ptxas still schedules)
ZK_DEV void mad_wide_noflags(uint64_t& c, uint32_t a, uint32_t b) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c) : "r"(a), "r"(b)); }
ZK_DEV uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }
ZK_DEV void mad_wide_top(uint32_t& lo, uint32_t& hi, uint32_t& top, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;" : "+r"(lo), "+r"(hi), "+r"(top) : "r"(a), "r"(b));
}
#endif
}  // namespace ptx

// The experiment routines keep their round-1 bodies; they used to be static members of Fp<FID>.
template <int FID> struct FpColumns {
    typedef FieldParams<FID> F;
    // ---------------------------------------------------------------- unreduced products without carry chains
    // Column accumulator: slot k collects every limb product a_i b_j with i + j == k as a 96-bit integer
    // (lo, hi, top) at weight 2^(32 k).  A product costs one carry-OUT-only IMAD.WIDE.U32 plus half an IADD3.X (ptxas
    // feeds two carry predicates into one IADD3.X), and the 64 products of a multiplication are independent of each
    // other -- no IMAD.WIDE.U32.X.  Up to 2^29 products per slot fit (8 x 2^29 x 2^64 < 2^96).  Measured on B200
    // (zk_arith_probe kind 7): 122 G products/s against 102 G/s for the chained mul_acc -- the carry-OUT form is not
    // the full-rate instruction either -- and 45 registers per accumulator instead of 17; the round kernels that
    // tried it (ZK_ROUND0_COLS) lost more to the halved occupancy than they gained.  Kept as a tested experiment.
    struct ColAcc {
        uint32_t lo[15], hi[15], top[15];
    };
    ZK_DEV static void cols_init(ColAcc& c) {
#pragma unroll
        for (int k = 0; k < 15; ++k) c.lo[k] = c.hi[k] = c.top[k] = 0;
    }
    ZK_DEV static void mul_acc_cols(ColAcc& c, const Fe& a, const Fe& b) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) ptx::mad_wide_top(c.lo[i + j], c.hi[i + j], c.top[i + j], a.v[i], b.v[j]);
    }
    // the slots as one 17-limb integer (the sum of < 2^32 products of canonical elements fits: < 2^542)
    ZK_DEV static void cols_to_limbs(uint32_t t[17], const ColAcc& c) {
        // limb p collects lo[p] + hi[p-1] + top[p-2]: three carry chains added one after the other
        t[0] = c.lo[0];
#pragma unroll
        for (int p = 1; p < 15; ++p) t[p] = c.lo[p];
        t[15] = t[16] = 0;
        t[1] = ptx::add_cc(t[1], c.hi[0]);
#pragma unroll
        for (int p = 2; p < 16; ++p) t[p] = ptx::addc_cc(t[p], c.hi[p - 1]);
        t[16] = ptx::addc(t[16], 0u);
        t[2] = ptx::add_cc(t[2], c.top[0]);
#pragma unroll
        for (int p = 3; p < 16; ++p) t[p] = ptx::addc_cc(t[p], c.top[p - 2]);
        t[16] = ptx::addc(t[16], c.top[14]);
    }

    // ---------------------------------------------------------------- flag-free products in radix 2^29 (EXPERIMENT)
    // Nine 29-bit digits per operand: a digit product is < 2^58 and a column collects at most 9 of them per
    // multiplication, so SIX multiplications accumulate in 64-bit columns with the flag-free IMAD.WIDE.U32 -- the one
    // form of the instruction that issues at full rate -- before the columns are carried out into the 32-bit-limb
    // accumulator.  81 multiplies instead of 64, none of them carry-chained.
    struct Digits29 {
        uint32_t v[9];
    };
    struct Cols29 {
        uint64_t c[17];   // sum_k c[k] 2^(29 k)
    };
    static constexpr int kCols29Budget = 6;   // multiplications per flush: 6 x 9 x 2^58 < 2^64
    ZK_DEV static void to_digits29(Digits29& o, const Fe& a) {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int bit = 29 * k, w = bit >> 5, sh = bit & 31;
            const uint32_t lo = a.v[w], hi = (w + 1 < 8) ? a.v[w + 1] : 0u;
            o.v[k] = ptx::funnel_r(lo, hi, sh) & 0x1fffffffu;
        }
    }
    ZK_DEV static void cols29_init(Cols29& c) {
#pragma unroll
        for (int k = 0; k < 17; ++k) c.c[k] = 0;
    }
    ZK_DEV static void mul_cols29(Cols29& c, const Digits29& a, const Digits29& b) {
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int j = 0; j < 9; ++j) ptx::mad_wide_noflags(c.c[i + j], a.v[i], b.v[j]);
    }
    // acc (17 x 32-bit limbs) += sum_k c[k] 2^(29 k); c = 0
    ZK_DEV static void cols29_flush(uint32_t acc[17], Cols29& c) {
        // radix-2^29 carry propagation: 18 exact digits + what is left above
        uint32_t dg[19];
        uint64_t carry = 0;
#pragma unroll
        for (int k = 0; k < 17; ++k) {
            const uint64_t t = c.c[k] + carry;
            dg[k] = (uint32_t)t & 0x1fffffffu;
            carry = t >> 29;
            c.c[k] = 0;
        }
        dg[17] = (uint32_t)carry & 0x1fffffffu;
        dg[18] = (uint32_t)(carry >> 29);
        // repack into 32-bit limbs: limb p = bits [32 p, 32 p + 32) of sum_k dg[k] 2^(29 k)
        uint32_t t[17];
#pragma unroll
        for (int p = 0; p < 17; ++p) {
            const int bit = 32 * p, k = bit / 29, off = bit - 29 * k;   // limb p starts `off` bits into digit k
            uint32_t v = dg[k] >> off;
            if (k + 1 < 19) v |= dg[k + 1] << (29 - off);
            if (29 - off + 29 < 32 && k + 2 < 19) v |= dg[k + 2] << (58 - off);
            t[p] = v;
        }
        acc[0] = ptx::add_cc(acc[0], t[0]);
#pragma unroll
        for (int p = 1; p < 16; ++p) acc[p] = ptx::addc_cc(acc[p], t[p]);
        acc[16] = ptx::addc(acc[16], t[16]);
    }


    // ---------------------------------------------------------------- column form of FoldScalar::fold
    // Column form (EXPERIMENT, off by default): limb j of every table row lands in slot j, so the 64 limb products fall
    // into 8 independent 96-bit slots -- 64 carry-OUT-only IMAD.WIDE.U32 and ~33 IADD3.X that collect two carries each,
    // instead of 19 + 45 carry-chained IMAD.WIDE.U32.X.  Same integer S, same Barrett step, bit-identical results;
    // on B200 the carry-out form issues no faster than the carry-in form (fold probe 86 vs 91 G/s).
    ZK_DEV static void fold_cols(Fe& out, const Fe& lo, const Fe& hi, const FoldTable& tab) {
        Fe d;
        Fp<FID>::sub_lazy(d, hi, lo);  // hi - lo + p in (0, 2p)
        uint32_t cl[8], ch[8], ct[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            cl[j] = lo.v[j];
            ch[j] = ct[j] = 0;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) ptx::mad_wide_top(cl[j], ch[j], ct[j], d.v[i], tab.w[i][j]);
        // S = sum_j (cl[j] + ch[j] 2^32 + ct[j] 2^64) 2^(32 j) < p (1 + 2^35) < 2^291
        uint32_t s[10];
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] = cl[k];
        s[8] = s[9] = 0;
        s[1] = ptx::add_cc(s[1], ch[0]);
#pragma unroll
        for (int k = 2; k < 9; ++k) s[k] = ptx::addc_cc(s[k], ch[k - 1]);
        s[9] = ptx::addc(s[9], 0u);
        s[2] = ptx::add_cc(s[2], ct[0]);
#pragma unroll
        for (int k = 3; k < 9; ++k) s[k] = ptx::addc_cc(s[k], ct[k - 2]);
        s[9] = ptx::addc(s[9], ct[7]);
        Fp<FID>::barrett(out.v, s);
    }
};

}  // namespace zk
