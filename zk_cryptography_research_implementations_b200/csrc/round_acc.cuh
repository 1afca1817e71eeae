// round_acc.cuh -- per-thread accumulation of the round polynomial and its exact grid-wide finish.
//
// A thread accumulates the UNREDUCED contributions of its pairs to s(X), X = 0..D, as multi-limb integers
// (9 limbs for the plain sums, 17 for sums of products).  Nothing is reduced per thread: the limbs of all
// threads are summed column by column as exact 64-bit integers (warp REDUX, shared memory, one RED per column
// per block -- kernels.cuh), the column totals are carry-propagated once, and ONE Montgomery reduction per
// evaluation yields the canonical element the reference computes (sumcheck_gkr_protocol.rs:127-137,
// prover.rs:74-89).  Integer addition is associative, so the result does not depend on the grid shape.
//
// Compiles for the host with -DZK_HOST_EMU (tests/host_emu) so the finish is unit-tested without a GPU.
#pragma once
#include "fp.cuh"

namespace zk {

constexpr int kMaxTables = 8;   // P * D (+ linear tables)
constexpr int kMaxEvals = 5;    // D + 1
constexpr int kMaxCols = kMaxEvals * 17 + 18;   // RoundAcc::NC upper bound

struct TablePtrs {
    Fe* t[kMaxTables];
};

// Given the (lo, hi) values of all T = P*D tables at one pair index, accumulate the unreduced
// contributions to s(X), X = 0..D:  s(X) += sum_p prod_d (lo + X (hi - lo)).
// (sumcheck_gkr_protocol.rs:127-137 evaluates the same sums with D+1 full folds of every table.)
// D == 1 (plain sumcheck, prover.rs:74-89) accumulates 9-limb sums; D >= 2 accumulates 17-limb
// unreduced products.  SKIP1 leaves s(1) out (the prover derives it from the running claim).
// NLIN extra tables enter the sum LINEARLY (a product with the all-ones table, which is then neither stored, folded
// nor multiplied): s(X) += sum_l (lo_l + X (hi_l - lo_l)).  The sparse GKR layer prover's phases have exactly that
// shape, h1*W + h2*1 (gkr_wide.cu).  Only the two half sums are accumulated; finalize() spreads them over the s(X).
template <int FID, int P, int D, bool SKIP1, int NLIN = 0> struct RoundAcc {
    static constexpr int NE = D + 1;
    static constexpr int W = (D == 1) ? 9 : 17;
    static constexpr int T = P * D + NLIN;
    static constexpr int NC = NE * W + (NLIN > 0 ? 18 : 0);   // 32-bit columns a thread contributes
    uint32_t acc[NE][W];
    uint32_t lin[NLIN > 0 ? 2 : 1][9];
    ZK_DEV void init() {
#pragma unroll
        for (int e = 0; e < NE; ++e)
#pragma unroll
            for (int k = 0; k < W; ++k) acc[e][k] = 0;
#pragma unroll
        for (int e = 0; e < (NLIN > 0 ? 2 : 1); ++e)
#pragma unroll
            for (int k = 0; k < 9; ++k) lin[e][k] = 0;
    }
    ZK_DEV void add_point(int e, const Fe (&v)[T]) {
        if (D == 1) {
#pragma unroll
            for (int p = 0; p < P; ++p) Fp<FID>::acc9_add(acc[e], v[p]);
        } else {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                Fe prod = v[p * D];
#pragma unroll
                for (int d = 1; d < D - 1; ++d) Fp<FID>::mont_mul(prod, prod, v[p * D + d]);
                Fp<FID>::mul_acc(acc[e], prod, v[p * D + D - 1]);
            }
        }
    }
    ZK_DEV void add_pair(const Fe (&lo)[T], const Fe (&hi)[T]) {
        add_point(0, lo);
        if (!SKIP1) add_point(1, hi);
#pragma unroll
        for (int l = 0; l < NLIN; ++l) {
            Fp<FID>::acc9_add(lin[0], lo[P * D + l]);
            Fp<FID>::acc9_add(lin[1], hi[P * D + l]);
        }
        if (D >= 2) {
            Fe cur[T], diff[T];
#pragma unroll
            for (int t = 0; t < P * D; ++t) {
                Fp<FID>::sub(diff[t], hi[t], lo[t]);
                Fp<FID>::add(cur[t], hi[t], diff[t]);  // value at X = 2
            }
            add_point(2, cur);
#pragma unroll
            for (int x = 3; x <= D; ++x) {
#pragma unroll
                for (int t = 0; t < P * D; ++t) Fp<FID>::add(cur[t], cur[t], diff[t]);
                add_point(x, cur);
            }
        }
    }
    // column c of this thread: acc[e][k] at c = e*W + k, then the two linear half sums
    ZK_DEV void columns(uint32_t (&col)[NC]) const {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
#pragma unroll
            for (int k = 0; k < W; ++k) col[e * W + k] = acc[e][k];
        }
        if (NLIN > 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int k = 0; k < 9; ++k) col[NE * W + h * 9 + k] = lin[h][k];
        }
    }
    // n 64-bit column totals -> n 32-bit limbs (the total fits: see the bounds above)
    template <int N> ZK_DEV static void carry_columns(uint32_t (&limbs)[N], const unsigned long long* tot) {
        unsigned long long c = 0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            c += tot[k];
            limbs[k] = (uint32_t)c;
            c >>= 32;
        }
    }
    // s(e) from the grid-wide column totals (exact integers), canonical.  Under SKIP1 the value for e == 1 is
    // meaningless (never read).
    ZK_DEV static void finalize(Fe& out, int e, const unsigned long long* tot) {
        uint32_t limbs[W];
        carry_columns<W>(limbs, tot + e * W);
        if (D == 1) Fp<FID>::reduce9(out, limbs);
        else Fp<FID>::redc_wide(out, limbs);
        if (NLIN > 0) {   // linear part at X = e: S_lo + e (S_hi - S_lo)
            uint32_t l9[9];
            Fe slo, shi, diff, cur;
            carry_columns<9>(l9, tot + NE * W);
            Fp<FID>::reduce9(slo, l9);
            carry_columns<9>(l9, tot + NE * W + 9);
            Fp<FID>::reduce9(shi, l9);
            Fp<FID>::sub(diff, shi, slo);
            cur = slo;
            for (int x = 0; x < e; ++x) Fp<FID>::add(cur, cur, diff);
            Fp<FID>::add(out, out, cur);
        }
    }
};

}  // namespace zk
