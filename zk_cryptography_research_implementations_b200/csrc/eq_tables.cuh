// eq_tables.cuh -- tables of the multilinear equality polynomial eq(r, .) over the boolean hypercube, built on the GPU.
// Used by the sparse GKR prover / verifier (gkr_wide.cu: the bound `a` weights, eq(u, .), eq(v, .)) and by
// MultilinearPolynomial::evaluate in inner-product form (mle_eval.cu).
#pragma once
#include "kernels.cuh"

namespace zk {

// ---- eq tables.  w(a) = s1 eq(r1, a) [+ s2 eq(r2, a)] over k variables is built from half-width tables
// (eq over the leading kh and the trailing kl variables): one launch fills the two halves of a term straight from
// factors passed as kernel arguments (no staging copy, no stream synchronisation), one launch forms the outer
// products and adds the two terms.  A direct product would cost k multiplies per entry; this costs one per term.
constexpr int kEqHalfBits = 15;   // a half-table has at most 2^15 entries (layers are at most 2^30 wide)
struct EqHalfArgs {
    Fe* out[2];                   // hi table (scaled), lo table
    uint32_t bits[2];
    Fe factors[2][2 * kEqHalfBits];   // [half][2 v + bit]: 1 - r_v, r_v
    Fe scale;                     // multiplies the hi table
};
template <int FID> __global__ void __launch_bounds__(kThreads) eq_halves_kernel(const __grid_constant__ EqHalfArgs a) {
    const int h = blockIdx.y;
    const uint32_t nbits = a.bits[h];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, n = 1ull << nbits;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Fe acc;
        if (h == 0) acc = a.scale;
        else {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc.v[k] = FieldParams<FID>::r2(k);
            Fe one;
#pragma unroll
            for (int k = 0; k < 8; ++k) one.v[k] = (k == 0);
            Fp<FID>::mont_mul(acc, acc, one);   // R^2 / R = R: the Montgomery one
        }
        for (uint32_t v = 0; v < nbits; ++v) {
            const uint32_t bit = (uint32_t)(i >> (nbits - 1 - v)) & 1u;
            Fp<FID>::mont_mul(acc, acc, a.factors[h][2 * v + bit]);
        }
        st256(a.out[h] + i, acc);
    }
}
// out[a] = hi1[a >> lo_bits] lo1[a & mask] (+ hi2[..] lo2[..])
template <int FID, bool TWO>
__global__ void __launch_bounds__(kThreads) eq_outer2_kernel(Fe* out, const Fe* hi1, const Fe* lo1, const Fe* hi2, const Fe* lo2, uint32_t lo_bits, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, mask = (1ull << lo_bits) - 1;
    for (uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += stride) {
        Fe h = hi1[a >> lo_bits], l = lo1[a & mask], o;
        Fp<FID>::mont_mul(o, h, l);
        if (TWO) {
            Fe h2 = hi2[a >> lo_bits], l2 = lo2[a & mask], o2;
            Fp<FID>::mont_mul(o2, h2, l2);
            Fp<FID>::add(o, o, o2);
        }
        st256(out + a, o);
    }
}

}  // namespace zk
