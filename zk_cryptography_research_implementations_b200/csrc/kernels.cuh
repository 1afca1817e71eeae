// kernels.cuh -- sm_100a kernels of the sumcheck / GKR prover hot path.
//
// Data layout in HBM: a table is N x 32 bytes, entry i = evaluation at the hypercube point whose
// MOST significant index bit is variable 0 (evaluation_form.rs:74-84), each entry 8 x u32 limbs of
// the Montgomery form.  Every prover fold binds variable 0, i.e. pairs (j, j + N/2): two perfectly
// coalesced streams.  Each thread moves whole 32-byte elements with one 256-bit LDG/STG.
//
// Per round ONE kernel runs (round k >= 1): it reads table_{k-1}, folds it by the challenge
// r_{k-1} in registers, writes table_k in place, and accumulates the round-k evaluations of the
// round polynomial from the freshly folded values; block partials go to a scratch array and the
// last block to finish (atomic ticket) reduces them and publishes the d+1 field elements.
#pragma once
#include <cuda_runtime.h>
#include "fp.cuh"

namespace zk {

#ifndef ZK_TWO_BLOCK_TABLES
#define ZK_TWO_BLOCK_TABLES 4   // up to this many tables the round kernels are compiled for two resident blocks per SM
#endif
constexpr int kMaxTables = 8;   // P * D
constexpr int kMaxEvals = 5;    // D + 1
#ifndef ZK_THREADS
#define ZK_THREADS 256
#endif
constexpr int kThreads = ZK_THREADS;   // threads per block of every kernel

struct TablePtrs {
    Fe* t[kMaxTables];
};

template <int FID> __device__ __forceinline__ void fold_by_scalar(Fe& out, const Fe& lo, const Fe& hi, const FoldTable& ft) {
    FoldScalar<FID>::fold(out, lo, hi, ft);
}

// ---------------------------------------------------------------- 256-bit global access
__device__ __forceinline__ Fe ld256(const Fe* p) {
    Fe r;
    asm volatile("ld.global.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                   "=r"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st256(Fe* p, const Fe& r) {
    asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.v[0]),
                 "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7])
                 : "memory");
}
// pull one 32-byte sector into L2 ahead of use (no register cost)
__device__ __forceinline__ void prefetch_l2(const Fe* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// coherent (L2) read of data another block published
__device__ __forceinline__ Fe ld256_cg(const Fe* p) {
    Fe r;
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                   "=r"(r.v[7])
                 : "l"(p));
    return r;
}

// ---------------------------------------------------------------- grid-wide reduction of NE field elements
// The published result of a round kernel: the d+1 field elements plus a sequence number written LAST
// (after a system-scope fence).  The mailbox lives in mapped pinned host memory -- for sharded runs in a
// segment shared by all rank processes -- so the host sees the round's result a PCIe write after the
// last block finishes, without a stream synchronisation or a D2H copy.
struct Mailbox {
    Fe vals[kMaxEvals];
    unsigned seq;
    unsigned pad[7];
};
struct ReduceScratch {
    Fe* partials;        // [gridDim.x][NE]
    unsigned* ticket;    // zero before the launch; reset by the last block
    Mailbox* out;        // mapped pinned host memory (device address)
    unsigned seq;        // value to publish in out->seq
};

template <int FID> __device__ __forceinline__ Fe warp_sum(Fe v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Fe o;
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = __shfl_down_sync(0xffffffffu, v.v[k], off);
        Fp<FID>::add(v, v, o);
    }
    return v;
}

// Sum `vals` over the block; result valid in thread 0.
template <int FID, int NE> __device__ __forceinline__ void block_sum(Fe (&vals)[NE]) {
    __shared__ Fe sm[NE][kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    __syncthreads();  // protect sm against a previous use
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        Fe w = warp_sum<FID>(vals[e]);
        if (lane == 0) sm[e][warp] = w;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            Fe w;
            if (lane < nwarps) w = sm[e][lane];
            else {
#pragma unroll
                for (int k = 0; k < 8; ++k) w.v[k] = 0;
            }
            vals[e] = warp_sum<FID>(w);
        }
    }
}

// Called by every thread of every block with its per-thread sums.
template <int FID, int NE> __device__ __forceinline__ void grid_sum_publish(Fe (&vals)[NE], const ReduceScratch& rs) {
    __shared__ bool is_last;
    block_sum<FID, NE>(vals);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int e = 0; e < NE; ++e) st256(&rs.partials[(size_t)blockIdx.x * NE + e], vals[e]);
        __threadfence();
        unsigned t = atomicAdd(rs.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        Fe acc;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc.v[k] = 0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
            Fe v = ld256_cg(&rs.partials[(size_t)b * NE + e]);
            Fp<FID>::add(acc, acc, v);
        }
        vals[e] = acc;
    }
    block_sum<FID, NE>(vals);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int e = 0; e < NE; ++e) rs.out->vals[e] = vals[e];
        *rs.ticket = 0;
        __threadfence_system();
        *reinterpret_cast<volatile unsigned*>(&rs.out->seq) = rs.seq;
    }
}

// ---------------------------------------------------------------- evaluation of the round polynomial
// Given the (lo, hi) values of all T = P*D tables at one pair index, accumulate the unreduced
// contributions to s(X), X = 0..D:  s(X) += sum_p prod_d (lo + X (hi - lo)).
// (sumcheck_gkr_protocol.rs:127-137 evaluates the same sums with D+1 full folds of every table.)
// D == 1 (plain sumcheck, prover.rs:74-89) accumulates 9-limb sums; D >= 2 accumulates 17-limb
// unreduced products.  SKIP1 leaves s(1) out (the host derives it from the running claim).
// NLIN extra tables enter the sum LINEARLY (a product with the all-ones table, which is then neither stored, folded
// nor multiplied): s(X) += sum_l (lo_l + X (hi_l - lo_l)).  The sparse GKR layer prover's phases have exactly that
// shape, h1*W + h2*1 (gkr_wide.cu).  Only the two half sums are accumulated; finish() spreads them over the s(X).
template <int FID, int P, int D, bool SKIP1, int NLIN = 0> struct RoundAcc {
    static constexpr int NE = D + 1;
    static constexpr int W = (D == 1) ? 9 : 17;
    static constexpr int T = P * D + NLIN;
    uint32_t acc[NE][W];
    uint32_t lin[NLIN > 0 ? 2 : 1][9];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int e = 0; e < NE; ++e)
#pragma unroll
            for (int k = 0; k < W; ++k) acc[e][k] = 0;
#pragma unroll
        for (int e = 0; e < (NLIN > 0 ? 2 : 1); ++e)
#pragma unroll
            for (int k = 0; k < 9; ++k) lin[e][k] = 0;
    }
    __device__ __forceinline__ void add_point(int e, const Fe (&v)[T]) {
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
    __device__ __forceinline__ void add_pair(const Fe (&lo)[T], const Fe (&hi)[T]) {
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
    __device__ __forceinline__ void finish(Fe (&out)[NE]) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (D == 1) Fp<FID>::reduce9(out[e], acc[e]);
            else Fp<FID>::redc_wide(out[e], acc[e]);
        }
        if (NLIN > 0) {   // linear part at X = 0, 1, 2, ...: S_lo, S_hi, 2 S_hi - S_lo, ...
            Fe slo, shi, cur, diff;
            Fp<FID>::reduce9(slo, lin[0]);
            Fp<FID>::reduce9(shi, lin[1]);
            Fp<FID>::add(out[0], out[0], slo);
            if (!SKIP1) Fp<FID>::add(out[1], out[1], shi);
            Fp<FID>::sub(diff, shi, slo);
            cur = shi;
#pragma unroll
            for (int x = 2; x <= D; ++x) {
                Fp<FID>::add(cur, cur, diff);
                Fp<FID>::add(out[x], out[x], cur);
            }
        }
    }
};

// Round 0: evaluations only.  half = N/2 pairs (j, j + half).
template <int FID, int P, int D, int NLIN = 0>
__global__ void __launch_bounds__(kThreads, (P * D + NLIN <= ZK_TWO_BLOCK_TABLES && D <= 2) ? 2 : 1) round_evals_kernel(TablePtrs tp, uint64_t half, ReduceScratch rs) {
    constexpr int T = P * D + NLIN;
    RoundAcc<FID, P, D, false, NLIN> ra;
    ra.init();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < half; j += stride) {
        if (j + stride < half) {   // next iteration's sectors start their trip from HBM now
#pragma unroll
            for (int t = 0; t < T; ++t) {
                prefetch_l2(tp.t[t] + j + stride);
                prefetch_l2(tp.t[t] + j + stride + half);
            }
        }
        Fe lo[T], hi[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            lo[t] = ld256(tp.t[t] + j);
            hi[t] = ld256(tp.t[t] + j + half);
        }
        ra.add_pair(lo, hi);
    }
    Fe out[D + 1];
    ra.finish(out);
    grid_sum_publish<FID, D + 1>(out, rs);
}

// Rounds k >= 1: fold table_{k-1} (4q entries per table) by r in place into table_k (2q entries) and
// evaluate round k on the folded values.  Thread j owns old entries j, j+q, j+2q, j+3q and new
// entries j, j+q -- it never touches another thread's data, so the in-place update is race free.
#ifndef ZK_FOLD_MIN_BLOCKS
#define ZK_FOLD_MIN_BLOCKS 2   // register cap so that two 256-thread blocks stay resident per SM for up to 3 tables
#endif
#ifndef ZK_PREFETCH_DIST
#define ZK_PREFETCH_DIST 1
#endif
template <int FID, int P, int D, bool SKIP1, int NLIN = 0>
__global__ void __launch_bounds__(kThreads, (P * D + NLIN <= ZK_TWO_BLOCK_TABLES && D <= 2) ? ZK_FOLD_MIN_BLOCKS : 1)
    fold_evals_kernel(TablePtrs tp, uint64_t q, const __grid_constant__ FoldTable ft, ReduceScratch rs) {
    constexpr int T = P * D + NLIN;
    RoundAcc<FID, P, D, SKIP1, NLIN> ra;
    ra.init();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < q; j += stride) {
        if (j + ZK_PREFETCH_DIST * stride < q) {   // a later iteration's sectors start their trip from HBM now
#pragma unroll
            for (int t = 0; t < T; ++t) {
#pragma unroll
                for (int s = 0; s < 4; ++s) prefetch_l2(tp.t[t] + j + ZK_PREFETCH_DIST * stride + s * q);
            }
        }
        Fe lo[T], hi[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            Fe a0 = ld256(tp.t[t] + j), a1 = ld256(tp.t[t] + j + q);
            Fe a2 = ld256(tp.t[t] + j + 2 * q), a3 = ld256(tp.t[t] + j + 3 * q);
            fold_by_scalar<FID>(lo[t], a0, a2, ft);
            fold_by_scalar<FID>(hi[t], a1, a3, ft);
            st256(tp.t[t] + j, lo[t]);
            st256(tp.t[t] + j + q, hi[t]);
        }
        ra.add_pair(lo, hi);
    }
    Fe out[D + 1];
    ra.finish(out);
    grid_sum_publish<FID, D + 1>(out, rs);
}

// Plain fold of variable 0, in place: table[j] = table[j] + r (table[j+half] - table[j]).
// (evaluation_form.rs:61-106 with evaluating_variable == 0.)
template <int FID>
__global__ void __launch_bounds__(kThreads)
    fold0_kernel(TablePtrs tp, int ntables, uint64_t half, const __grid_constant__ FoldTable ft) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (int t = 0; t < ntables; ++t) {
        Fe* tab = tp.t[t];
        for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < half; j += stride) {
            Fe a0 = ld256(tab + j), a1 = ld256(tab + j + half), o;
            fold_by_scalar<FID>(o, a0, a1, ft);
            st256(tab + j, o);
        }
    }
}

// Fold of an arbitrary variable, out of place (evaluation_form.rs:61-106, any evaluating_variable).
// out[i] = in[j] + r (in[j | 1<<power] - in[j]),  j = i with a zero inserted at bit `power`.
template <int FID>
__global__ void __launch_bounds__(kThreads)
    fold_var_kernel(const Fe* in, Fe* out, uint64_t half, uint32_t power, const __grid_constant__ FoldTable ft) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t low_mask = (1ull << power) - 1;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride) {
        uint64_t j = ((i & ~low_mask) << 1) | (i & low_mask);
        Fe a0 = ld256(in + j), a1 = ld256(in + (j | (1ull << power))), o;
        fold_by_scalar<FID>(o, a0, a1, ft);
        st256(out + i, o);
    }
}

// MLE evaluate: bind K leading variables in one pass (one read of the table, 2^-K of it written).
// in has m * 2^K entries, out has m.  Safe for out == in (thread j reads in[j + c m], writes out[j]).
struct FoldTables3 {
    FoldTable t[3];
};
template <int FID, int K>
__global__ void __launch_bounds__(kThreads)
    fold_multi_kernel(const Fe* in, Fe* out, uint64_t m, const __grid_constant__ FoldTables3 fts) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride) {
        Fe v[1 << K];
#pragma unroll
        for (int c = 0; c < (1 << K); ++c) v[c] = ld256(in + j + (uint64_t)c * m);
#pragma unroll
        for (int lvl = 0; lvl < K; ++lvl) {
            const int h = 1 << (K - 1 - lvl);
#pragma unroll
            for (int c = 0; c < h; ++c) fold_by_scalar<FID>(v[c], v[c], v[c + h], fts.t[lvl]);
        }
        st256(out + j, v[0]);
    }
}

// Sum of the two halves of one table (prover.rs:74-89) without folding == round_evals_kernel<FID,1,1>.

// ---------------------------------------------------------------- element-wise helpers (GKR table builders)
enum EwOp { EW_ADD = 0, EW_MUL = 1 };
// out[i] = a[i] (op) b[i]                      (evaluation_form.rs:145-163, product_polynomial.rs:66-70)
template <int FID, int OP> __global__ void __launch_bounds__(kThreads) ew_kernel(const Fe* a, const Fe* b, Fe* out, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Fe x = ld256(a + i), y = ld256(b + i), o;
        if (OP == EW_ADD) Fp<FID>::add(o, x, y);
        else Fp<FID>::mont_mul(o, x, y);
        st256(out + i, o);
    }
}
// out[b * n + c] = wb[b] (op) wc[c]             (evaluation_form.rs:108-143)
template <int FID, int OP>
__global__ void __launch_bounds__(kThreads) tensor_kernel(const Fe* wb, const Fe* wc, Fe* out, uint64_t n, uint32_t log_n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, total = n * n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        Fe x = ld256(wb + (i >> log_n)), y = ld256(wc + (i & (n - 1))), o;
        if (OP == EW_ADD) Fp<FID>::add(o, x, y);
        else Fp<FID>::mont_mul(o, x, y);
        st256(out + i, o);
    }
}
// out[i] = s * a[i]                             (evaluation_form.rs:49-57)
template <int FID> __global__ void __launch_bounds__(kThreads) scale_kernel(const Fe* a, Fe* out, uint64_t n, Fe s) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Fe x = ld256(a + i), o;
        Fp<FID>::mont_mul(o, x, s);
        st256(out + i, o);
    }
}
// SumPolynomial::add_polynomials_element_wise (sum_polynomial.rs:57-76): out[i] = sum_p prod_d t[p][d][i]
template <int FID> __global__ void __launch_bounds__(kThreads) sumpoly_reduce_kernel(TablePtrs tp, int P, int D, Fe* out, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Fe acc;
        for (int p = 0; p < P; ++p) {
            Fe prod = ld256(tp.t[p * D] + i);
            for (int d = 1; d < D; ++d) {
                Fe y = ld256(tp.t[p * D + d] + i);
                Fp<FID>::mont_mul(prod, prod, y);
            }
            if (p == 0) acc = prod;
            else Fp<FID>::add(acc, acc, prod);
        }
        st256(out + i, acc);
    }
}

// ---------------------------------------------------------------- convert_to_bytes (evaluation_form.rs:35-43)
// 32-byte big-endian canonical encoding of every entry: from-Montgomery, then byte reversal.
template <int FID> __global__ void __launch_bounds__(kThreads) to_bytes_be_kernel(const Fe* in, Fe* out, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    Fe one;
#pragma unroll
    for (int k = 0; k < 8; ++k) one.v[k] = (k == 0);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Fe x = ld256(in + i), c, o;
        Fp<FID>::mont_mul(c, x, one);
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = __byte_perm(c.v[7 - k], 0, 0x0123);
        st256(out + i, o);
    }
}

// ---------------------------------------------------------------- synthetic tables
// entry i of table `tid` under `seed` = from_le_bytes_mod_order(le64(w0)|le64(w1)|le64(w2)|le64(w3)),
// w_l = splitmix64(seed ^ tid * GOLDEN, counter = 4 i + l).  `first`/`step` select a shard: local
// entry j is global entry first + j * step.
__device__ __forceinline__ uint64_t splitmix64_at(uint64_t base, uint64_t ctr) {
    uint64_t z = base + (ctr + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
template <int FID>
__global__ void __launch_bounds__(kThreads)
    generate_kernel(Fe* out, uint64_t n, uint64_t seed, uint64_t table_id, uint64_t first, uint64_t step) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t base = seed ^ (table_id * 0x9E3779B97F4A7C15ull);
    Fe r2;
#pragma unroll
    for (int k = 0; k < 8; ++k) r2.v[k] = FieldParams<FID>::r2(k);
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const uint64_t g = first + j * step;
        uint32_t s[10];
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            uint64_t w = splitmix64_at(base, 4 * g + l);
            s[2 * l] = (uint32_t)w;
            s[2 * l + 1] = (uint32_t)(w >> 32);
        }
        s[8] = s[9] = 0;
        Fe plain, m;
        Fp<FID>::barrett(plain.v, s);
        Fp<FID>::mont_mul(m, plain, r2);
        st256(out + j, m);
    }
}

// ---------------------------------------------------------------- integer-multiply roofline probe
// Register-resident arithmetic, no memory traffic: every thread runs `iters` rounds of four
// independent chains.  KIND 0: mont_mul, 1: FoldScalar::fold, 2: mul_acc (unreduced product).
// The measured rate is the IMAD-pipe ceiling the round kernels are compared against.
template <int FID, int KIND>
__global__ void __launch_bounds__(kThreads) arith_probe_kernel(Fe* out, uint32_t iters, const __grid_constant__ FoldTable ft) {
    Fe x[4], y[4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x[c].v[k] = (threadIdx.x * 2654435761u + blockIdx.x + 977u * c + k) & 0x0fffffffu;
            y[c].v[k] = (threadIdx.x * 40503u + 31u * blockIdx.x + 13u * c + 7u * k) & 0x0fffffffu;
        }
    uint32_t acc[4][17];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int k = 0; k < 17; ++k) acc[c][k] = 0;
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (KIND == 0) Fp<FID>::mont_mul(x[c], x[c], y[c]);
            else if (KIND == 1) fold_by_scalar<FID>(x[c], x[c], y[c], ft);
            else { Fp<FID>::mul_acc(acc[c], x[c], y[c]); x[c].v[0] ^= acc[c][16]; }
        }
    }
    Fe r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) r.v[k] += x[c].v[k] * (2 * c + 1) + acc[c][k] + acc[c][k + 8];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// FP64 pipe probe: 8 independent DFMA chains per thread (kind 3 of zk_arith_probe).  Not used by any kernel;
// it answers whether a double-precision limb product (Emmart-style 52-bit limbs, 2 DFMA per product) could
// relieve the half-rate IMAD.WIDE pipe in a later round.
template <int UNUSED = 0> __global__ void __launch_bounds__(kThreads) dfma_probe_kernel(double* out, uint32_t iters) {
    double x[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = 1.0 + 1e-9 * (threadIdx.x + 13 * c + blockIdx.x);
    const double a = 1.0000001, b = 1e-12;
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = fma(x[c], a, b);
    }
    double r = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) r += x[c];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// integer-multiply pipe probes (kinds 4, 5, 6 of zk_arith_probe): 8 independent chains per thread of
//   4: mad.wide.u32 (IMAD.WIDE.U32, 64-bit accumulate, no carry flag)
//   5: mad.lo.u32   (IMAD, 32-bit)
//   6: mad.lo.cc / madc.hi.cc pairs (IMAD.WIDE.U32.X, carry chained) -- what fp.cuh emits
template <int KIND> __global__ void __launch_bounds__(kThreads) imad_probe_kernel(uint64_t* out, uint32_t iters) {
    uint64_t acc[8];
    uint32_t a[8], b = threadIdx.x * 2654435761u + 12345u;
#pragma unroll
    for (int c = 0; c < 8; ++c) { acc[c] = blockIdx.x + c; a[c] = threadIdx.x * 40503u + 977u * c + 1u; }
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (KIND == 4) {
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a[c]), "r"(b));
            } else if (KIND == 5) {
                uint32_t lo = (uint32_t)acc[c];
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(a[c]), "r"(b));
                acc[c] = lo;
            } else {
                uint32_t lo = (uint32_t)acc[c], hi = (uint32_t)(acc[c] >> 32);
                if (c == 0) ptx::mad_wide_cc(lo, hi, a[c], b);
                else ptx::madc_wide_cc(lo, hi, a[c], b);
                acc[c] = ((uint64_t)hi << 32) | lo;
            }
        }
    }
    uint64_t r = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) r += acc[c] * (2 * c + 1);
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

}  // namespace zk
