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
#include "round_acc.cuh"

namespace zk {

#ifndef ZK_TWO_BLOCK_TABLES
#define ZK_TWO_BLOCK_TABLES 4   // up to this many tables the round kernels are compiled for two resident blocks per SM
#endif
#ifndef ZK_THREADS
#define ZK_THREADS 256
#endif
constexpr int kThreads = ZK_THREADS;   // threads per block of every kernel

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
// cached (L1-allocating) 256-bit load, for small tables that are re-read many times (eq half tables)
__device__ __forceinline__ Fe ld256_ca(const Fe* p) {
    Fe r;
    asm volatile("ld.global.ca.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                   "=r"(r.v[7])
                 : "l"(p));
    return r;
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

// ---------------------------------------------------------------- grid-wide reduction of the round sums
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
    unsigned long long* gacc;   // [kMaxCols] grid-wide column totals; zero before the launch, re-zeroed by the last block
    unsigned* ticket;           // zero before the launch; reset by the last block
    Mailbox* out;               // mapped pinned host memory (device address)
    unsigned seq;               // value to publish in out->seq
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

// Sum `vals` over the block with modular additions; result valid in thread 0.  (General-purpose helper of the
// GKR table builders; the round kernels use the exact column sums below.)
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

// Exact integer sum of every thread's NC 32-bit columns over the block: tot[c] = sum over threads of col[c], valid
// for all threads after the call.  One REDUX pair per column and warp (16-bit halves, so nothing wraps), then NC
// threads add the per-warp sums.  Needs blockDim.x >= NC.
template <int NC> __device__ __forceinline__ void block_column_sums(const uint32_t (&col)[NC], unsigned long long* tot /* shared, [NC] */) {
    static_assert(NC <= kThreads, "one thread per column");
    __shared__ unsigned long long warp_cols[kThreads / 32][NC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        unsigned lo = __reduce_add_sync(0xffffffffu, col[c] & 0xffffu);
        unsigned hi = __reduce_add_sync(0xffffffffu, col[c] >> 16);
        if (lane == (c & 31)) warp_cols[warp][c] = (unsigned long long)lo + ((unsigned long long)hi << 16);
    }
    __syncthreads();
    if (threadIdx.x < NC) {
        unsigned long long s = 0;
        for (int w = 0; w < nwarps; ++w) s += warp_cols[w][threadIdx.x];
        tot[threadIdx.x] = s;
    }
    __syncthreads();
}

// Grid-wide version: every block adds its column sums into rs.gacc (one RED per column), takes a ticket, and the
// last block to arrive gets the grid totals in tot[] and returns true (all its threads); it also re-arms gacc and
// the ticket for the next launch.  A one-block grid skips the global stage altogether.
template <int NC> __device__ __forceinline__ bool grid_column_sums(const uint32_t (&col)[NC], const ReduceScratch& rs, unsigned long long* tot) {
    __shared__ bool is_last;
    block_column_sums<NC>(col, tot);
    if (gridDim.x == 1) return true;
    if (threadIdx.x < NC) {
        atomicAdd(rs.gacc + threadIdx.x, tot[threadIdx.x]);
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(rs.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    if (threadIdx.x < NC) {
        tot[threadIdx.x] = __ldcg(rs.gacc + threadIdx.x);
        rs.gacc[threadIdx.x] = 0;
    }
    if (threadIdx.x == 0) *rs.ticket = 0;
    __syncthreads();
    return true;
}

// Epilogue of a round kernel: called by every thread of every block with its accumulator.
template <class RA> __device__ __forceinline__ void publish_round(const RA& ra, const ReduceScratch& rs) {
    __shared__ unsigned long long tot[RA::NC];
    uint32_t col[RA::NC];
    ra.columns(col);
    if (!grid_column_sums<RA::NC>(col, rs, tot)) return;
    if (threadIdx.x < RA::NE) {
        Fe out;
        RA::finalize(out, threadIdx.x, tot);
        rs.out->vals[threadIdx.x] = out;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned*>(&rs.out->seq) = rs.seq;
    }
}

// Round 0: evaluations only.  half = N/2 pairs (j, j + half).
// SKIP1: s(1) is left out (the caller derives it from a claimed sum it knows to be true: the GKR layer prover).
template <int FID, int P, int D, int NLIN = 0, bool SKIP1 = false>
__global__ void __launch_bounds__(kThreads, (P * D + NLIN <= ZK_TWO_BLOCK_TABLES && D <= 2) ? 2 : 1) round_evals_kernel(TablePtrs tp, uint64_t half, ReduceScratch rs) {
    constexpr int T = P * D + NLIN;
    RoundAcc<FID, P, D, SKIP1, NLIN> ra;
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
    publish_round(ra, rs);
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
    publish_round(ra, rs);
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

}  // namespace zk
