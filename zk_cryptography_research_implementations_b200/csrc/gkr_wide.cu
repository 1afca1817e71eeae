// gkr_wide.cu -- GKR layer prover for WIDE layers: sparse two-phase form of the reference's layer sumcheck.
//
// The reference materialises add_i / mul_i as dense 2^(3i+2) tables and W(b)+W(c), W(b)W(c) as 4^(i+1)
// tensors (circuit/src/arithmetic_circuit.rs:126-163, gkr/src/utils.rs:8-68): quadratic in the layer width.
// With m = log2(width of the layer below), the same 2m round polynomials are obtained from tables of 2^m
// entries (checked numerically there and by tests/test_gpu_gkr.py against the dense
// oracle on reference-shaped circuits):
//
//   phase 1 (rounds over b):  sum_c f(b,c) = h1(b) W(b) + h2(b) * 1
//        h1(b) = sum_{gates g: left_g = b} w(out_g) (add_g ? 1 : W(right_g))
//        h2(b) = sum_{add gates g: left_g = b} w(out_g) W(right_g)
//   phase 2 (rounds over c, b bound to u):  f(u,c) = A(c) * 1 + B(c) W(c)
//        add_u(c) = sum_{add gates g: right_g = c} w(out_g) eq(u, left_g),  mul_u likewise
//        A(c) = W(u) add_u(c),   B(c) = add_u(c) + W(u) mul_u(c)
//
// where w(a) = eq(r_a, a) at the output layer and alpha eq(r_b, a) + beta eq(r_c, a) below it
// (gkr_protocol.rs:60-82).  Both phases are the fused <P=2,D=2> round kernel over 2^m entries; W(u) and W(v)
// fall out of the folded W tables, so the reference's evaluate_wb_wc (utils.rs:70-82) costs nothing extra.
//
// Layer shapes are explicit (`layer_bits`), so this also covers layered circuits the reference's rigid
// "layer i has i output bits" packing cannot express (BASELINE.json configs[3]: width 2^22 at every depth).
// For those the output claim generalises the reference's single challenge r_a to layer_bits[0] successive
// challenges; parity with the reference is defined (and tested) on reference-shaped circuits.
// Gate lists must be duplicate-free (the reference's dense indicator stores `= one`, its evaluation `+=`).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "internal.h"
#include "gkr_wide.h"
#include "eq_tables.cuh"

using namespace zk;

#define ZK_CUDA(call)                                                        \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) {                                            \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__); \
            return ZK_ERR_CUDA;                                              \
        }                                                                    \
    } while (0)

#define ZK_FID_SWITCH(ctx, EXPR)                              \
    switch ((ctx)->fid) {                                     \
        case 0: { constexpr int FID = 0; EXPR; } break;       \
        case 1: { constexpr int FID = 1; EXPR; } break;       \
        default: { constexpr int FID = 2; EXPR; } break;      \
    }

namespace {
inline int grid_of(const zk_ctx* ctx, uint64_t work, int bps) {
    uint64_t blocks = (work + kThreads - 1) / kThreads, cap = (uint64_t)ctx->sm_count * bps;
    if (blocks > cap) blocks = cap;
    if (ctx->grid_cap > 0 && blocks > (uint64_t)ctx->grid_cap) blocks = ctx->grid_cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

// Circuit::evaluate for one layer (arithmetic_circuit.rs:82-97): out[o] = sum over gates with output o of op(in[l], in[r])
template <int FID>
__global__ void __launch_bounds__(kThreads) eval_layer_kernel(GateCsr g, const Fe* in, Fe* out, uint64_t n_out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_out; o += stride) {
        Fe acc;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc.v[k] = 0;
        for (uint64_t i = g.off[o]; i < g.off[o + 1]; ++i) {
            Fe l = ld256(in + g.x[i]), r = ld256(in + g.y[i]), v;
            if (g.op[i] == 0) Fp<FID>::add(v, l, r);
            else Fp<FID>::mont_mul(v, l, r);
            Fp<FID>::add(acc, acc, v);
        }
        st256(out + o, acc);
    }
}

// same, one BLOCK per output: for reduction layers where a few outputs collect very many gates
template <int FID>
__global__ void __launch_bounds__(kThreads) eval_layer_block_kernel(GateCsr g, const Fe* in, Fe* out, uint64_t n_out) {
    for (uint64_t o = blockIdx.x; o < n_out; o += gridDim.x) {
        Fe acc[1];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[0].v[k] = 0;
        for (uint64_t i = g.off[o] + threadIdx.x; i < g.off[o + 1]; i += blockDim.x) {
            Fe l = ld256(in + g.x[i]), r = ld256(in + g.y[i]), v;
            if (g.op[i] == 0) Fp<FID>::add(v, l, r);
            else Fp<FID>::mont_mul(v, l, r);
            Fp<FID>::add(acc[0], acc[0], v);
        }
        block_sum<FID, 1>(acc);
        if (threadIdx.x == 0) st256(out + o, acc[0]);
    }
}

// heavy fan-in, few outputs (a reduction layer: 2 outputs collecting 2^21 gates each): every output is cut into S
// slices, one block per (output, slice) writes a partial sum, sum_slices_kernel adds the S partials.  A block per
// output alone would leave all but n_out SMs idle.
template <int FID>
__global__ void __launch_bounds__(kThreads) eval_layer_slices_kernel(GateCsr g, const Fe* in, Fe* partial, uint64_t n_out, uint32_t S) {
    for (uint64_t w = blockIdx.x; w < n_out * S; w += gridDim.x) {
        const uint64_t o = w / S, s = w % S;
        const uint64_t b = g.off[o], e = g.off[o + 1], chunk = (e - b + S - 1) / S;
        const uint64_t lo = b + s * chunk, hi = lo + chunk < e ? lo + chunk : e;
        Fe acc[1];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[0].v[k] = 0;
        for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            Fe l = ld256(in + g.x[i]), r = ld256(in + g.y[i]), v;
            if (g.op[i] == 0) Fp<FID>::add(v, l, r);
            else Fp<FID>::mont_mul(v, l, r);
            Fp<FID>::add(acc[0], acc[0], v);
        }
        block_sum<FID, 1>(acc);
        if (threadIdx.x == 0) st256(partial + w, acc[0]);
    }
}
// one block per output adds its S partial sums
template <int FID> __global__ void __launch_bounds__(kThreads) sum_slices_kernel(const Fe* partial, Fe* out, uint64_t n_out, uint32_t S) {
    for (uint64_t o = blockIdx.x; o < n_out; o += gridDim.x) {
        Fe acc[1];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[0].v[k] = 0;
        for (uint32_t s = threadIdx.x; s < S; s += blockDim.x) {
            Fe v = ld256(partial + o * S + s);
            Fp<FID>::add(acc[0], acc[0], v);
        }
        block_sum<FID, 1>(acc);
        if (threadIdx.x == 0) st256(out + o, acc[0]);
    }
}

// ---------------------------------------------------------------- warp-segmented bucket sums
// The table builders sum, per wire, a value over that wire's gates (its CSR bucket).  One thread per wire with a loop over
// its bucket wastes most of a warp: bucket lengths are Poisson-like (a third of the wires have no gate, a few have five), the
// warp runs max-over-lanes iterations with a dependent load chain each, and the random gathers -- what these kernels are
// made of -- are issued at a quarter of the possible rate (phase2_kernel: 410 us where a gate-parallel pass with the same
// gathers takes ~150 us, profiles/r02).  Here a WARP owns 32 consecutive wires, i.e. one contiguous range of gates, and
// walks that range gate-parallel (lane l takes gates S + l, S + l + 32, ...: coalesced index loads, every lane busy, all
// gathers in flight at once).  A gate's value goes to its wire's accumulator in shared memory as 32-bit limb columns
// (atomicAdd on the low word, the rare carry into a high word: exact integer sums, so the order is irrelevant and the
// result is the canonical field sum after ONE Barrett step per wire); the owning lane is found by a 5-step binary search
// over the 32 bucket starts held in the warp's registers.
constexpr int kSegWarps = kThreads / 32;
#ifndef ZK_SEG_MIN_BLOCKS
#define ZK_SEG_MIN_BLOCKS 4   // these kernels are chains of dependent loads per tile: resident warps, not registers, buy throughput
#endif
template <int FID, class Op>
__global__ void __launch_bounds__(kThreads, ZK_SEG_MIN_BLOCKS) seg_bucket_kernel(GateCsr g, uint64_t n_keys, const __grid_constant__ Op op) {
    constexpr int NV = Op::NV;
    __shared__ uint32_t acc_lo[kSegWarps][NV][8][32];
    __shared__ uint32_t acc_hi[kSegWarps][NV][8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t n_tiles = (n_keys + 31) / 32;
    for (uint64_t tile = (uint64_t)blockIdx.x * kSegWarps + warp; tile < n_tiles; tile += (uint64_t)gridDim.x * kSegWarps) {
        const uint64_t key = tile * 32 + lane;
        const uint64_t s = g.off[key < n_keys ? key : n_keys];                 // this lane's bucket start (sorted over the warp)
        const uint64_t e = g.off[key + 1 < n_keys ? key + 1 : n_keys];
        const uint64_t S = __shfl_sync(0xffffffffu, s, 0), E = __shfl_sync(0xffffffffu, e, 31);
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc_lo[warp][v][k][lane] = acc_hi[warp][v][k][lane] = 0;
        __syncwarp();
        for (uint64_t base = S; base < E; base += 32) {
            const uint64_t i = base + lane;
            const bool active = i < E;
            // owner = the last lane whose bucket starts at or before gate i
            int lo = 0, hi = 31;
#pragma unroll
            for (int step = 0; step < 5; ++step) {
                const int mid = (lo + hi + 1) >> 1;
                const uint64_t sm = __shfl_sync(0xffffffffu, s, mid);
                if (sm <= i) lo = mid;
                else hi = mid - 1;
            }
            if (active) {
                Fe val[NV];
                uint32_t present = 0;
                op.template gate<FID>(i, g.x[i], g.y[i], g.op[i], val, present);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    if (!((present >> v) & 1u)) continue;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t x = val[v].v[k];
                        const uint32_t old = atomicAdd(&acc_lo[warp][v][k][lo], x);
                        if (old + x < old) atomicAdd(&acc_hi[warp][v][k][lo], 1u);
                    }
                }
            }
        }
        __syncwarp();
        if (key < n_keys) {
            Fe sum[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                uint32_t limbs[9];
                unsigned long long c = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    c += (unsigned long long)acc_lo[warp][v][k][lane] + ((unsigned long long)acc_hi[warp][v][k][lane] << 32);
                    limbs[k] = (uint32_t)c;
                    c >>= 32;
                }
                limbs[8] = (uint32_t)c;
                Fp<FID>::reduce9(sum[v], limbs);
            }
            op.template finish<FID>(key, sum);
        }
        __syncwarp();
    }
}
// phase 1: h1(b) = sum_{left = b} w(out) (add ? 1 : W(right)),  h2(b) = sum_{add, left = b} w(out) W(right)     (x = out, y = right)
struct Phase1Op {
    static constexpr int NV = 2;
    const Fe *w, *W;
    Fe *h1, *h2;
    template <int FID> __device__ __forceinline__ void gate(uint64_t, uint32_t x, uint32_t y, uint8_t op, Fe (&val)[2], uint32_t& present) const {
        const Fe wv = ld256(w + x), wr = ld256(W + y);
        Fe t;
        Fp<FID>::mont_mul(t, wv, wr);
        if (op == 0) { val[0] = wv; val[1] = t; present = 3u; }
        else { val[0] = t; present = 1u; }
    }
    template <int FID> __device__ __forceinline__ void finish(uint64_t b, const Fe (&sum)[2]) const {
        st256(h1 + b, sum[0]);
        st256(h2 + b, sum[1]);
    }
};
// phase 2: add_u(c) = sum_{add, right = c} w(out) eq(u, left), mul_u likewise; A = W(u) add_u, B = add_u + W(u) mul_u   (x = out, y = left)
struct Phase2Op {
    static constexpr int NV = 2;
    const Fe *w, *equ;
    Fe *A, *B;
    FoldTable Wu;
    template <int FID> __device__ __forceinline__ void gate(uint64_t, uint32_t x, uint32_t y, uint8_t op, Fe (&val)[2], uint32_t& present) const {
        const Fe wv = ld256(w + x), e = ld256(equ + y);
        Fp<FID>::mont_mul(val[op ? 1 : 0], wv, e);
        present = op ? 2u : 1u;
    }
    template <int FID> __device__ __forceinline__ void finish(uint64_t c, const Fe (&sum)[2]) const {
        Fe zero, a, m, bsum;
#pragma unroll
        for (int k = 0; k < 8; ++k) zero.v[k] = 0;
        FoldScalar<FID>::fold(a, zero, sum[0], Wu);   // W(u) * add_u(c): a product by a per-launch constant
        FoldScalar<FID>::fold(m, zero, sum[1], Wu);
        Fp<FID>::add(bsum, sum[0], m);
        st256(A + c, a);
        st256(B + c, bsum);
    }
};
// phase 2 after the overlapped gate-wise half (phase2_pre_kernel): P_g EL[left & mask]
struct Phase2FinOp {
    static constexpr int NV = 2;
    const Fe *pg, *el;
    uint32_t mask;
    Fe *A, *B;
    FoldTable Wu;
    template <int FID> __device__ __forceinline__ void gate(uint64_t i, uint32_t, uint32_t y, uint8_t op, Fe (&val)[2], uint32_t& present) const {
        const Fe p = ld256(pg + i), e = ld256(el + (y & mask));
        Fp<FID>::mont_mul(val[op ? 1 : 0], p, e);
        present = op ? 2u : 1u;
    }
    template <int FID> __device__ __forceinline__ void finish(uint64_t c, const Fe (&sum)[2]) const {
        Fe zero, a, m, bsum;
#pragma unroll
        for (int k = 0; k < 8; ++k) zero.v[k] = 0;
        FoldScalar<FID>::fold(a, zero, sum[0], Wu);
        FoldScalar<FID>::fold(m, zero, sum[1], Wu);
        Fp<FID>::add(bsum, sum[0], m);
        st256(A + c, a);
        st256(B + c, bsum);
    }
};
inline int seg_grid(const zk_ctx* ctx, uint64_t n_keys) {
    uint64_t blocks = ((n_keys + 31) / 32 + kSegWarps - 1) / kSegWarps, cap = (uint64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (ctx->grid_cap > 0 && blocks > (uint64_t)ctx->grid_cap) blocks = ctx->grid_cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

// phase-1 tables: one thread per b.  (first, step): the b this launch covers are first + j * step, j < nb, written to
// position j -- a rank's shard of the phase tables (low index bits, comm.cu) or, with (0, 1), the whole table.
template <int FID>
__global__ void __launch_bounds__(kThreads) phase1_kernel(GateCsr g, const Fe* w, const Fe* W, Fe* h1, Fe* h2, uint64_t nb, uint64_t first, uint64_t step) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nb; j += stride) {
        const uint64_t b = first + j * step;
        Fe a1, a2;
#pragma unroll
        for (int k = 0; k < 8; ++k) a1.v[k] = a2.v[k] = 0;
        for (uint64_t i = g.off[b]; i < g.off[b + 1]; ++i) {
            Fe wv = ld256(w + g.x[i]), wc = ld256(W + g.y[i]), t;
            Fp<FID>::mont_mul(t, wv, wc);
            if (g.op[i] == 0) {
                Fp<FID>::add(a1, a1, wv);
                Fp<FID>::add(a2, a2, t);
            } else {
                Fp<FID>::add(a1, a1, t);
            }
        }
        st256(h1 + j, a1);
        st256(h2 + j, a2);
    }
}
// out[j] = in[first + j * step]: a rank's shard of a replicated table
__global__ void __launch_bounds__(kThreads) strided_copy_kernel(const Fe* in, Fe* out, uint64_t n, uint64_t first, uint64_t step) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) st256(out + j, ld256(in + first + j * step));
}

// phase-2 tables: one thread per c
template <int FID>
__global__ void __launch_bounds__(kThreads)
    phase2_kernel(GateCsr g, const Fe* w, const Fe* equ, const __grid_constant__ FoldTable Wu, Fe* A, Fe* B, uint64_t nc, uint64_t first, uint64_t step) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    Fe zero;
#pragma unroll
    for (int k = 0; k < 8; ++k) zero.v[k] = 0;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nc; j += stride) {
        const uint64_t c = first + j * step;
        Fe addu, mulu;
#pragma unroll
        for (int k = 0; k < 8; ++k) addu.v[k] = mulu.v[k] = 0;
        for (uint64_t i = g.off[c]; i < g.off[c + 1]; ++i) {
            Fe wv = ld256(w + g.x[i]), e = ld256(equ + g.y[i]), t;
            Fp<FID>::mont_mul(t, wv, e);
            if (g.op[i] == 0) Fp<FID>::add(addu, addu, t);
            else Fp<FID>::add(mulu, mulu, t);
        }
        Fe a, m, bsum;
        FoldScalar<FID>::fold(a, zero, addu, Wu);   // W(u) * add_u(c): a product by a per-launch constant (0 + W(u) (x - 0))
        FoldScalar<FID>::fold(m, zero, mulu, Wu);
        Fp<FID>::add(bsum, addu, m);
        st256(A + j, a);
        st256(B + j, bsum);
    }
}

// Phase 2 split at the challenges known when phase 1's rounds move into the persistent launch.  eq(u, left) factors as
// EH[left >> lo_bits] EL[left & mask] with EH = eq(u_0..u_{k-1}, .) over the first k challenges -- known by then -- and EL over the
// rest, known only when phase 1 ends.  While the (shrinking) round loop leaves most SMs idle, phase2_pre_kernel forms
// P_g = w(out_g) EH[left_g >> lo_bits] per gate, in by_right order: that is where the random HBM gather of w(.) happens, off the
// critical path.  Afterwards phase2_fin_kernel only streams P_g and looks EL up in a table of 2^(m-k) entries (L2-resident):
//   add_u(c) = sum_{add gates g: right_g = c} P_g EL[left_g & mask],  mul_u likewise;   A = W(u) add_u,  B = add_u + W(u) mul_u.
template <int FID>
__global__ void __launch_bounds__(kThreads) phase2_pre_kernel(GateCsr g, const Fe* w, const Fe* eh, uint32_t lo_bits, Fe* pg, uint64_t n_gates) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_gates; i += stride) {
        Fe wv = ld256(w + g.x[i]), e = ld256_ca(eh + (g.y[i] >> lo_bits)), t;
        Fp<FID>::mont_mul(t, wv, e);
        st256(pg + i, t);
    }
}
template <int FID>
__global__ void __launch_bounds__(kThreads)
    phase2_fin_kernel(GateCsr g, const Fe* pg, const Fe* el, uint32_t lo_bits, const __grid_constant__ FoldTable Wu, Fe* A, Fe* B, uint64_t nc, uint64_t first, uint64_t step) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t mask = (1u << lo_bits) - 1u;
    Fe zero;
#pragma unroll
    for (int k = 0; k < 8; ++k) zero.v[k] = 0;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nc; j += stride) {
        const uint64_t c = first + j * step;
        Fe addu, mulu;
#pragma unroll
        for (int k = 0; k < 8; ++k) addu.v[k] = mulu.v[k] = 0;
        for (uint64_t i = g.off[c]; i < g.off[c + 1]; ++i) {
            Fe p = ld256(pg + i), e = ld256(el + (g.y[i] & mask)), t;
            Fp<FID>::mont_mul(t, p, e);
            if (g.op[i] == 0) Fp<FID>::add(addu, addu, t);
            else Fp<FID>::add(mulu, mulu, t);
        }
        Fe a, m, bsum;
        FoldScalar<FID>::fold(a, zero, addu, Wu);
        FoldScalar<FID>::fold(m, zero, mulu, Wu);
        Fp<FID>::add(bsum, addu, m);
        st256(A + j, a);
        st256(B + j, bsum);
    }
}

// verifier: add_i / mul_i at the sumcheck point with `a` bound (gkr/src/utils.rs:84-135) from the gate list,
//   add_r = sum over add gates of w(out_g) eq(u, left_g) eq(v, right_g),   mul_r likewise,
// one thread per b = left index (by_left CSR: x = out, y = right); per-block partial sums, added on the host.
template <int FID>
__global__ void __launch_bounds__(kThreads) wiring_eval_kernel(GateCsr g, const Fe* w, const Fe* equ, const Fe* eqv, Fe* partial /* [gridDim.x][2] */, uint64_t nb) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    Fe acc[2];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[0].v[k] = acc[1].v[k] = 0;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += stride) {
        Fe sa, sm;
#pragma unroll
        for (int k = 0; k < 8; ++k) sa.v[k] = sm.v[k] = 0;
        for (uint64_t i = g.off[b]; i < g.off[b + 1]; ++i) {
            Fe wv = ld256(w + g.x[i]), ev = ld256(eqv + g.y[i]), t;
            Fp<FID>::mont_mul(t, wv, ev);
            if (g.op[i] == 0) Fp<FID>::add(sa, sa, t);
            else Fp<FID>::add(sm, sm, t);
        }
        Fe eu = ld256(equ + b), t;
        Fp<FID>::mont_mul(t, sa, eu);
        Fp<FID>::add(acc[0], acc[0], t);
        Fp<FID>::mont_mul(t, sm, eu);
        Fp<FID>::add(acc[1], acc[1], t);
    }
    block_sum<FID, 2>(acc);
    if (threadIdx.x == 0) {
        st256(partial + 2 * blockIdx.x, acc[0]);
        st256(partial + 2 * blockIdx.x + 1, acc[1]);
    }
}

// Row form of the outer product for wide tables: along a row the hi factor is a constant, so its eight multiples
// h 2^(32 i) mod p are formed once per row (8 threads, one Montgomery product each) and every entry costs a
// fold-by-scalar (84 multiplies) instead of a Montgomery product (137).  Rows have 2^lo_bits >= 256 entries.
struct Pow32Args {
    Fe pow32[8];   // Montgomery forms of 2^(32 i)
};
template <int FID, bool TWO>
__global__ void __launch_bounds__(kThreads) eq_outer_rows_kernel(Fe* out, const Fe* hi1, const Fe* lo1, const Fe* hi2, const Fe* lo2, uint32_t lo_bits,
                                                                 uint64_t n_rows, const __grid_constant__ Pow32Args pw) {
    __shared__ FoldTable tab[TWO ? 2 : 1];
    const uint64_t row_len = 1ull << lo_bits;
    Fe zero;
#pragma unroll
    for (int k = 0; k < 8; ++k) zero.v[k] = 0;
    for (uint64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
        __syncthreads();   // the previous row's readers are done with tab
        if (threadIdx.x < (TWO ? 16u : 8u)) {
            const int t = threadIdx.x >> 3, i = threadIdx.x & 7;
            Fe h = (t ? hi2 : hi1)[row], plain, r;
            Fp<FID>::redc256(plain.v, h.v);          // out of Montgomery form
            Fp<FID>::cond_sub_p(plain.v);
            Fp<FID>::mont_mul(r, plain, pw.pow32[i]);   // h 2^(32 i) mod p as a plain integer
#pragma unroll
            for (int k = 0; k < 8; ++k) tab[t].w[i][k] = r.v[k];
        }
        __syncthreads();
        for (uint64_t j = threadIdx.x; j < row_len; j += blockDim.x) {
            Fe x = lo1[j], o;
            FoldScalar<FID>::fold(o, zero, x, tab[0]);
            if (TWO) {
                Fe y = lo2[j], o2;
                FoldScalar<FID>::fold(o2, zero, y, tab[TWO ? 1 : 0]);
                Fp<FID>::add(o, o, o2);
            }
            st256(out + (row << lo_bits) + j, o);
        }
    }
}

int launch_eq_halves(zk_ctx* ctx, const std::vector<HFe>& r, const HFe& scale, Fe* hi, Fe* lo, uint32_t kh, uint32_t kl) {
    const HostField& f = ctx->field;
    EqHalfArgs a;
    a.out[0] = hi; a.out[1] = lo;
    a.bits[0] = kh; a.bits[1] = kl;
    for (uint32_t v = 0; v < kh + kl; ++v) {
        const int h = v < kh ? 0 : 1;
        const uint32_t j = v < kh ? v : v - kh;
        HFe omr = f.sub(f.one(), r[v]);
        memcpy(a.factors[h][2 * j].v, omr.l, 32);
        memcpy(a.factors[h][2 * j + 1].v, r[v].l, 32);
    }
    memcpy(a.scale.v, scale.l, 32);
    const uint64_t widest = 1ull << (kh > kl ? kh : kl);
    dim3 grid((unsigned)grid_of(ctx, widest, 4), 2);
    ZK_FID_SWITCH(ctx, (eq_halves_kernel<FID><<<grid, kThreads, 0, ctx->stream>>>(a)));
    ctx->launches++;
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}
// out[a] = s1 eq(r1, a) + s2 eq(r2, a) for all a in {0,1}^k (variable 0 = most significant bit); r2 == nullptr: one term
int build_eq2(zk_ctx* ctx, zk_wide_circuit* wc, const std::vector<HFe>& r1, const HFe& s1, const std::vector<HFe>* r2, const HFe& s2, Fe* out) {
    const uint32_t k = (uint32_t)r1.size(), kh = k / 2, kl = k - kh;
    if (kl > (uint32_t)kEqHalfBits) return fail(ctx, ZK_ERR_ARG, "layer wider than 2^30");
    int rc = launch_eq_halves(ctx, r1, s1, wc->half_hi.p, wc->half_lo.p, kh, kl);
    if (rc) return rc;
    // rows of at least one block's worth of entries; ZKB200_EQ_ROWS=0 forces the entry-wise kernel (test hook: the two
    // must give the same proof)
    const char* knob = getenv("ZKB200_EQ_ROWS");
    const bool rows = kl >= 8 && !(knob && knob[0] == '0');
    Pow32Args pw;
    memcpy(pw.pow32, ctx->pow32, sizeof pw.pow32);
    const uint64_t n_rows = 1ull << kh;
    const int row_grid = (int)std::min<uint64_t>(n_rows, (uint64_t)ctx->sm_count * 4);
    if (r2) {
        rc = launch_eq_halves(ctx, *r2, s2, wc->half_hi2.p, wc->half_lo2.p, kh, kl);
        if (rc) return rc;
        if (rows) {
            ZK_FID_SWITCH(ctx, (eq_outer_rows_kernel<FID, true><<<row_grid, kThreads, 0, ctx->stream>>>(
                                    out, wc->half_hi.p, wc->half_lo.p, wc->half_hi2.p, wc->half_lo2.p, kl, n_rows, pw)));
        } else {
            ZK_FID_SWITCH(ctx, (eq_outer2_kernel<FID, true><<<grid_of(ctx, 1ull << k, 4), kThreads, 0, ctx->stream>>>(
                                    out, wc->half_hi.p, wc->half_lo.p, wc->half_hi2.p, wc->half_lo2.p, kl, 1ull << k)));
        }
    } else if (rows) {
        ZK_FID_SWITCH(ctx, (eq_outer_rows_kernel<FID, false><<<row_grid, kThreads, 0, ctx->stream>>>(
                                out, wc->half_hi.p, wc->half_lo.p, nullptr, nullptr, kl, n_rows, pw)));
    } else {
        ZK_FID_SWITCH(ctx, (eq_outer2_kernel<FID, false><<<grid_of(ctx, 1ull << k, 4), kThreads, 0, ctx->stream>>>(
                                out, wc->half_hi.p, wc->half_lo.p, nullptr, nullptr, kl, 1ull << k)));
    }
    ctx->launches++;
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}
}  // namespace

// gkr_protocol::prove (gkr_protocol.rs:26-143), sparse two-phase layers.  Outputs as zk_gkr_prove; `output` may be
// NULL (wide output layers).  flags: ZK_FLAG_SKIP_ABSORB leaves the output layer out of the transcript (its absorb
// is a serial host Keccak over 32 * 2^bits[0] bytes).
static int gkr_prove_wide_impl(zk_ctx* ctx, const zk_wide_circuit* wc_, const uint64_t* inputs, const Fe* device_inputs, uint64_t n_inputs,
                              uint64_t* output, uint64_t* claimed_sum, uint64_t* layer_claims, uint64_t* coeffs_out,
                              uint64_t* challenges_out, uint64_t* wb_out, uint64_t* wc_out, uint32_t flags, bool sharded = false,
                              uint64_t collapse_len = 1 << 12);

extern "C" int zk_gkr_prove_wide(zk_ctx* ctx, const zk_wide_circuit* wc, const uint64_t* inputs, uint64_t n_inputs,
                                 uint64_t* output, uint64_t* claimed_sum, uint64_t* layer_claims, uint64_t* coeffs_out,
                                 uint64_t* challenges_out, uint64_t* wb_out, uint64_t* wc_out, uint32_t flags) {
    return gkr_prove_wide_impl(ctx, wc, inputs, nullptr, n_inputs, output, claimed_sum, layer_claims, coeffs_out, challenges_out,
                               wb_out, wc_out, flags);
}
// same, the input layer already resident in HBM (a zk_table of 2^layer_bits[L] entries; left untouched)
extern "C" int zk_gkr_prove_wide_device(zk_ctx* ctx, const zk_wide_circuit* wc, const zk_table* inputs,
                                        uint64_t* output, uint64_t* claimed_sum, uint64_t* layer_claims, uint64_t* coeffs_out,
                                        uint64_t* challenges_out, uint64_t* wb_out, uint64_t* wc_out, uint32_t flags) {
    return gkr_prove_wide_impl(ctx, wc, nullptr, inputs->d, inputs->len, output, claimed_sum, layer_claims, coeffs_out, challenges_out,
                               wb_out, wc_out, flags);
}

// gkr_protocol::prove with every layer sumcheck spread over the ranks of the communicator ("GKR: shard the
// 2^m phase tables the same way").  The circuit, its layer values and the transcript are REPLICATED -- every rank holds the
// circuit object, evaluates all layers and runs the same Fiat-Shamir steps -- while the expensive part is split: rank q
// builds the phase tables h1, h2 / A, B only for the wires b (resp. c) whose low index bits are q (1/G of the gate list's
// buckets, the random gathers go to the replicated w(.), eq(u, .) and W tables), takes its shard of W, and the phase's
// sumcheck runs as zk_prove_product_sharded (partial evaluations exchanged per round, collapse, redundant tail).
// Every rank returns the same proof.  `inputs`: the whole input layer on every rank.
extern "C" int zk_gkr_prove_wide_sharded(zk_ctx* ctx, const zk_wide_circuit* wc, const zk_table* inputs, uint64_t* output,
                                         uint64_t* claimed_sum, uint64_t* layer_claims, uint64_t* coeffs_out, uint64_t* challenges_out,
                                         uint64_t* wb_out, uint64_t* wc_out, uint32_t flags, uint64_t collapse_len) {
    return gkr_prove_wide_impl(ctx, wc, nullptr, inputs->d, inputs->len, output, claimed_sum, layer_claims, coeffs_out, challenges_out,
                               wb_out, wc_out, flags, zk_comm_world(ctx) > 1, collapse_len ? collapse_len : 1);
}

static int gkr_prove_wide_impl(zk_ctx* ctx, const zk_wide_circuit* wc_, const uint64_t* inputs, const Fe* device_inputs, uint64_t n_inputs,
                              uint64_t* output, uint64_t* claimed_sum, uint64_t* layer_claims, uint64_t* coeffs_out,
                              uint64_t* challenges_out, uint64_t* wb_out, uint64_t* wc_out, uint32_t flags, bool sharded, uint64_t collapse_len) {
    zk_wide_circuit* wc = const_cast<zk_wide_circuit*>(wc_);   // the workspace inside the circuit object is mutable
    // ZKB200_TRACE=1: coarse host-side timeline of one prove (stream synchronised at each mark)
    const bool trace = getenv("ZKB200_TRACE") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    double t_acc[6] = {0, 0, 0, 0, 0, 0};
    auto mark = [&](int slot) {
        if (!trace) return;
        cudaStreamSynchronize(ctx->stream);
        auto now = std::chrono::steady_clock::now();
        t_acc[slot] += std::chrono::duration<double, std::milli>(now - t_prev).count();
        t_prev = now;
    };
    const HostField& f = ctx->field;
    const uint32_t L = wc->L;
    if (n_inputs != (1ull << wc->bits[L])) return fail(ctx, ZK_ERR_ASSERT, "different number of variables");
    // ---- circuit.evaluate on the device: all layer values stay resident (gkr_protocol.rs:27)
    std::vector<DevBuf>& W = wc->W;
    if (device_inputs) ZK_CUDA(cudaMemcpyAsync(W[L].p, device_inputs, n_inputs * sizeof(Fe), cudaMemcpyDeviceToDevice, ctx->stream));
    else ZK_CUDA(cudaMemcpyAsync(W[L].p, inputs, n_inputs * sizeof(Fe), cudaMemcpyHostToDevice, ctx->stream));
    for (uint32_t li = L; li-- > 0;) {
        const uint64_t n_out = 1ull << wc->bits[li];
        const uint64_t full_grid = (uint64_t)ctx->sm_count * 4;
        if (wc->layers[li].n_gates / n_out >= 64 && n_out * 2 <= full_grid) {   // heavy fan-in, few outputs: slices
            const uint32_t S = (uint32_t)std::min<uint64_t>(full_grid / n_out, (wc->layers[li].n_gates / n_out + kThreads - 1) / kThreads);
            ZK_FID_SWITCH(ctx, (eval_layer_slices_kernel<FID><<<(int)(n_out * S), kThreads, 0, ctx->stream>>>(wc->layers[li].by_out, W[li + 1].p, wc->half_hi.p, n_out, S)));
            ZK_FID_SWITCH(ctx, (sum_slices_kernel<FID><<<(int)n_out, kThreads, 0, ctx->stream>>>(wc->half_hi.p, W[li].p, n_out, S)));
            ctx->launches++;
        } else if (wc->layers[li].n_gates / n_out >= 64) {   // heavy fan-in, many outputs: a block per output
            int blocks = (int)std::min<uint64_t>(n_out, full_grid);
            ZK_FID_SWITCH(ctx, (eval_layer_block_kernel<FID><<<blocks, kThreads, 0, ctx->stream>>>(wc->layers[li].by_out, W[li + 1].p, W[li].p, n_out)));
        } else {
            ZK_FID_SWITCH(ctx, (eval_layer_kernel<FID><<<grid_of(ctx, n_out, 4), kThreads, 0, ctx->stream>>>(wc->layers[li].by_out, W[li + 1].p, W[li].p, n_out)));
        }
        ctx->launches++;
    }
    ZK_CUDA(cudaGetLastError());
    mark(0);

    // ---- output layer: absorb W_0, bind its variables (gkr_protocol.rs:39-51; one challenge in the reference shape)
    HostTranscript tr;
    const uint64_t n0 = 1ull << wc->bits[0];
    std::vector<HFe> w0(n0);
    ZK_CUDA(cudaMemcpyAsync(w0.data(), W[0].p, n0 * sizeof(Fe), cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (output) memcpy(output, w0.data(), n0 * sizeof(Fe));
    if (!(flags & ZK_FLAG_SKIP_ABSORB))
        for (const HFe& x : w0) tr.append_be(f, x);
    std::vector<HFe> ra(wc->bits[0]);
    for (HFe& r : ra) r = tr.challenge(f);
    {
        std::vector<HFe> cur = w0;
        for (const HFe& r : ra) {
            size_t half = cur.size() / 2;
            for (size_t j = 0; j < half; ++j) cur[j] = f.add(cur[j], f.mul(r, f.sub(cur[j + half], cur[j])));
            cur.resize(half);
        }
        w0.assign(1, cur[0]);
    }
    HFe claim = w0[0];

    // how the layer sumchecks run their rounds / exchange their partial sums is the caller's choice; the proof does not depend on it
    // (the claims handed to them are sums this prover computed itself: round 0 may derive s(1) from them)
    uint32_t sc_flags = (flags & (ZK_FLAG_HOST_ROUNDS | ZK_FLAG_NCCL_EXCHANGE | ZK_FLAG_HOST_EXCHANGE | ZK_FLAG_DIRECT_S1)) | ZK_FLAG_TRUSTED_CLAIM;
    // sharded layers: the three-table phase rounds are latency-bound and the host's Keccak is faster than the device's, so
    // the per-round exchange goes through the host mailboxes by default (26.9 vs 29.3 ms at 8 GPUs, profiles/r02);
    // ZKB200_GKR_PEER_EXCHANGE=1 runs the sharded rounds in the persistent kernels with the in-kernel NVLink exchange
    if (sharded) {
        const char* knob = getenv("ZKB200_GKR_PEER_EXCHANGE");
        if (!(knob && knob[0] == '1')) sc_flags |= ZK_FLAG_HOST_EXCHANGE;
    }
    // overlap of the gate-wise phase-2 work with phase 1's latency rounds (ZKB200_GKR_OVERLAP=0 switches it off: A/B, tests);
    // a traced run synchronises at every stage mark, which would serialise it anyway
    const char* seg_knob = getenv("ZKB200_GKR_SEG");   // 0: the wire-per-thread builders (A/B, tests)
    const bool seg_buckets = !(seg_knob && seg_knob[0] == '0');
    const char* ov_knob = getenv("ZKB200_GKR_OVERLAP");
    const bool overlap = !(ov_knob && ov_knob[0] == '0') && !trace && wc->pre_pg.p != nullptr;
    if (overlap && !ctx->side_stream) {
        ZK_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
        ZK_CUDA(cudaEventCreateWithFlags(&ctx->side_event, cudaEventDisableTiming));
    }
    // ---- scratch tables
    DevBuf &wtab = wc->wtab, &eqa = wc->eqa, &h1 = wc->h1, &h2 = wc->h2, &Wc = wc->Wc;


    HFe alpha = f.zero(), beta = f.zero();
    std::vector<HFe> rb, rcv;
    uint64_t round_off = 0;
    for (uint32_t li = 0; li < L; ++li)
        if (wc->bits[li + 1] == 0) return fail(ctx, ZK_ERR_ARG, "every layer must read at least two wires");
    for (uint32_t li = 0; li < L; ++li) {
        const uint32_t m = wc->bits[li + 1];
        const uint64_t nm = 1ull << m;
        const WideLayer& wl = wc->layers[li];
        int rc;
        // ---- w(a): eq(r_a, .) at the output layer, alpha eq(r_b, .) + beta eq(r_c, .) below
        if (li == 0) {
            if ((rc = build_eq2(ctx, wc, ra, f.one(), nullptr, f.one(), wtab.p))) return rc;
        } else {
            if ((rc = build_eq2(ctx, wc, rb, alpha, &rcv, beta, wtab.p))) return rc;
        }
        mark(1);
        // a layer is spread over the ranks when every rank gets at least two wires; narrower layers run redundantly on all
        const uint64_t G = (sharded && nm >= 2 * (uint64_t)ctx->world) ? (uint64_t)ctx->world : 1, q = G > 1 ? (uint64_t)ctx->rank : 0;
        const uint64_t nl = nm / G;   // this rank's wires: b = q + j G
        // ---- phase 1 tables and sumcheck over b
        if (G == 1 && seg_buckets) {   // a warp per 32 wires, gate-parallel (a rank's strided shard keeps the wire-per-thread kernel)
            const Phase1Op op1{wtab.p, W[li + 1].p, h1.p, h2.p};
            ZK_FID_SWITCH(ctx, (seg_bucket_kernel<FID, Phase1Op><<<seg_grid(ctx, nm), kThreads, 0, ctx->stream>>>(wl.by_left, nm, op1)));
        } else {
            ZK_FID_SWITCH(ctx, (phase1_kernel<FID><<<grid_of(ctx, nl, 4), kThreads, 0, ctx->stream>>>(wl.by_left, wtab.p, W[li + 1].p, h1.p, h2.p, nl, q, G)));
        }
        ctx->launches++;
        if (G > 1) {
            strided_copy_kernel<<<grid_of(ctx, nl, 4), kThreads, 0, ctx->stream>>>(W[li + 1].p, Wc.p, nl, q, G);
            ctx->launches++;
        } else {
            ZK_CUDA(cudaMemcpyAsync(Wc.p, W[li + 1].p, nm * sizeof(Fe), cudaMemcpyDeviceToDevice, ctx->stream));
        }
        ZK_CUDA(cudaGetLastError());
        // h1*W + h2*1: one product plus one LINEAR table -- the all-ones factor is never materialised
        zk_table t_h1, t_W, t_h2;
        zk_table* tabs1[3] = {&t_h1, &t_W, &t_h2};
        Fe* ptr1[3] = {h1.p, Wc.p, h2.p};
        for (int i = 0; i < 3; ++i) { tabs1[i]->d = ptr1[i]; tabs1[i]->len = nl; tabs1[i]->cap = nm; tabs1[i]->owned = false; }
        mark(2);
        zk_sumpoly sp1;
        sp1.P = 1; sp1.D = 2; sp1.nlin = 1; sp1.len = nl;
        sp1.tabs.assign(tabs1, tabs1 + 3);
        memcpy(layer_claims + 4 * li, claim.l, 32);
        zk_transcript wrap;
        wrap.t = tr;
        uint64_t* chal = challenges_out + 4 * round_off;
        uint64_t* coef = coeffs_out + 12 * round_off;
        HFe fin1[4], fin2[4];
        // When phase 1's rounds move into the persistent launch, the challenges u_0..u_{k-1} are known and most SMs fall idle:
        // queue the gate-wise half of the phase-2 build (phase2_pre_kernel) on the side stream behind that launch.
        uint32_t pre_k = 0;
        int pre_rc = ZK_OK;
        if (overlap && wl.n_gates > 0) {
            ctx->dev_hook = [&](const uint64_t* first_chal) {
                if (pre_k || pre_rc) return;
                const uint32_t k = (uint32_t)((first_chal - chal) / 4);
                if (k < 2 || k + 2 > m || k > (uint32_t)kEqHalfBits) return;
                cudaStream_t main_stream = ctx->stream;
                ctx->stream = ctx->side_stream;   // the eq helpers launch on the context's stream
                std::vector<HFe> uk(reinterpret_cast<const HFe*>(chal), reinterpret_cast<const HFe*>(chal) + k);
                pre_rc = launch_eq_halves(ctx, uk, f.one(), wc->pre_eh.p, wc->half_lo2.p, k, 0);
                if (!pre_rc) {
                    ZK_FID_SWITCH(ctx, (phase2_pre_kernel<FID><<<grid_of(ctx, wl.n_gates, 4), kThreads, 0, ctx->stream>>>(wl.by_right, wtab.p, wc->pre_eh.p, m - k, wc->pre_pg.p, wl.n_gates)));
                    ctx->launches++;
                    if (cudaGetLastError() != cudaSuccess || cudaEventRecord(ctx->side_event, ctx->stream) != cudaSuccess) pre_rc = ZK_ERR_CUDA;
                }
                ctx->stream = main_stream;
                if (!pre_rc) pre_k = k;
            };
        }
        if (G > 1) rc = zk_prove_product_sharded(ctx, &sp1, claim.l, &wrap, coef, chal, fin1[0].l, sc_flags, collapse_len);
        else rc = zk_prove_product(ctx, &sp1, claim.l, &wrap, coef, chal, fin1[0].l, sc_flags);          // rounds 0..m-1
        ctx->dev_hook = nullptr;
        if (rc) return rc;
        if (pre_rc) return fail(ctx, ZK_ERR_CUDA, "GKR: the overlapped phase-2 precomputation failed to launch");
        mark(3);
        const HFe Wu = fin1[1];                                                                           // W(r_b)
        std::vector<HFe> u(reinterpret_cast<HFe*>(chal), reinterpret_cast<HFe*>(chal) + m);
        // ---- phase 2 tables and sumcheck over c
        const FoldTable Wu_ft = make_fold_table(f, Wu);
        if (pre_k) {   // the gate-wise products are (being) computed on the side stream: only the low part of eq(u, .) is missing
            std::vector<HFe> ulo(u.begin() + pre_k, u.end());
            if ((rc = build_eq2(ctx, wc, ulo, f.one(), nullptr, f.one(), eqa.p))) return rc;
            ZK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->side_event, 0));
            if (G == 1 && seg_buckets) {
                const Phase2FinOp opf{wc->pre_pg.p, eqa.p, (1u << (m - pre_k)) - 1u, h1.p, h2.p, Wu_ft};
                ZK_FID_SWITCH(ctx, (seg_bucket_kernel<FID, Phase2FinOp><<<seg_grid(ctx, nm), kThreads, 0, ctx->stream>>>(wl.by_right, nm, opf)));
            } else {
                ZK_FID_SWITCH(ctx, (phase2_fin_kernel<FID><<<grid_of(ctx, nl, 4), kThreads, 0, ctx->stream>>>(wl.by_right, wc->pre_pg.p, eqa.p, m - pre_k, Wu_ft, h1.p, h2.p, nl, q, G)));
            }
        } else {
            if ((rc = build_eq2(ctx, wc, u, f.one(), nullptr, f.one(), eqa.p))) return rc;
            if (G == 1 && seg_buckets) {
                const Phase2Op op2{wtab.p, eqa.p, h1.p, h2.p, Wu_ft};
                ZK_FID_SWITCH(ctx, (seg_bucket_kernel<FID, Phase2Op><<<seg_grid(ctx, nm), kThreads, 0, ctx->stream>>>(wl.by_right, nm, op2)));
            } else {
                ZK_FID_SWITCH(ctx, (phase2_kernel<FID><<<grid_of(ctx, nl, 4), kThreads, 0, ctx->stream>>>(wl.by_right, wtab.p, eqa.p, Wu_ft, h1.p, h2.p, nl, q, G)));
            }
        }
        ctx->launches += 1;
        ZK_CUDA(cudaGetLastError());
        mark(4);
        for (int i = 0; i < 3; ++i) { tabs1[i]->len = nl; }
        if (G > 1) {   // this rank's shard of the layer values, again (phase 1 folded the first copy away)
            strided_copy_kernel<<<grid_of(ctx, nl, 4), kThreads, 0, ctx->stream>>>(W[li + 1].p, Wc.p, nl, q, G);
            ctx->launches++;
            ZK_CUDA(cudaGetLastError());
        } else {
            // this sumcheck is the last reader of the layer's values (the next layer works on W[li + 2]): fold them in place
            // instead of copying 2^m elements into the scratch table first
            t_W.d = W[li + 1].p;
        }
        zk_table* tabs2[3] = {&t_h2, &t_W, &t_h1};                                                       // B*W + A*1 (A in h1, B in h2)
        zk_sumpoly sp2;
        sp2.P = 1; sp2.D = 2; sp2.nlin = 1; sp2.len = nl;
        sp2.tabs.assign(tabs2, tabs2 + 3);
        // the running claim entering round m is s_{m-1}(r_{m-1}); it is not absorbed again (one 2m-round sumcheck)
        HFe mid = f.horner(reinterpret_cast<HFe*>(coef + 12 * (m - 1)), 3, u[m - 1]);
        if (G > 1) rc = zk_prove_product_sharded(ctx, &sp2, mid.l, &wrap, coef + 12 * m, chal + 4 * m, fin2[0].l, sc_flags | ZK_FLAG_NO_CLAIM_ABSORB, collapse_len);
        else rc = zk_prove_product(ctx, &sp2, mid.l, &wrap, coef + 12 * m, chal + 4 * m, fin2[0].l, sc_flags | ZK_FLAG_NO_CLAIM_ABSORB);   // rounds m..2m-1
        if (rc) return rc;
        tr = wrap.t;
        mark(5);
        const HFe Wv = fin2[1];                                                                           // W(r_c)
        if (li + 1 < L) {                                                                                 // gkr_protocol.rs:109-132
            memcpy(wb_out + 4 * li, Wu.l, 32);
            memcpy(wc_out + 4 * li, Wv.l, 32);
            rb = u;
            rcv.assign(reinterpret_cast<HFe*>(chal) + m, reinterpret_cast<HFe*>(chal) + 2 * m);
            tr.append_be(f, Wu);
            alpha = tr.challenge(f);
            tr.append_be(f, Wv);
            beta = tr.challenge(f);
            claim = f.add(f.mul(alpha, Wu), f.mul(beta, Wv));
        }
        round_off += 2ull * m;
    }
    memcpy(claimed_sum, claim.l, 32);
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (trace)
        fprintf(stderr, "zk_gkr_prove_wide ms: upload+evaluate %.2f | w tables %.2f | phase-1 build %.2f | phase-1 sumcheck %.2f | eq(u)+phase-2 build %.2f | phase-2 sumcheck %.2f\n",
                t_acc[0], t_acc[1], t_acc[2], t_acc[3], t_acc[4], t_acc[5]);
    return ZK_OK;
}

// gkr_protocol::verify (gkr_protocol.rs:146-236) for circuits of explicit layer widths; the proof laid out as
// zk_gkr_prove_wide writes it.  The transcript replay and the per-round checks of the layer sumchecks are host work
// (sumcheck_gkr_protocol.rs:69-106); the claim helpers (utils.rs:84-135) evaluate add_i / mul_i at the sumcheck point from
// the gate list on the GPU (eq tables of u, v and the bound `a` weights, one pass over the layer's gates), and the input
// layer's W(u), W(v) (gkr_protocol.rs:188-194, utils.rs:70-82) are the evaluate kernels.  *ok = 1 iff the reference would
// return true.  `flags` must match the prover's (ZK_FLAG_SKIP_ABSORB).
static int gkr_verify_wide_impl(zk_ctx* ctx, const zk_wide_circuit* wc_, const uint64_t* output, const uint64_t* layer_claims,
                                const uint64_t* coeffs, const uint64_t* wb, const uint64_t* wcv_in, const uint64_t* inputs,
                                const zk_table* device_inputs, uint64_t n_inputs, uint32_t flags, int* ok, bool succinct = false,
                                const uint64_t* input_evals = nullptr, uint64_t* last_challenges = nullptr) {
    zk_wide_circuit* wc = const_cast<zk_wide_circuit*>(wc_);
    const HostField& f = ctx->field;
    const uint32_t L = wc->L;
    *ok = 0;
    for (uint32_t li = 0; li < L; ++li)
        if (wc->bits[li + 1] == 0) return fail(ctx, ZK_ERR_ARG, "every layer must read at least two wires");
    HostTranscript tr;
    const uint64_t n0 = 1ull << wc->bits[0];
    std::vector<HFe> w0(reinterpret_cast<const HFe*>(output), reinterpret_cast<const HFe*>(output) + n0);   // :153-159
    if (!(flags & ZK_FLAG_SKIP_ABSORB))
        for (const HFe& x : w0) tr.append_be(f, x);                                                          // :161
    std::vector<HFe> ra(wc->bits[0]);
    for (HFe& r : ra) r = tr.challenge(f);
    for (const HFe& r : ra) {                                                                                // :164
        size_t half = w0.size() / 2;
        for (size_t j = 0; j < half; ++j) w0[j] = f.add(w0[j], f.mul(r, f.sub(w0[j + half], w0[j])));
        w0.resize(half);
    }
    HFe claimed = w0[0], alpha = f.zero(), beta = f.zero();
    std::vector<HFe> prev_b, prev_c;
    uint64_t round_off = 0;
    const int grid_cap = ctx->sm_count * 4;
    int rc = ensure_scratch(ctx, (size_t)grid_cap * 2 * sizeof(Fe));
    if (rc) return rc;
    std::vector<HFe> partial((size_t)grid_cap * 2);
    zk_table* in_tab = nullptr;
    for (uint32_t li = 0; li < L; ++li) {
        const uint32_t m = wc->bits[li + 1], rounds = 2 * m;
        const uint64_t nm = 1ull << m;
        HFe layer_claim;
        memcpy(layer_claim.l, layer_claims + 4 * li, 32);
        if (claimed != layer_claim) return ZK_OK;                                                            // :167-169
        std::vector<HFe> chal(rounds);
        HFe last;
        int valid = 0;
        zk_transcript wrap;
        wrap.t = tr;
        rc = zk_verify_product(ctx->fid, layer_claim.l, coeffs + 12 * round_off, rounds, 2, &wrap, chal[0].l, last.l, &valid);
        tr = wrap.t;
        if (rc) return rc;
        if (!valid) return ZK_OK;                                                                            // :172-176
        std::vector<HFe> u(chal.begin(), chal.begin() + m), v(chal.begin() + m, chal.end());
        HFe wbv, wcv;
        if (succinct && li + 1 == L) {
            // succinct_gkr_protocol.rs:205-262: the input layer is not evaluated -- the caller checks the KZG openings of the
            // committed input polynomial at (rb, rc) = the two halves of this layer's challenges.  The reference stops here
            // without tying the opened values to this sumcheck's last claim; with `input_evals` (the two opened values) that
            // check is made as for every other layer.
            if (last_challenges) memcpy(last_challenges, chal[0].l, (size_t)rounds * 32);
            if (!input_evals) break;
            memcpy(wbv.l, input_evals, 32);
            memcpy(wcv.l, input_evals + 4, 32);
        } else if (li + 1 < L) {                                                                             // :183-187
            memcpy(wbv.l, wb + 4 * li, 32);
            memcpy(wcv.l, wcv_in + 4 * li, 32);
        } else {                                                                                             // :188-194
            if (n_inputs != nm) return ZK_OK;
            const zk_table* t = device_inputs;
            if (!t) {
                rc = zk_table_upload(ctx, inputs, n_inputs, &in_tab);
                if (rc) return rc;
                t = in_tab;
            }
            rc = zk_mle_evaluate(ctx, t, u[0].l, m, wbv.l);                                                  // utils.rs:70-82
            if (!rc) rc = zk_mle_evaluate(ctx, t, v[0].l, m, wcv.l);
            if (in_tab) { zk_table_free(ctx, in_tab); in_tab = nullptr; }
            if (rc) return rc;
        }
        // bound `a` weights, eq(u, .), eq(v, .) -- utils.rs:84-135 without the dense 2^(3i+2) tables
        if (li == 0) rc = build_eq2(ctx, wc, ra, f.one(), nullptr, f.one(), wc->wtab.p);
        else rc = build_eq2(ctx, wc, prev_b, alpha, &prev_c, beta, wc->wtab.p);
        if (!rc) rc = build_eq2(ctx, wc, u, f.one(), nullptr, f.one(), wc->eqa.p);
        if (!rc) rc = build_eq2(ctx, wc, v, f.one(), nullptr, f.one(), wc->h1.p);
        if (rc) return rc;
        const int grid = grid_of(ctx, nm, 4);
        ZK_FID_SWITCH(ctx, (wiring_eval_kernel<FID><<<grid, kThreads, 0, ctx->stream>>>(wc->layers[li].by_left, wc->wtab.p, wc->eqa.p, wc->h1.p, (Fe*)ctx->scratch, nm)));
        ctx->launches++;
        ZK_CUDA(cudaGetLastError());
        ZK_CUDA(cudaMemcpyAsync(partial.data(), ctx->scratch, (size_t)grid * 2 * sizeof(Fe), cudaMemcpyDeviceToHost, ctx->stream));
        ZK_CUDA(cudaStreamSynchronize(ctx->stream));
        HFe add_r = f.zero(), mul_r = f.zero();
        for (int b = 0; b < grid; ++b) {
            add_r = f.add(add_r, partial[2 * b]);
            mul_r = f.add(mul_r, partial[2 * b + 1]);
        }
        const HFe expected = f.add(f.mul(add_r, f.add(wbv, wcv)), f.mul(mul_r, f.mul(wbv, wcv)));           // utils.rs:110,134
        if (expected != last) return ZK_OK;                                                                  // :220-222
        prev_b = u;                                                                                          // :224
        prev_c = v;
        tr.append_be(f, wbv);
        alpha = tr.challenge(f);                                                                             // :226-227
        tr.append_be(f, wcv);
        beta = tr.challenge(f);                                                                              // :229-230
        claimed = f.add(f.mul(alpha, wbv), f.mul(beta, wcv));                                                // :232
        round_off += rounds;
    }
    *ok = 1;
    return ZK_OK;
}

extern "C" int zk_gkr_verify_wide(zk_ctx* ctx, const zk_wide_circuit* wc, const uint64_t* output, const uint64_t* layer_claims,
                                  const uint64_t* coeffs, const uint64_t* wb, const uint64_t* wcv, const uint64_t* inputs,
                                  uint64_t n_inputs, uint32_t flags, int* ok) {
    return gkr_verify_wide_impl(ctx, wc, output, layer_claims, coeffs, wb, wcv, inputs, nullptr, n_inputs, flags, ok);
}
// verify_succinct's sumcheck half (succinct_gkr_protocol.rs:172-262): everything of zk_gkr_verify_wide except the input layer
extern "C" int zk_gkr_verify_wide_succinct(zk_ctx* ctx, const zk_wide_circuit* wc, const uint64_t* output, const uint64_t* layer_claims,
                                           const uint64_t* coeffs, const uint64_t* wb, const uint64_t* wcv, const uint64_t* input_evals,
                                           uint32_t flags, uint64_t* last_challenges, int* ok) {
    return gkr_verify_wide_impl(ctx, wc, output, layer_claims, coeffs, wb, wcv, nullptr, nullptr, 0, flags, ok, true, input_evals, last_challenges);
}
extern "C" int zk_gkr_verify_wide_device(zk_ctx* ctx, const zk_wide_circuit* wc, const uint64_t* output, const uint64_t* layer_claims,
                                         const uint64_t* coeffs, const uint64_t* wb, const uint64_t* wcv, const zk_table* inputs,
                                         uint32_t flags, int* ok) {
    return gkr_verify_wide_impl(ctx, wc, output, layer_claims, coeffs, wb, wcv, nullptr, inputs, inputs->len, flags, ok);
}
