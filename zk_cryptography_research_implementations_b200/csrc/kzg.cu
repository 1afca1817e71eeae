// kzg.cu -- multilinear KZG over BLS12-381 G1 on the GPU: the input commitment of succinct GKR
// (gkr/src/succinct_gkr_protocol.rs:35-169 -> multilinear_kzg/src/multilinear_kzg.rs:25-127, trusted_setup.rs:12-63).
//
// The reference commits with `sum_i power_i.mul_bigint(value_i)` -- 2^n independent double-and-add scalar
// multiplications -- and opens with n more such sums over the full setup, each against a quotient "blown up" to 2^n
// entries (multilinear_kzg.rs:93-108, :183-214).  Here:
//   * every sum is one bucket-method multi-scalar multiplication: signed c-bit digits of the canonical scalars, a
//     histogram / scan / scatter that groups the (window, digit) occurrences (no sort), bucket sums in three bounded
//     levels (one thread per segment of <= 128 entries adding affine points into an XYZZ accumulator, g1.cuh; segments run
//     longest first), running sums per chunk of a few buckets, the chunk results summed by
//     bit planes of the chunk index (one block per plane), and a host finish of ~450 group operations (one doubling per
//     scalar bit, one addition per plane, one inversion for the affine result);
//   * the blow-up is never materialised: the quotient of round k repeats with period 2^(n-k-1), so its sum against the
//     setup equals its sum against the setup FOLDED k+1 times (S_{k+1}[j] = S_k[j] + S_k[j + half]) -- the Lagrange basis
//     of the remaining variables -- which is built once per setup.  Openings cost 2^n point additions in total, not n 2^n;
//   * f - v is never formed: the quotient hi - lo does not see the constant, and v falls out of the last fold;
//   * the last 16 rounds of an opening (<= 2^15 points each) are summed in ONE grouped pass (MsmPlan::groups);
//   * several GPUs: every rank sums a contiguous share of the points, one 96-byte all-gather per call (zk_kzg_*_sharded).
// Group elements leave in affine form (canonical), so neither the coordinate system nor the order of the additions can show
// in a result: outputs are bit-identical to the reference's `P::G1` values converted with `into_affine()`.
#include <algorithm>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "internal.h"
#include "g1.cuh"
#include "host_curve.h"
#include "scan.cuh"
#include "msm_plan.cuh"
#include "msm_host.h"

using namespace zk;

#define ZK_CUDA(call)                                                        \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) {                                            \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__); \
            return ZK_ERR_CUDA;                                              \
        }                                                                    \
    } while (0)

namespace {
typedef Fp<BLS12_381_FR> Fr;
typedef unsigned long long u64;
#ifndef ZK_MSM_THREADS
#define ZK_MSM_THREADS 128
#endif
#ifndef ZK_MSM_MIN_BLOCKS
#define ZK_MSM_MIN_BLOCKS 3
#endif
#ifndef ZK_MSM_PREFETCH
#define ZK_MSM_PREFETCH 1
#endif
constexpr int kMsmThreads = ZK_MSM_THREADS;   // the group law needs ~150-250 registers: small blocks keep the SMs evenly filled
// open_and_prove sums the quotients of its last levels (2^kBatchLog points and fewer: each too small to fill the GPU) in ONE pass
constexpr uint32_t kBatchLog = 15;
constexpr int kFixedWindows = 32;    // fixed-base table of the generator: 32 windows of 8 bits

// ---------------------------------------------------------------- 16-byte vector moves of points
__device__ __forceinline__ G1Affine load_affine(const G1Affine* p) {
    G1Affine r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int i = 0; i < 6; ++i) d[i] = __ldg(s + i);
    return r;
}
__device__ __forceinline__ void store_affine(G1Affine* p, const G1Affine& v) {
    uint4* d = reinterpret_cast<uint4*>(p);
    const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
    for (int i = 0; i < 6; ++i) d[i] = s[i];
}
__device__ __forceinline__ G1Xyzz load_xyzz(const G1Xyzz* p) {
    G1Xyzz r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int i = 0; i < 12; ++i) d[i] = s[i];
    return r;
}
__device__ __forceinline__ void store_xyzz(G1Xyzz* p, const G1Xyzz& v) {
    uint4* d = reinterpret_cast<uint4*>(p);
    const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
    for (int i = 0; i < 12; ++i) d[i] = s[i];
}

// off[key + 1] += 1 for every non-zero digit (off zeroed before); key = window * B + bucket
__global__ void __launch_bounds__(kThreads) msm_count_kernel(const Fe* scalars, uint64_t n, MsmPlan pl, u64* off) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t k[8];
        canonical_scalar(k, scalars[i]);
        for_each_digit(k, pl, group_of(i, pl), [&](int w, uint32_t b, bool) { atomicAdd(off + (uint64_t)w * pl.B + b + 1, 1ull); });
    }
}
// entry = point index | sign << 31, to the next free place of its bucket (cursor starts as a copy of off[0..W*B))
__global__ void __launch_bounds__(kThreads) msm_scatter_kernel(const Fe* scalars, uint64_t n, MsmPlan pl, u64* cursor, uint32_t* sorted) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t k[8];
        canonical_scalar(k, scalars[i]);
        for_each_digit(k, pl, group_of(i, pl), [&](int w, uint32_t b, bool neg) {
            const u64 p = atomicAdd(cursor + (uint64_t)w * pl.B + b, 1ull);
            sorted[p] = (uint32_t)i | (neg ? 0x80000000u : 0u);
        });
    }
}

// ---------------------------------------------------------------- bucket sums
// A bucket can hold anything from nothing to every point (small or equal scalars put whole tables into a handful of buckets,
// and the top window of a width that does not divide 256 has only a few buckets in use), so the work is cut into pieces of
// bounded size in three levels, none of which needs the host:
//   level 0: every bucket is split into segments of <= cap0 entries; one thread per segment adds its affine points;
//   level 1: the segment sums of a bucket are grouped <= kCap1 at a time; one thread per group adds them;
//   level 2: one thread per bucket takes its single group sum -- or, where a bucket still has several, its warp adds them
//            lane-strided and folds the lanes with a butterfly.
// In the common case (every bucket within cap0) levels 1 and 2 are copies.
constexpr uint32_t kCap1 = 32;   // level 0's bound is MsmPlan::cap0 (8 .. 128)

// seg[b + 1] = ceil((off[b + 1] - off[b]) / cap), seg[0] = 0; an inclusive scan turns it into segment offsets
__global__ void __launch_bounds__(kThreads) msm_segments_kernel(const u64* off, uint64_t n_keys, uint32_t cap, u64* seg) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_keys; b += stride) {
        seg[b + 1] = (off[b + 1] - off[b] + cap - 1) / cap;
        if (b == 0) seg[0] = 0;
    }
}
// the bucket that owns segment s: the largest b with seg[b] <= s (s < seg[n_keys])
__device__ __forceinline__ uint64_t owner_of(const u64* seg, uint64_t n_keys, u64 s) {
    uint64_t lo = 0, hi = n_keys;   // invariant: seg[lo] <= s < seg[hi]
    while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (seg[mid] <= s) lo = mid;
        else hi = mid;
    }
    return lo;
}
// Level 0 runs its segments longest first (a counting sort on the segment length, <= 128): the 32 threads of a warp then
// add the same number of points, where bucket order would leave most of them waiting for the warp's fullest bucket.
constexpr int kLenBins = 129;
// where segment s starts in `sorted` and how long it is; histogram of the lengths (block-private, then global)
__global__ void __launch_bounds__(kThreads) msm_segdesc_kernel(const u64* off, const u64* seg0, uint64_t n_keys, uint64_t max_segments, uint32_t cap0,
                                                               u64* seg_lo, uint8_t* seg_len, unsigned* ghist) {
    __shared__ unsigned hist[kLenBins];
    for (int i = threadIdx.x; i < kLenBins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const uint64_t sidx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sidx < max_segments && sidx < seg0[n_keys]) {
        const uint64_t b = owner_of(seg0, n_keys, sidx);
        const u64 lo = off[b] + (sidx - seg0[b]) * cap0;
        const uint32_t len = (uint32_t)min((u64)cap0, off[b + 1] - lo);
        seg_lo[sidx] = lo;
        seg_len[sidx] = (uint8_t)len;
        atomicAdd(&hist[len], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kLenBins; i += blockDim.x)
        if (hist[i]) atomicAdd(&ghist[i], hist[i]);
}
// first rank of every length, longest first
__global__ void msm_lenscan_kernel(const unsigned* ghist, unsigned* gcursor) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned run = 0;
        for (int l = kLenBins - 1; l >= 0; --l) {
            gcursor[l] = run;
            run += ghist[l];
        }
    }
}
// order[rank] = segment: a block reserves a range per length, its threads take places inside
__global__ void __launch_bounds__(kThreads) msm_segplace_kernel(const u64* seg0, uint64_t n_keys, uint64_t max_segments, const uint8_t* seg_len,
                                                                unsigned* gcursor, uint32_t* order) {
    __shared__ unsigned hist[kLenBins], base[kLenBins];
    for (int i = threadIdx.x; i < kLenBins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const uint64_t sidx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = sidx < max_segments && sidx < seg0[n_keys];
    unsigned len = 0, rank = 0;
    if (valid) {
        len = seg_len[sidx];
        rank = atomicAdd(&hist[len], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kLenBins; i += blockDim.x)
        if (hist[i]) base[i] = atomicAdd(&gcursor[i], hist[i]);
    __syncthreads();
    if (valid) order[base[len] + rank] = (uint32_t)sidx;
}
__global__ void __launch_bounds__(kMsmThreads, ZK_MSM_MIN_BLOCKS) msm_bucket_kernel(const u64* seg0, const u64* seg_lo, const uint8_t* seg_len,
                                                                                      const uint32_t* order, const uint32_t* sorted,
                                                                                      const G1Affine* bases, uint64_t n_keys, uint64_t max_segments,
                                                                                      G1Xyzz* part0) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= max_segments || t >= seg0[n_keys]) return;
    const uint64_t sidx = order ? order[t] : t;
    const u64 lo = seg_lo[sidx], hi = lo + seg_len[sidx];
    G1Xyzz acc = G1::infinity();
#if ZK_MSM_PREFETCH
    // the next point travels while the current one is added: index load -> 96-byte gather is a dependent pair of misses
    uint32_t v = sorted[lo];
    G1Affine p = load_affine(bases + (v & 0x7fffffffu));
#pragma unroll 1
    for (u64 e = lo; e < hi; ++e) {
        uint32_t vn = 0;
        G1Affine pn = p;
        if (e + 1 < hi) {
            vn = sorted[e + 1];
            pn = load_affine(bases + (vn & 0x7fffffffu));
        }
        G1::add_affine(acc, p, (v >> 31) != 0);
        v = vn;
        p = pn;
    }
#else
#pragma unroll 1
    for (u64 e = lo; e < hi; ++e) {
        const uint32_t v = sorted[e];
        const G1Affine p = load_affine(bases + (v & 0x7fffffffu));
        G1::add_affine(acc, p, (v >> 31) != 0);
    }
#endif
    store_xyzz(part0 + sidx, acc);
}
__global__ void __launch_bounds__(kMsmThreads) msm_merge_kernel(const u64* seg0, const u64* seg1, const G1Xyzz* part0, uint64_t n_keys,
                                                                uint64_t max_segments, G1Xyzz* part1) {
    const uint64_t sidx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sidx >= max_segments || sidx >= seg1[n_keys]) return;
    const uint64_t b = owner_of(seg1, n_keys, sidx);
    const u64 lo = seg0[b] + (sidx - seg1[b]) * kCap1;
    const u64 hi = min(lo + (u64)kCap1, seg0[b + 1]);
    G1Xyzz acc = load_xyzz(part0 + lo);
#pragma unroll 1
    for (u64 e = lo + 1; e < hi; ++e) {
        const G1Xyzz v = load_xyzz(part0 + e);
        G1::add(acc, v);
    }
    store_xyzz(part1 + sidx, acc);
}
__device__ __forceinline__ G1Xyzz shfl_xyzz(const G1Xyzz& v, int src, bool down) {
    G1Xyzz r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 48; ++i) d[i] = down ? __shfl_down_sync(0xffffffffu, s[i], src) : __shfl_sync(0xffffffffu, s[i], src);
    return r;
}
__global__ void __launch_bounds__(kMsmThreads) msm_finish_kernel(const u64* seg1, const G1Xyzz* part1, uint64_t n_keys, G1Xyzz* buckets) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = b < n_keys;
    const u64 lo = valid ? seg1[b] : 0;
    const uint32_t cnt = valid ? (uint32_t)(seg1[b + 1] - lo) : 0;
    G1Xyzz acc = G1::infinity();
    if (cnt == 1) acc = load_xyzz(part1 + lo);
    unsigned big = __ballot_sync(0xffffffffu, cnt > 1);
    while (big) {   // the warp adds one oversized bucket at a time
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const u64 slo = __shfl_sync(0xffffffffu, lo, src);
        const uint32_t scnt = __shfl_sync(0xffffffffu, cnt, src);
        G1Xyzz part = G1::infinity();
#pragma unroll 1
        for (uint32_t t = lane; t < scnt; t += 32) {
            const G1Xyzz v = load_xyzz(part1 + slo + t);
            G1::add(part, v);
        }
#pragma unroll 1
        for (int delta = 16; delta >= 1; delta >>= 1) {
            const G1Xyzz o = shfl_xyzz(part, delta, true);
            if (lane < delta) G1::add(part, o);
        }
        const G1Xyzz total = shfl_xyzz(part, 0, false);
        if (lane == src) acc = total;
    }
    if (valid) store_xyzz(buckets + b, acc);
}

// ---------------------------------------------------------------- window sums: sum_b (b + 1) bucket[b]
// level 1: one thread per chunk of S buckets: run = sum of the chunk, acc = sum (b - lo + 1) bucket[b]
__global__ void __launch_bounds__(kMsmThreads) msm_chunk_kernel(const G1Xyzz* buckets, uint64_t n_chunks, uint32_t S, G1Xyzz* chunk_acc,
                                                                G1Xyzz* chunk_run) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_chunks) return;
    G1Xyzz run = G1::infinity(), acc = G1::infinity();
    const G1Xyzz* b = buckets + t * S;
#pragma unroll 1
    for (int i = (int)S - 1; i >= 0; --i) {
        const G1Xyzz v = load_xyzz(b + i);
        G1::add(run, v);
        G1::add(acc, run);
    }
    store_xyzz(chunk_acc + t, acc);
    store_xyzz(chunk_run + t, run);
}
// level 2: sum_t acc_t and sum_t t run_t over a window's nT chunks, the second as bit planes of t:
//   plane p < nb:  P_p = sum of run_t over the t whose bit p is set   (sum_t t run_t = sum_p 2^p P_p)
//   plane nb:      A   = sum of acc_t
// One block per (window, plane): every thread adds its share of the chunks, a five-step butterfly folds each warp, and the
// warp results meet in shared memory.
__global__ void __launch_bounds__(kMsmThreads) msm_plane_kernel(const G1Xyzz* chunk_acc, const G1Xyzz* chunk_run, uint32_t nT, uint32_t nb,
                                                                G1Xyzz* planes) {
    __shared__ G1Xyzz warp_sum[kMsmThreads / 32];
    const uint32_t plane = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t w = plane / (nb + 1), p = plane % (nb + 1);
    const G1Xyzz* src = (p == nb ? chunk_acc : chunk_run) + (uint64_t)w * nT;
    G1Xyzz acc = G1::infinity();
#pragma unroll 1
    for (uint32_t t = threadIdx.x; t < nT; t += blockDim.x) {
        if (p == nb || ((t >> p) & 1u)) {
            const G1Xyzz v = load_xyzz(src + t);
            G1::add(acc, v);
        }
    }
#pragma unroll 1
    for (int delta = 16; delta >= 1; delta >>= 1) {
        const G1Xyzz o = shfl_xyzz(acc, delta, true);
        if ((int)lane < delta) G1::add(acc, o);
    }
    if (lane == 0) warp_sum[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll 1
        for (int k = 1; k < kMsmThreads / 32; ++k) G1::add(acc, warp_sum[k]);
        store_xyzz(planes + plane, acc);
    }
}

// ---------------------------------------------------------------- trusted setup on the GPU
// compute_lagrange_basis (trusted_setup.rs:26-52): basis[index] = prod_i (bit_i(index) ? tau_i : 1 - tau_i), variable 0 = top bit
__global__ void __launch_bounds__(kThreads) lagrange_basis_kernel(const Fe* taus, uint32_t n, const __grid_constant__ Fe one, Fe* out) {
    const uint64_t len = 1ull << n, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t index = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; index < len; index += stride) {
        Fe e = one;
        for (uint32_t i = 0; i < n; ++i) {
            Fe f = taus[i];
            if (!((index >> (n - 1 - i)) & 1)) Fr::sub(f, one, f);
            Fr::mont_mul(e, e, f);
        }
        out[index] = e;
    }
}
// compute_g1_powers_of_tau (trusted_setup.rs:54-63): out[i] = scalar_i * G from the byte-window table of the generator
// (table[w][d] = d * 256^w * G, d = 0 unused), then one inversion per point for the affine form
__global__ void __launch_bounds__(kMsmThreads) fixed_base_kernel(const Fe* scalars, uint64_t n, const G1Affine* table, G1Affine* out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k[8];
    canonical_scalar(k, scalars[i]);
    G1Xyzz acc = G1::infinity();
#pragma unroll 1
    for (int w = 0; w < kFixedWindows; ++w) {
        const uint32_t d = (k[w >> 2] >> ((w & 3) * 8)) & 0xffu;
        if (d) {
            const G1Affine p = load_affine(table + w * 256 + d);
            G1::add_affine(acc, p);
        }
    }
    G1Affine r;
    if (G1::is_inf(acc)) {
        r.x = Fq381::zero();
        r.y = Fq381::zero();
    } else {
        Fq i3;
        G1::inv(i3, acc.zzz);
        G1::to_affine_with_inverse(r, acc, i3);
    }
    store_affine(out + i, r);
}
// the setup folded once more: out[j] = in[j] + in[j + half]
__global__ void __launch_bounds__(kMsmThreads) fold_points_kernel(const G1Affine* in, uint64_t half, G1Affine* out) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= half) return;
    G1Xyzz acc = G1::from_affine(load_affine(in + j));
    G1::add_affine(acc, load_affine(in + j + half));
    G1Affine r;
    if (G1::is_inf(acc)) {
        r.x = Fq381::zero();
        r.y = Fq381::zero();
    } else {
        Fq i3;
        G1::inv(i3, acc.zzz);
        G1::to_affine_with_inverse(r, acc, i3);
    }
    store_affine(out + j, r);
}
// flags |= 1 if a point is neither infinity nor on the curve (coordinates canonical and y^2 == x^3 + 4)
__global__ void __launch_bounds__(kThreads) check_points_kernel(const G1Affine* pts, uint64_t n, unsigned* flags) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const G1Affine p = load_affine(pts + i);
        if (G1::is_inf(p)) continue;
        Fq l, r, b, t;
        // canonical: adding zero through the reducing adder leaves a canonical value unchanged
        Fq z = Fq381::zero();
        Fq381::add(t, p.x, z);
        bool ok = Fq381::eq(t, p.x);
        Fq381::add(t, p.y, z);
        ok = ok && Fq381::eq(t, p.y);
        {
            constexpr uint32_t bm[12] = ZKC_B_32;   // 4 in Montgomery form
#pragma unroll
            for (int k = 0; k < 12; ++k) b.v[k] = bm[k];
        }
        Fq381::sqr(l, p.y);
        Fq381::sqr(r, p.x);
        Fq381::mul(r, r, p.x);
        Fq381::add(r, r, b);
        if (!ok || !Fq381::eq(l, r)) atomicOr(flags, 1u);
    }
}

// one round of open_and_prove (multilinear_kzg.rs:78-119): q[j] = hi - lo (the quotient, :166-181) and, in place,
// cur[j] = lo + r (hi - lo) (the remainder, partial_evaluate at variable 0)
__global__ void __launch_bounds__(kThreads) quotient_fold_kernel(Fe* cur, uint64_t half, const __grid_constant__ Fe r, Fe* q) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < half; j += stride) {
        const Fe lo = cur[j], hi = cur[j + half];
        Fe d, m;
        Fr::sub(d, hi, lo);
        Fr::mont_mul(m, d, r);
        Fr::add(m, m, lo);
        q[j] = d;
        cur[j] = m;
    }
}

// register-resident arithmetic, no memory traffic: the ceiling the bucket kernel is measured against.
// KIND 0: four independent chains of Fq products per thread; KIND 1: one chain of mixed additions (acc += P) per thread
template <int KIND> __global__ void __launch_bounds__(kMsmThreads, ZK_MSM_MIN_BLOCKS) g1_probe_kernel(G1Xyzz* out, uint32_t iters, const __grid_constant__ G1Affine g) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    G1Xyzz acc = G1::from_affine(g);
    acc.x.v[0] ^= (uint32_t)tid & 0xffu;     // per-thread data so nothing folds to a constant (not a curve point; timing only)
    if (KIND == 0) {
#pragma unroll 1
        for (uint32_t i = 0; i < iters; ++i) {
            Fq381::mul(acc.x, acc.x, g.x);
            Fq381::mul(acc.y, acc.y, g.y);
            Fq381::mul(acc.zz, acc.zz, g.x);
            Fq381::mul(acc.zzz, acc.zzz, g.y);
        }
    } else {
        G1Affine p = g;
#pragma unroll 1
        for (uint32_t i = 0; i < iters; ++i) {
            G1::add_affine(acc, p);
            p.x.v[0] ^= acc.x.v[0] & 1u;     // keeps the operand live without leaving the arithmetic
        }
    }
    store_xyzz(out + tid, acc);
}

int blocks_for(const zk_ctx* ctx, uint64_t work, int threads, int bps) {
    uint64_t blocks = (work + threads - 1) / threads, cap = (uint64_t)ctx->sm_count * bps;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}
}  // namespace

// ---------------------------------------------------------------- the setup object
struct zk_kzg_setup {
    uint32_t n = 0;                       // variables
    int device = 0;
    std::vector<G1Affine*> level;         // level[k]: the setup folded k times, 2^(n-k) points (level[0] = g1_powers_of_tau)
    G1Affine* storage = nullptr;          // all levels, 2^(n+1) points
    // workspace of the multi-scalar multiplication, sized for 2^n points
    u64 *off = nullptr, *cursor = nullptr, *scan_scratch = nullptr, *seg0 = nullptr, *seg1 = nullptr;
    uint32_t *sorted = nullptr, *order = nullptr;
    u64* seg_lo = nullptr;
    uint8_t* seg_len = nullptr;
    unsigned* len_hist = nullptr;       // [2][kLenBins]: histogram of the segment lengths, then the placement cursors
    G1Xyzz *buckets = nullptr, *part0 = nullptr, *part1 = nullptr, *chunk_acc = nullptr, *chunk_run = nullptr, *win = nullptr;
    HG1Xyzz* win_host = nullptr;          // pinned
    Fe *cur = nullptr, *quot = nullptr;   // open_and_prove: the remainder and quotient tables
    uint64_t cap_points = 0, cap_keys = 0, cap_chunks = 0;
};

namespace {
int msm_reserve(zk_ctx* ctx, zk_kzg_setup* s, uint64_t max_points) {
    uint64_t keys = 0, entries = 0, chunks = 0, segs0 = 0, planes = 0;
    auto account = [&](const MsmPlan& pl, uint64_t n) {
        const uint64_t k = (uint64_t)pl.groups * pl.W * pl.B, nT = pl.B / pl.S;
        uint32_t nb = 0;
        while ((1ull << nb) < nT) ++nb;
        keys = std::max<uint64_t>(keys, k);
        entries = std::max<uint64_t>(entries, n * pl.W);
        chunks = std::max<uint64_t>(chunks, (uint64_t)pl.groups * pl.W * nT);
        segs0 = std::max<uint64_t>(segs0, k + n * pl.W / pl.cap0);
        planes = std::max<uint64_t>(planes, (uint64_t)pl.groups * pl.W * (nb + 1));
    };
    for (uint64_t p2 = 1; p2 / 2 < max_points; p2 <<= 1) {   // every size open_and_prove will use, and max_points itself
        const uint64_t n = std::min(p2, max_points);
        account(plan_for(n), n);
    }
    for (uint32_t g = 1; g <= kBatchLog + 1; ++g) account(plan_for((2ull << (g - 1)) - 1, g, g - 1), (2ull << (g - 1)) - 1);   // batched small levels
    if (getenv("ZKB200_MSM_WINDOW")) {   // a forced window width: size for the widest plan
        keys = std::max<uint64_t>(keys, 128ull * 32768);
        entries = std::max<uint64_t>(entries, max_points * 128);
        chunks = std::max<uint64_t>(chunks, 128ull * 32768);
        segs0 = std::max<uint64_t>(segs0, keys + entries / 8);
        planes = std::max<uint64_t>(planes, (uint64_t)(kBatchLog + 1) * 128 * 17);
    }
    if (max_points <= s->cap_points && keys <= s->cap_keys && chunks <= s->cap_chunks) return ZK_OK;
    if (s->cap_points) return fail(ctx, ZK_ERR_ARG, "multi-scalar multiplication workspace is sized once");
    ZK_CUDA(cudaMalloc(&s->off, (keys + 1) * sizeof(u64)));
    ZK_CUDA(cudaMalloc(&s->cursor, keys * sizeof(u64)));
    ZK_CUDA(cudaMalloc(&s->scan_scratch, (scan::scan_chunks(keys + 1) + 1) * sizeof(u64)));
    ZK_CUDA(cudaMalloc(&s->sorted, entries * sizeof(uint32_t)));
    ZK_CUDA(cudaMalloc(&s->seg0, (keys + 1) * sizeof(u64)));
    ZK_CUDA(cudaMalloc(&s->seg1, (keys + 1) * sizeof(u64)));
    ZK_CUDA(cudaMalloc(&s->buckets, keys * sizeof(G1Xyzz)));
    ZK_CUDA(cudaMalloc(&s->part0, (segs0 + 1) * sizeof(G1Xyzz)));
    ZK_CUDA(cudaMalloc(&s->order, (segs0 + 1) * sizeof(uint32_t)));
    ZK_CUDA(cudaMalloc(&s->seg_lo, (segs0 + 1) * sizeof(u64)));
    ZK_CUDA(cudaMalloc(&s->seg_len, segs0 + 1));
    ZK_CUDA(cudaMalloc(&s->len_hist, 2 * 129 * sizeof(unsigned)));
    ZK_CUDA(cudaMalloc(&s->part1, (keys + segs0 / kCap1 + 1) * sizeof(G1Xyzz)));
    ZK_CUDA(cudaMalloc(&s->chunk_acc, chunks * sizeof(G1Xyzz)));
    ZK_CUDA(cudaMalloc(&s->chunk_run, chunks * sizeof(G1Xyzz)));
    ZK_CUDA(cudaMalloc(&s->win, planes * sizeof(G1Xyzz)));
    ZK_CUDA(cudaHostAlloc(&s->win_host, planes * sizeof(HG1Xyzz), cudaHostAllocDefault));
    s->cap_points = max_points;
    s->cap_keys = keys;
    s->cap_chunks = chunks;
    return ZK_OK;
}

// sum_i scalars[i] * bases[i] over n device-resident pairs -> affine result on the host.  groups > 1: `groups` sums in one
// pass over consecutive index ranges of halving size (n = 2^groups - 1, see MsmPlan); out[g] receives the sum of range g.
int g1_msm(zk_ctx* ctx, zk_kzg_setup* s, const Fe* scalars, const G1Affine* bases, uint64_t n, HG1Affine* out, uint32_t groups = 1) {
    static_assert(sizeof(HG1Xyzz) == sizeof(G1Xyzz) && sizeof(HG1Affine) == sizeof(G1Affine), "host and device point layouts");
    if (n >= (1ull << 31)) return fail(ctx, ZK_ERR_ARG, "multi-scalar multiplication over 2^31 or more points");
    if (groups > 1 && n != (1ull << groups) - 1) return fail(ctx, ZK_ERR_ARG, "grouped sum: n must be 2^groups - 1");
    const MsmPlan pl = plan_for(n, groups, groups - 1);
    const int WG = pl.W * (int)groups;                      // windows of all groups
    const uint64_t keys = (uint64_t)WG * pl.B;
    const uint32_t nT = pl.B / pl.S;
    cudaStream_t st = ctx->stream;
    ZK_CUDA(cudaMemsetAsync(s->off, 0, (keys + 1) * sizeof(u64), st));
    msm_count_kernel<<<blocks_for(ctx, n, kThreads, 8), kThreads, 0, st>>>(scalars, n, pl, s->off);
    scan::inclusive_scan(st, s->off, keys + 1, s->scan_scratch);
    ZK_CUDA(cudaMemcpyAsync(s->cursor, s->off, keys * sizeof(u64), cudaMemcpyDeviceToDevice, st));
    msm_scatter_kernel<<<blocks_for(ctx, n, kThreads, 8), kThreads, 0, st>>>(scalars, n, pl, s->cursor, s->sorted);
    // bucket sums in three bounded levels (see msm_bucket_kernel); the segment counts never come to the host, the grids
    // are sized by their upper bounds
    const uint64_t max0 = keys + n * (uint64_t)pl.W / pl.cap0, max1 = keys + max0 / kCap1;
    msm_segments_kernel<<<blocks_for(ctx, keys, kThreads, 8), kThreads, 0, st>>>(s->off, keys, pl.cap0, s->seg0);
    scan::inclusive_scan(st, s->seg0, keys + 1, s->scan_scratch);
    msm_segments_kernel<<<blocks_for(ctx, keys, kThreads, 8), kThreads, 0, st>>>(s->seg0, keys, kCap1, s->seg1);
    scan::inclusive_scan(st, s->seg1, keys + 1, s->scan_scratch);
    static const bool by_length = !(getenv("ZKB200_MSM_SORT") && atoi(getenv("ZKB200_MSM_SORT")) == 0);
    ZK_CUDA(cudaMemsetAsync(s->len_hist, 0, 2 * kLenBins * sizeof(unsigned), st));
    msm_segdesc_kernel<<<(unsigned)((max0 + kThreads - 1) / kThreads), kThreads, 0, st>>>(s->off, s->seg0, keys, max0, pl.cap0, s->seg_lo, s->seg_len, s->len_hist);
    if (by_length) {
        msm_lenscan_kernel<<<1, 32, 0, st>>>(s->len_hist, s->len_hist + kLenBins);
        msm_segplace_kernel<<<(unsigned)((max0 + kThreads - 1) / kThreads), kThreads, 0, st>>>(s->seg0, keys, max0, s->seg_len, s->len_hist + kLenBins, s->order);
    }
    msm_bucket_kernel<<<(unsigned)((max0 + kMsmThreads - 1) / kMsmThreads), kMsmThreads, 0, st>>>(s->seg0, s->seg_lo, s->seg_len, by_length ? s->order : nullptr,
                                                                                                   s->sorted, bases, keys, max0, s->part0);
    msm_merge_kernel<<<(unsigned)((max1 + kMsmThreads - 1) / kMsmThreads), kMsmThreads, 0, st>>>(s->seg0, s->seg1, s->part0, keys, max1, s->part1);
    msm_finish_kernel<<<(unsigned)((keys + kMsmThreads - 1) / kMsmThreads), kMsmThreads, 0, st>>>(s->seg1, s->part1, keys, s->buckets);
    const uint64_t n_chunks = (uint64_t)WG * nT;
    msm_chunk_kernel<<<(unsigned)((n_chunks + kMsmThreads - 1) / kMsmThreads), kMsmThreads, 0, st>>>(s->buckets, n_chunks, pl.S, s->chunk_acc, s->chunk_run);
    uint32_t nb = 0, log_s = 0;
    while ((1u << nb) < nT) ++nb;
    while ((1u << log_s) < pl.S) ++log_s;
    const uint32_t n_planes = (uint32_t)WG * (nb + 1);
    msm_plane_kernel<<<n_planes, kMsmThreads, 0, st>>>(s->chunk_acc, s->chunk_run, nT, nb, s->win);
    ctx->launches += 21;
    ZK_CUDA(cudaGetLastError());
    ZK_CUDA(cudaMemcpyAsync(s->win_host, s->win, (size_t)n_planes * sizeof(G1Xyzz), cudaMemcpyDeviceToHost, st));
    ZK_CUDA(cudaStreamSynchronize(st));
    // window_w = A_w + S sum_p 2^p P_{w,p};  result = sum_w 2^(c w) window_w = sum over bit positions: A_w sits at c w and
    // P_{w,p} at c w + log2(S) + p < c (w + 1).  One pass from the top bit down: a doubling per position, an addition per term.
    auto combine = [&](uint32_t g) { out[g] = msm_combine_planes(s->win_host + (size_t)g * pl.W * (nb + 1), pl, nb, log_s); };
    if (groups == 1) {
        combine(0);
    } else {   // the groups' finishes are independent: a few host threads
        const uint32_t nthreads = std::min<uint32_t>(groups, std::max(1u, std::min(8u, std::thread::hardware_concurrency())));
        std::vector<std::thread> pool;
        for (uint32_t t = 0; t < nthreads; ++t)
            pool.emplace_back([&, t]() {
                for (uint32_t g = t; g < groups; g += nthreads) combine(g);
            });
        for (std::thread& th : pool) th.join();
    }
    return ZK_OK;
}

int require_fr(zk_ctx* ctx) {
    if (ctx->fid != BLS12_381_FR) return fail(ctx, ZK_ERR_ARG, "multilinear KZG needs a BLS12-381 Fr context (the scalar field of the curve)");
    return ZK_OK;
}

// level[k + 1] from level[k], k = 0..n-1
int fold_levels(zk_ctx* ctx, zk_kzg_setup* s) {
    for (uint32_t k = 0; k < s->n; ++k) {
        const uint64_t half = 1ull << (s->n - k - 1);
        fold_points_kernel<<<(unsigned)((half + kMsmThreads - 1) / kMsmThreads), kMsmThreads, 0, ctx->stream>>>(s->level[k], half, s->level[k + 1]);
        ++ctx->launches;
    }
    ZK_CUDA(cudaGetLastError());
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

int setup_alloc(zk_ctx* ctx, uint32_t n, std::unique_ptr<zk_kzg_setup, void (*)(zk_kzg_setup*)>& s) {
    s->n = n;
    s->device = ctx->device;
    const uint64_t len = 1ull << n;
    ZK_CUDA(cudaSetDevice(ctx->device));
    ZK_CUDA(cudaMalloc(&s->storage, 2 * len * sizeof(G1Affine)));
    s->level.resize(n + 1);
    uint64_t o = 0;
    for (uint32_t k = 0; k <= n; ++k) {
        s->level[k] = s->storage + o;
        o += len >> k;
    }
    ZK_CUDA(cudaMalloc(&s->cur, len * sizeof(Fe)));
    ZK_CUDA(cudaMalloc(&s->quot, len * sizeof(Fe)));   // every round's quotient of one opening: 2^n - 1 elements
    return msm_reserve(ctx, s.get(), len);
}
void setup_delete(zk_kzg_setup* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    cudaFree(s->storage); cudaFree(s->off); cudaFree(s->cursor); cudaFree(s->scan_scratch); cudaFree(s->sorted);
    cudaFree(s->seg0); cudaFree(s->seg1); cudaFree(s->part0); cudaFree(s->part1);
    cudaFree(s->order); cudaFree(s->seg_lo); cudaFree(s->seg_len); cudaFree(s->len_hist);
    cudaFree(s->buckets); cudaFree(s->chunk_acc); cudaFree(s->chunk_run); cudaFree(s->win); cudaFree(s->cur); cudaFree(s->quot);
    if (s->win_host) cudaFreeHost(s->win_host);
    delete s;
}

// table[w][d] = d * 256^w * G in affine form (d = 0: infinity), built on the host with one shared inversion
void generator_table(std::vector<HG1Affine>& table) {
    typedef HostFq F;
    std::vector<HG1Xyzz> pts((size_t)kFixedWindows * 256, HostG1::infinity());
    HG1Xyzz base = HostG1::from_affine(HostG1::generator());
    for (int w = 0; w < kFixedWindows; ++w) {
        HG1Xyzz cur = base;
        for (int d = 1; d < 256; ++d) {
            pts[(size_t)w * 256 + d] = cur;
            cur = HostG1::add(cur, base);
        }
        base = cur;   // 256 * base
    }
    // Montgomery's trick over the ZZZ coordinates of the finite points
    std::vector<HFq> prefix(pts.size());
    HFq run = F::one();
    for (size_t i = 0; i < pts.size(); ++i) {
        prefix[i] = run;
        if (!pts[i].is_inf()) run = F::mul(run, pts[i].zzz);
    }
    HFq inv = F::inv(run);
    table.assign(pts.size(), HG1Affine{F::zero(), F::zero()});
    for (size_t i = pts.size(); i-- > 0;) {
        if (pts[i].is_inf()) continue;
        const HFq i3 = F::mul(inv, prefix[i]);
        inv = F::mul(inv, pts[i].zzz);
        const HFq izz = F::mul(F::sqr(i3), F::sqr(pts[i].zz));
        table[i] = HG1Affine{F::mul(pts[i].x, izz), F::mul(pts[i].y, i3)};
    }
}
}  // namespace

// ---------------------------------------------------------------- C-ABI
extern "C" int zk_kzg_setup_create(zk_ctx* ctx, const uint64_t* taus, uint32_t n, zk_kzg_setup** out) {
    int rc = require_fr(ctx);
    if (rc) return rc;
    if (n == 0) return fail(ctx, ZK_ERR_ASSERT, "requires at least one variable");   // trusted_setup.rs:28
    if (n > 28) return fail(ctx, ZK_ERR_ARG, "trusted setup over more than 28 variables");
    std::unique_ptr<zk_kzg_setup, void (*)(zk_kzg_setup*)> s(new zk_kzg_setup(), setup_delete);
    rc = setup_alloc(ctx, n, s);
    if (rc) return rc;
    const uint64_t len = 1ull << n;
    std::vector<HG1Affine> table;
    generator_table(table);
    G1Affine* d_table = nullptr;
    Fe* d_taus = nullptr;
    ZK_CUDA(cudaMalloc(&d_table, table.size() * sizeof(G1Affine)));
    std::unique_ptr<G1Affine, void (*)(G1Affine*)> g1(d_table, [](G1Affine* p) { cudaFree(p); });
    ZK_CUDA(cudaMalloc(&d_taus, (size_t)n * sizeof(Fe)));
    std::unique_ptr<Fe, void (*)(Fe*)> g2(d_taus, [](Fe* p) { cudaFree(p); });
    ZK_CUDA(cudaMemcpyAsync(d_table, table.data(), table.size() * sizeof(G1Affine), cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(cudaMemcpyAsync(d_taus, taus, (size_t)n * sizeof(Fe), cudaMemcpyHostToDevice, ctx->stream));
    Fe one;
    {
        const HFe h = ctx->field.one();
        memcpy(one.v, h.l, sizeof one);
    }
    lagrange_basis_kernel<<<blocks_for(ctx, len, kThreads, 8), kThreads, 0, ctx->stream>>>(d_taus, n, one, s->cur);
    fixed_base_kernel<<<(unsigned)((len + kMsmThreads - 1) / kMsmThreads), kMsmThreads, 0, ctx->stream>>>(s->cur, len, d_table, s->level[0]);
    ctx->launches += 2;
    ZK_CUDA(cudaGetLastError());
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    rc = fold_levels(ctx, s.get());
    if (rc) return rc;
    *out = s.release();
    return ZK_OK;
}

extern "C" int zk_kzg_setup_from_points(zk_ctx* ctx, const uint64_t* g1_points, uint32_t n, zk_kzg_setup** out) {
    int rc = require_fr(ctx);
    if (rc) return rc;
    if (n == 0) return fail(ctx, ZK_ERR_ASSERT, "requires at least one variable");
    if (n > 28) return fail(ctx, ZK_ERR_ARG, "trusted setup over more than 28 variables");
    std::unique_ptr<zk_kzg_setup, void (*)(zk_kzg_setup*)> s(new zk_kzg_setup(), setup_delete);
    rc = setup_alloc(ctx, n, s);
    if (rc) return rc;
    const uint64_t len = 1ull << n;
    ZK_CUDA(cudaMemcpyAsync(s->level[0], g1_points, len * sizeof(G1Affine), cudaMemcpyHostToDevice, ctx->stream));
    unsigned* flags = reinterpret_cast<unsigned*>(s->off);
    ZK_CUDA(cudaMemsetAsync(flags, 0, sizeof(unsigned), ctx->stream));
    check_points_kernel<<<blocks_for(ctx, len, kThreads, 8), kThreads, 0, ctx->stream>>>(s->level[0], len, flags);
    ++ctx->launches;
    unsigned h = 0;
    ZK_CUDA(cudaMemcpyAsync(&h, flags, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h) return fail(ctx, ZK_ERR_ARG, "a setup point is not on the curve");
    rc = fold_levels(ctx, s.get());
    if (rc) return rc;
    *out = s.release();
    return ZK_OK;
}

extern "C" void zk_kzg_setup_free(zk_ctx* ctx, zk_kzg_setup* s) {
    if (!s) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    setup_delete(s);
}
extern "C" uint32_t zk_kzg_setup_num_vars(const zk_kzg_setup* s) { return s->n; }

extern "C" int zk_kzg_setup_points(zk_ctx* ctx, const zk_kzg_setup* s, uint32_t level, uint64_t* out) {
    if (level > s->n) return fail(ctx, ZK_ERR_ARG, "setup level out of range");
    ZK_CUDA(cudaMemcpyAsync(out, s->level[level], ((size_t)1 << (s->n - level)) * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

extern "C" int zk_kzg_commit_device(zk_ctx* ctx, zk_kzg_setup* s, const zk_table* t, uint64_t out[12]) {
    int rc = require_fr(ctx);
    if (rc) return rc;
    if (t->len != (1ull << s->n)) return fail(ctx, ZK_ERR_ASSERT, "Polynomial evaluation must match g1 length");   // multilinear_kzg.rs:29-33
    HG1Affine r;
    rc = g1_msm(ctx, s, t->d, s->level[0], t->len, &r);
    if (rc) return rc;
    memcpy(out, &r, sizeof r);
    return ZK_OK;
}
extern "C" int zk_kzg_commit(zk_ctx* ctx, zk_kzg_setup* s, const uint64_t* vals, uint64_t len, uint64_t out[12]) {
    int rc = require_fr(ctx);
    if (rc) return rc;
    if (len != (1ull << s->n)) return fail(ctx, ZK_ERR_ASSERT, "Polynomial evaluation must match g1 length");
    ZK_CUDA(cudaMemcpyAsync(s->cur, vals, (size_t)len * sizeof(Fe), cudaMemcpyHostToDevice, ctx->stream));
    zk_table t;
    t.d = s->cur;
    t.len = t.cap = len;
    t.owned = false;
    return zk_kzg_commit_device(ctx, s, &t, out);
}

namespace {
// `cur` holds the polynomial (it is consumed).  All quotients are formed first -- round i's goes where level i + 1 of the
// setup sits relative to level 1, so the quotients of the last rounds and the setup levels they meet are two parallel runs
// of halving blocks -- then the large rounds are summed one by one and the small ones (<= 2^kBatchLog points) in one pass.
// sum over ranks, in rank order, of one partial affine point per rank and item: all[q * items + i]
void sum_partials(const std::vector<HG1Affine>& all, int world, size_t items, uint64_t* out) {
    for (size_t i = 0; i < items; ++i) {
        HG1Xyzz acc = HostG1::infinity();
        for (int q = 0; q < world; ++q) acc = HostG1::add(acc, HostG1::from_affine(all[(size_t)q * items + i]));
        const HG1Affine r = HostG1::to_affine(acc);
        memcpy(out + 12 * i, &r, sizeof r);
    }
}

int open_impl(zk_ctx* ctx, zk_kzg_setup* s, const uint64_t* opening, uint64_t eval[4], uint64_t* proofs, bool sharded = false) {
    const uint32_t n = s->n;
    const int G = sharded ? ctx->world : 1;
    static const bool batch_small = !(getenv("ZKB200_KZG_BATCH") && atoi(getenv("ZKB200_KZG_BATCH")) == 0);
    std::vector<uint64_t> qoff(n + 1, 0);
    for (uint32_t i = 0; i < n; ++i) qoff[i + 1] = qoff[i] + (1ull << (n - i - 1));
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t half = 1ull << (n - i - 1);
        Fe r;
        memcpy(r.v, opening + 4 * i, sizeof r);
        quotient_fold_kernel<<<blocks_for(ctx, half, kThreads, 8), kThreads, 0, ctx->stream>>>(s->cur, half, r, s->quot + qoff[i]);
        ++ctx->launches;
    }
    ZK_CUDA(cudaGetLastError());
    ZK_CUDA(cudaMemcpyAsync(eval, s->cur, sizeof(Fe), cudaMemcpyDeviceToHost, ctx->stream));
    uint32_t first_small = n;   // rounds first_small .. n-1 are summed together
    if (batch_small)
        for (uint32_t i = 0; i < n; ++i)
            if (n - i - 1 <= kBatchLog) { first_small = i; break; }
    if (n - first_small < 2) first_small = n;
    // the large rounds: over several ranks every rank sums its contiguous share of the points, the partial sums are
    // exchanged once at the end and added in rank order by every rank (group elements: the affine result is canonical)
    std::vector<HG1Affine> mine(first_small);
    for (uint32_t i = 0; i < first_small; ++i) {
        const uint64_t half = 1ull << (n - i - 1), cnt = half / G, lo = cnt * (G > 1 ? ctx->rank : 0);
        int rc = g1_msm(ctx, s, s->quot + qoff[i] + lo, s->level[i + 1] + lo, cnt, &mine[i]);
        if (rc) return rc;
    }
    if (first_small && G > 1) {
        std::vector<HG1Affine> all((size_t)G * first_small);   // G == ctx->world here: the gather fills world * bytes
        int rc = allgather_host_bytes(ctx, mine.data(), first_small * sizeof(HG1Affine), all.data());
        if (rc) return rc;
        sum_partials(all, G, first_small, proofs);
    } else if (first_small) {
        memcpy(proofs, mine.data(), first_small * sizeof(HG1Affine));
    }
    if (first_small < n) {
        const uint32_t groups = n - first_small;
        std::vector<HG1Affine> pr(groups);
        int rc = g1_msm(ctx, s, s->quot + qoff[first_small], s->level[first_small + 1], (1ull << groups) - 1, pr.data(), groups);
        if (rc) return rc;
        memcpy(proofs + 12 * first_small, pr.data(), groups * sizeof(HG1Affine));
    }
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}
int open_checks(zk_ctx* ctx, const zk_kzg_setup* s, uint64_t len, uint32_t n_opening) {
    int rc = require_fr(ctx);
    if (rc) return rc;
    if (len == 0 || (len & (len - 1))) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");
    uint32_t nv = 0;
    while ((1ull << nv) < len) ++nv;
    if (nv != n_opening) return fail(ctx, ZK_ERR_ASSERT, "number of polynomial variables must match length of opening values");   // :56-60
    if (n_opening != s->n) return fail(ctx, ZK_ERR_ASSERT, "Opening values must match number of variables from trusted setup"); // :61-65
    return ZK_OK;
}
}  // namespace

extern "C" int zk_kzg_open_device(zk_ctx* ctx, zk_kzg_setup* s, const zk_table* t, const uint64_t* opening, uint32_t n_opening,
                                  uint64_t eval[4], uint64_t* proofs) {
    int rc = open_checks(ctx, s, t->len, n_opening);
    if (rc) return rc;
    ZK_CUDA(cudaMemcpyAsync(s->cur, t->d, (size_t)t->len * sizeof(Fe), cudaMemcpyDeviceToDevice, ctx->stream));
    return open_impl(ctx, s, opening, eval, proofs);
}
extern "C" int zk_kzg_open(zk_ctx* ctx, zk_kzg_setup* s, const uint64_t* vals, uint64_t len, const uint64_t* opening, uint32_t n_opening,
                           uint64_t eval[4], uint64_t* proofs) {
    int rc = open_checks(ctx, s, len, n_opening);
    if (rc) return rc;
    ZK_CUDA(cudaMemcpyAsync(s->cur, vals, (size_t)len * sizeof(Fe), cudaMemcpyHostToDevice, ctx->stream));
    return open_impl(ctx, s, opening, eval, proofs);
}

// ---- several GPUs (one process per GPU, zk_comm_init done): setup and polynomial replicated on every rank (like the circuit and
// the input layer of zk_gkr_prove_wide_sharded); rank q sums the q-th contiguous share of the points of every large sum, the
// 96-byte partial results are all-gathered over NCCL and added in rank order, so every rank returns the same canonical points.
extern "C" int zk_kzg_commit_sharded(zk_ctx* ctx, zk_kzg_setup* s, const zk_table* t, uint64_t out[12]) {
    int rc = require_fr(ctx);
    if (rc) return rc;
    if (t->len != (1ull << s->n)) return fail(ctx, ZK_ERR_ASSERT, "Polynomial evaluation must match g1 length");
    const int G = ctx->world;
    if (G == 1 || t->len < (uint64_t)G * 2) return zk_kzg_commit_device(ctx, s, t, out);
    const uint64_t cnt = t->len / G, lo = cnt * ctx->rank;
    HG1Affine mine;
    rc = g1_msm(ctx, s, t->d + lo, s->level[0] + lo, cnt, &mine);
    if (rc) return rc;
    std::vector<HG1Affine> all(G);
    rc = allgather_host_bytes(ctx, &mine, sizeof mine, all.data());
    if (rc) return rc;
    sum_partials(all, G, 1, out);
    return ZK_OK;
}
extern "C" int zk_kzg_open_sharded(zk_ctx* ctx, zk_kzg_setup* s, const zk_table* t, const uint64_t* opening, uint32_t n_opening,
                                   uint64_t eval[4], uint64_t* proofs) {
    int rc = open_checks(ctx, s, t->len, n_opening);
    if (rc) return rc;
    ZK_CUDA(cudaMemcpyAsync(s->cur, t->d, (size_t)t->len * sizeof(Fe), cudaMemcpyDeviceToDevice, ctx->stream));
    return open_impl(ctx, s, opening, eval, proofs, ctx->world > 1);
}

// sum_i scalars[i] * points[i] for caller-supplied points (host arrays; any n >= 1)
extern "C" int zk_g1_msm(zk_ctx* ctx, const uint64_t* scalars, const uint64_t* points, uint64_t n, uint64_t out[12]) {
    int rc = require_fr(ctx);
    if (rc) return rc;
    if (n == 0) {
        memset(out, 0, 96);
        return ZK_OK;
    }
    std::unique_ptr<zk_kzg_setup, void (*)(zk_kzg_setup*)> s(new zk_kzg_setup(), setup_delete);
    s->device = ctx->device;
    ZK_CUDA(cudaSetDevice(ctx->device));
    ZK_CUDA(cudaMalloc(&s->storage, n * sizeof(G1Affine)));
    ZK_CUDA(cudaMalloc(&s->cur, n * sizeof(Fe)));
    rc = msm_reserve(ctx, s.get(), n);
    if (rc) return rc;
    ZK_CUDA(cudaMemcpyAsync(s->storage, points, n * sizeof(G1Affine), cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(cudaMemcpyAsync(s->cur, scalars, n * sizeof(Fe), cudaMemcpyHostToDevice, ctx->stream));
    unsigned* flags = reinterpret_cast<unsigned*>(s->off);
    ZK_CUDA(cudaMemsetAsync(flags, 0, sizeof(unsigned), ctx->stream));
    check_points_kernel<<<blocks_for(ctx, n, kThreads, 8), kThreads, 0, ctx->stream>>>(s->storage, n, flags);
    unsigned h = 0;
    ZK_CUDA(cudaMemcpyAsync(&h, flags, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h) return fail(ctx, ZK_ERR_ARG, "a point is not on the curve");
    HG1Affine r;
    rc = g1_msm(ctx, s.get(), s->cur, s->storage, n, &r);
    if (rc) return rc;
    memcpy(out, &r, sizeof r);
    return ZK_OK;
}

// Times register-resident curve arithmetic on a full grid: kind 0 = Fq (381-bit) Montgomery products (4 * iters per thread),
// kind 1 = mixed additions acc += P (iters per thread).  Returns operations per second: the IMAD-pipe ceiling of the MSM.
extern "C" int zk_g1_arith_probe(zk_ctx* ctx, int kind, uint32_t iters, int blocks_per_sm, double* ops_per_s, double* ms_out) {
    if (kind < 0 || kind > 1 || blocks_per_sm < 1 || blocks_per_sm > 16) return fail(ctx, ZK_ERR_ARG, "bad probe arguments");
    const int grid = ctx->sm_count * blocks_per_sm;
    int rc = ensure_scratch(ctx, (size_t)grid * kMsmThreads * sizeof(G1Xyzz));
    if (rc) return rc;
    G1Affine g;
    {
        const HG1Affine h = HostG1::generator();
        memcpy(&g, &h, sizeof g);
    }
    cudaEvent_t e0, e1;
    ZK_CUDA(cudaEventCreate(&e0));
    ZK_CUDA(cudaEventCreate(&e1));
    for (int pass = 0; pass < 2; ++pass) {   // first pass warms up
        ZK_CUDA(cudaEventRecord(e0, ctx->stream));
        if (kind == 0) g1_probe_kernel<0><<<grid, kMsmThreads, 0, ctx->stream>>>((G1Xyzz*)ctx->scratch, iters, g);
        else g1_probe_kernel<1><<<grid, kMsmThreads, 0, ctx->stream>>>((G1Xyzz*)ctx->scratch, iters, g);
        ++ctx->launches;
        ZK_CUDA(cudaGetLastError());
        ZK_CUDA(cudaEventRecord(e1, ctx->stream));
        ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    float ms = 0;
    ZK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_out) *ms_out = ms;
    if (ops_per_s) *ops_per_s = (double)grid * kMsmThreads * (kind == 0 ? 4.0 : 1.0) * iters / (ms * 1e-3);
    return ZK_OK;
}
