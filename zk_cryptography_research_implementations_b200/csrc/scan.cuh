// scan.cuh -- in-place inclusive prefix sum of a 64-bit array on the device in three launches (per-chunk totals, a
// one-block scan of the totals, the chunks again).  Used by the CSR builders of the wide GKR circuit and by the bucket
// sort of the multi-scalar multiplication.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.cuh"

namespace zk { namespace scan {
typedef unsigned long long u64;
constexpr int kScanItems = 8;                          // items per thread of the scan kernels
constexpr int kScanChunk = kThreads * kScanItems;      // items per block

__device__ __forceinline__ u64 block_exclusive_scan(u64 v, u64* total) {   // v: this thread's value; returns the sum of all lower threads'
    __shared__ u64 warp_tot[kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u64 o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    __syncthreads();   // protects warp_tot against the previous call's readers
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    u64 base = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        if (w < warp) base += warp_tot[w];
        all += warp_tot[w];
    }
    *total = all;
    return base + inc - v;
}
static __global__ void __launch_bounds__(kThreads) scan_chunk_totals_kernel(const u64* data, uint64_t n, u64* chunk_tot) {
    const uint64_t base = (uint64_t)blockIdx.x * kScanChunk + (uint64_t)threadIdx.x * kScanItems;
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) s += data[base + k];
    u64 total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) chunk_tot[blockIdx.x] = total;
}
static __global__ void __launch_bounds__(kThreads) scan_totals_kernel(u64* chunk_tot, uint64_t n_chunks) {   // one block: exclusive scan in place
    u64 carry = 0;
    for (uint64_t t0 = 0; t0 < n_chunks; t0 += kThreads) {
        const uint64_t i = t0 + threadIdx.x;
        const u64 v = i < n_chunks ? chunk_tot[i] : 0;
        u64 total;
        const u64 ex = block_exclusive_scan(v, &total);
        if (i < n_chunks) chunk_tot[i] = carry + ex;
        carry += total;
    }
}
static __global__ void __launch_bounds__(kThreads) scan_apply_kernel(u64* data, uint64_t n, const u64* chunk_prefix) {
    const uint64_t base = (uint64_t)blockIdx.x * kScanChunk + (uint64_t)threadIdx.x * kScanItems;
    u64 v[kScanItems], s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = base + k < n ? data[base + k] : 0;
        s += v[k];
    }
    u64 total;
    u64 run = chunk_prefix[blockIdx.x] + block_exclusive_scan(s, &total);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        run += v[k];
        if (base + k < n) data[base + k] = run;
    }
}

// scratch: at least scan_chunks(n) elements
inline uint64_t scan_chunks(uint64_t n) { return (n + kScanChunk - 1) / kScanChunk; }
inline void inclusive_scan(cudaStream_t stream, u64* data, uint64_t n, u64* scratch) {
    const uint64_t n_chunks = scan_chunks(n);
    scan_chunk_totals_kernel<<<(unsigned)n_chunks, kThreads, 0, stream>>>(data, n, scratch);
    scan_totals_kernel<<<1, kThreads, 0, stream>>>(scratch, n_chunks);
    scan_apply_kernel<<<(unsigned)n_chunks, kThreads, 0, stream>>>(data, n, scratch);
}
}}  // namespace zk::scan
