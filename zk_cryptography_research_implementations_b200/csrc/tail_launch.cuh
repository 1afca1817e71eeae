// tail_launch.cuh -- the single-block tail kernel (tail.cuh's body under the CUDA execution policy) and its
// dispatch over (P, D).  Instantiated once per field in tail_fid{0,1,2}.cu so the fields compile in parallel.
#pragma once
#include "../../include/zk_sumcheck.h"
#include "engine.h"
#include "kernels.cuh"
#include "tail.cuh"

namespace zk {

struct CudaExec {
    uint32_t wk_a, wk_b;   // this lane's routing words for the warp-wide Keccak (dev_transcript.cuh)
    __device__ __forceinline__ CudaExec() : wk_a(kWkA[threadIdx.x & 31u]), wk_b(kWkB[threadIdx.x & 31u]) {}
    __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31u); }
    __device__ __forceinline__ int warp() const { return (int)(threadIdx.x >> 5); }
    // Keccak-f[1600] on a state in shared memory, by the whole (converged) warp
    __device__ __forceinline__ void permute(uint64_t* s) const {
        __syncwarp();
        const int l = lane();
        const uint64_t a = l < 25 ? s[l] : 0ull;
        uint32_t lo = (uint32_t)a, hi = (uint32_t)(a >> 32);
        warp_keccak_f1600<uint32_t>(lo, hi, wk_a, wk_b);
        if (l < 25) s[l] = (uint64_t)lo | ((uint64_t)hi << 32);
        __syncwarp();
    }
    // Keccak-256 digest of a CLONE of the sponge (pad 0x01 .. 0x80, permute, first four words); s is left untouched
    __device__ __forceinline__ void finalize(const uint64_t* s, uint32_t pos, uint64_t* digest) const {
        __syncwarp();
        const int l = lane();
        uint64_t a = l < 25 ? s[l] : 0ull;
        if (l == (int)(pos >> 3)) a ^= 0x01ull << (8 * (pos & 7));
        if (l == 16) a ^= 0x8000000000000000ull;
        uint32_t lo = (uint32_t)a, hi = (uint32_t)(a >> 32);
        warp_keccak_f1600<uint32_t>(lo, hi, wk_a, wk_b);
        if (l < 4) digest[l] = (uint64_t)lo | ((uint64_t)hi << 32);
        __syncwarp();
    }
    __device__ __forceinline__ int tid() const { return (int)threadIdx.x; }
    __device__ __forceinline__ int nthreads() const { return (int)blockDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    template <int NC> __device__ __forceinline__ void column_sums(const uint32_t (&col)[NC], unsigned long long* tot) const {
        block_column_sums<NC>(col, tot);
    }
    __device__ __forceinline__ Fe load(const Fe* p) const { return ld256(p); }
    __device__ __forceinline__ void store(Fe* p, const Fe& v) const { st256(p, v); }
    __device__ __forceinline__ void publish(uint32_t* seq, uint32_t v) const {
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t*>(seq) = v;
    }
};

// One block; every remaining round of the sumcheck, the Fiat-Shamir transcript included.
template <int FID, int P, int D, int NLIN>
__global__ void __launch_bounds__(kThreads, 1) sumcheck_tail_kernel(const __grid_constant__ TailArgs a) {
    __shared__ TailShared sh;
    CudaExec ex;
    sumcheck_tail_body<FID, P, D, NLIN>(a, sh, ex);
}

template <int FID> int launch_tail_pd(zk_ctx* ctx, int P, int D, int nlin, const TailArgs& a);

#ifdef ZK_INSTANTIATE_TAIL
template <int FID> int launch_tail_pd(zk_ctx* ctx, int P, int D, int nlin, const TailArgs& a) {
    if (nlin == 1 && P == 1 && D == 2) {
        sumcheck_tail_kernel<FID, 1, 2, 1><<<1, kThreads, 0, ctx->stream>>>(a);
    }
#define ZK_CASE(PP, DD) else if (nlin == 0 && P == PP && D == DD) { sumcheck_tail_kernel<FID, PP, DD, 0><<<1, kThreads, 0, ctx->stream>>>(a); }
    ZK_CASE(1, 1) ZK_CASE(1, 2) ZK_CASE(2, 2) ZK_CASE(1, 3) ZK_CASE(2, 3) ZK_CASE(3, 2) ZK_CASE(4, 2)
#undef ZK_CASE
    else {
        ctx->err = "unsupported (P, D) for the device tail";
        return ZK_ERR_ARG;
    }
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        ctx->err = std::string("tail kernel launch: ") + cudaGetErrorString(e);
        return ZK_ERR_CUDA;
    }
    return ZK_OK;
}
template int launch_tail_pd<ZK_INSTANTIATE_TAIL>(zk_ctx*, int, int, int, const TailArgs&);
#endif

}  // namespace zk
