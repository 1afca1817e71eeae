// devrounds_launch.cuh -- the device-resident round loop (devrounds.cuh) under the CUDA execution policy, its kernel and
// the dispatch over (P, D).  Instantiated once per field in devrounds_fid{0,1,2}.cu so the fields compile in parallel.
#pragma once
#include "../../include/zk_sumcheck.h"
#include "engine.h"
#include "kernels.cuh"
#include "devrounds.cuh"

namespace zk {

#ifndef ZK_DEV_MIN_BLOCKS
#define ZK_DEV_MIN_BLOCKS 1   // resident blocks per SM the kernel is compiled for (1: the whole register file per block)
#endif
constexpr unsigned long long kDevSpinLimitNs = 5ull * 1000 * 1000 * 1000;     // arrivals / release inside one GPU
constexpr unsigned long long kDevPeerLimitNs = 60ull * 1000 * 1000 * 1000;    // a peer rank may still be uploading its tables

__device__ __forceinline__ unsigned long long dev_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct DevCudaExec {
    uint32_t wk_a, wk_b;   // this lane's routing words for the warp-wide Keccak (dev_transcript.cuh)
    uint32_t* sflag;       // shared word: outcome of a wait, broadcast to the block
    uint32_t fail_code;
    __device__ __forceinline__ DevCudaExec(uint32_t* flag) : wk_a(kWkA[threadIdx.x & 31u]), wk_b(kWkB[threadIdx.x & 31u]), sflag(flag), fail_code(0) {}
    __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31u); }
    __device__ __forceinline__ int warp() const { return (int)(threadIdx.x >> 5); }
    __device__ __forceinline__ uint32_t lanes() const { return 32u; }
    __device__ __forceinline__ unsigned long long now_ns() const { return dev_now_ns(); }
    __device__ __forceinline__ int tid() const { return (int)threadIdx.x; }
    __device__ __forceinline__ int nthreads() const { return (int)blockDim.x; }
    __device__ __forceinline__ uint32_t bid() const { return blockIdx.x; }
    __device__ __forceinline__ uint32_t nblocks() const { return gridDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ uint32_t failure() const { return fail_code; }
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
    template <int NC> __device__ __forceinline__ void column_sums(const uint32_t (&col)[NC], unsigned long long* tot) const {
        block_column_sums<NC>(col, tot);
    }
    // tables: another block may have written the entry a round ago -> L2-only loads
    __device__ __forceinline__ Fe load(const Fe* p) const { return ld256_cg(p); }
    __device__ __forceinline__ void store(Fe* p, const Fe& v) const { st256(p, v); }
    __device__ __forceinline__ void prefetch(const Fe* p) const { prefetch_l2(p); }
    __device__ __forceinline__ uint32_t load_word(const uint32_t* p) const { return __ldcg(p); }
    __device__ __forceinline__ void store_word(uint32_t* p, uint32_t v) const { __stcg(p, v); }
    __device__ __forceinline__ void grid_add(unsigned long long* p, unsigned long long v) const { atomicAdd(p, v); }
    __device__ __forceinline__ unsigned long long grid_take(unsigned long long* p) const {
        const unsigned long long v = __ldcg(p);
        __stcg(p, 0ull);
        return v;
    }
    // every thread's table stores and RED contributions are ordered before the arrival
    __device__ __forceinline__ void arrive(DevGlobal* g) const {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(&g->arrive, 1u);
    }
    __device__ __forceinline__ bool wait_arrivals(DevGlobal* g, uint32_t target) {
        if (threadIdx.x == 0) {
            uint32_t ok = 1;
            const unsigned long long t0 = dev_now_ns();
            for (uint32_t spins = 0; ld_acquire_gpu(&g->arrive) < target; ++spins) {
                if ((spins & 0xfffu) == 0xfffu && dev_now_ns() - t0 > kDevSpinLimitNs) { ok = 0; break; }
            }
            *sflag = ok;
        }
        __syncthreads();
        const bool ok = *sflag != 0;
        __syncthreads();
        if (!ok) fail_code = kDevTimeoutArrive;
        return ok;
    }
    __device__ __forceinline__ void release(DevGlobal* g, uint32_t round) const {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) st_release_gpu(&g->release, round);
    }
    __device__ __forceinline__ bool wait_release(DevGlobal* g, uint32_t round) {
        if (threadIdx.x == 0) {
            uint32_t ok = 1;
            const unsigned long long t0 = dev_now_ns();
            for (uint32_t spins = 0; ld_acquire_gpu(&g->release) < round; ++spins) {
                __nanosleep(40);   // up to ~150 blocks poll this word: keep the L2 slice free for the leader's release
                if ((spins & 0xffu) == 0xffu) {
                    if (ld_volatile_u32(&g->abort_) != 0 || dev_now_ns() - t0 > kDevSpinLimitNs) { ok = 0; break; }
                }
            }
            *sflag = ok;
        }
        __syncthreads();
        const bool ok = *sflag != 0;
        __syncthreads();
        if (!ok) fail_code = kDevTimeoutRelease;
        return ok;
    }
    // peer exchange: plain stores into the peer-mapped slot, system-scope fence, then the sequence word
    __device__ __forceinline__ void store_peer(Fe* p, const Fe& v) const {
        st256(p, v);
        __threadfence_system();
    }
    __device__ __forceinline__ void publish_peer(uint32_t* seq, uint32_t v) const {
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t*>(seq) = v;
    }
    __device__ __forceinline__ bool wait_peers(PeerSlot* mine, uint32_t world, uint32_t xs) {
        int ok = 1;
        if (threadIdx.x < world) {
            const unsigned long long t0 = dev_now_ns();
            for (uint32_t spins = 0; ld_volatile_u32(&mine[threadIdx.x].seq) != xs; ++spins) {
                if ((spins & 0xfffu) == 0xfffu && dev_now_ns() - t0 > kDevPeerLimitNs) { ok = 0; break; }
            }
            __threadfence_system();
        }
        ok = __syncthreads_and(ok);
        if (!ok) fail_code = kDevTimeoutPeer;
        return ok != 0;
    }
    __device__ __forceinline__ Fe load_peer(const Fe* p) const {
        Fe r;
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "l"(p) : "memory");
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(reinterpret_cast<const char*>(p) + 16) : "memory");
        return r;
    }
    // end of the launch: leave the global state zero for the next one (after a failure the host re-zeroes it)
    __device__ __forceinline__ void rearm(DevGlobal* g, bool failed) const {
        if (threadIdx.x == 0) {
            if (failed) {
                *reinterpret_cast<volatile uint32_t*>(&g->abort_) = 1u;
            } else {
                g->arrive = 0;
                g->release = 0;
            }
            __threadfence();
        }
    }
    __device__ __forceinline__ void publish(uint32_t* seq, uint32_t v) const {
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t*>(seq) = v;
    }
};

// The persistent grid: every block runs the rounds it has work in.  Must be launched cooperatively when gridDim.x > 1
// (all blocks co-resident: the barrier spins).
template <int FID, int P, int D, int NLIN>
__global__ void __launch_bounds__(kThreads, ZK_DEV_MIN_BLOCKS) sumcheck_rounds_kernel(const __grid_constant__ DevArgs a) {
    __shared__ DevShared sh;
    __shared__ uint32_t flag;
    DevCudaExec ex(&flag);
    DevRounds<FID, P, D, NLIN, DevCudaExec> blk(a, sh, ex);
    blk.init();
    while (blk.step()) {
    }
}

// launches the loop on `grid` blocks; max_blocks (out, may be null): co-resident capacity of this kernel on the device
template <int FID> int launch_dev_rounds_pd(zk_ctx* ctx, int P, int D, int nlin, const DevArgs& a, int grid, int* max_blocks);

#ifdef ZK_INSTANTIATE_DEVROUNDS
template <int FID, int P, int D, int NLIN> static int launch_dev_rounds_one(zk_ctx* ctx, const DevArgs& a, int grid, int* max_blocks) {
    auto kernel = sumcheck_rounds_kernel<FID, P, D, NLIN>;
    if (max_blocks) {
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
        if (e != cudaSuccess || per_sm < 1) {
            ctx->err = std::string("occupancy query of the round-loop kernel: ") + cudaGetErrorString(e);
            return ZK_ERR_CUDA;
        }
        *max_blocks = per_sm * ctx->sm_count;
        return ZK_OK;
    }
    cudaError_t e;
    if (grid > 1) {
        void* params[] = {const_cast<DevArgs*>(&a)};
        e = cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(kThreads), params, 0, ctx->stream);
    } else {
        kernel<<<1, kThreads, 0, ctx->stream>>>(a);
        e = cudaGetLastError();
    }
    ctx->launches++;
    if (e != cudaSuccess) {
        ctx->err = std::string("round-loop kernel launch: ") + cudaGetErrorString(e);
        return ZK_ERR_CUDA;
    }
    return ZK_OK;
}
template <int FID> int launch_dev_rounds_pd(zk_ctx* ctx, int P, int D, int nlin, const DevArgs& a, int grid, int* max_blocks) {
    if (nlin == 1 && P == 1 && D == 2) return launch_dev_rounds_one<FID, 1, 2, 1>(ctx, a, grid, max_blocks);
#define ZK_CASE(PP, DD) if (nlin == 0 && P == PP && D == DD) return launch_dev_rounds_one<FID, PP, DD, 0>(ctx, a, grid, max_blocks);
    ZK_CASE(1, 1) ZK_CASE(1, 2) ZK_CASE(2, 2) ZK_CASE(1, 3) ZK_CASE(2, 3) ZK_CASE(3, 2) ZK_CASE(4, 2)
#undef ZK_CASE
    ctx->err = "unsupported (P, D) for the device-resident round loop";
    return ZK_ERR_ARG;
}
template int launch_dev_rounds_pd<ZK_INSTANTIATE_DEVROUNDS>(zk_ctx*, int, int, int, const DevArgs&, int, int*);
#endif

}  // namespace zk
