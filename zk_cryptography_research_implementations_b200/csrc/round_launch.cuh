// round_launch.cuh -- dispatch of the round kernels over (P, D).  Instantiated once per field in
// round_fid{0,1,2}_{evals,fold}.cu so the heavy kernels compile in parallel.
#pragma once
#include "../../include/zk_sumcheck.h"
#include "engine.h"
#include "kernels.cuh"

namespace zk {

inline int launch_grid(const zk_ctx* ctx, uint64_t work, int blocks_per_sm) {
    uint64_t blocks = (work + kThreads - 1) / kThreads;
    uint64_t cap = (uint64_t)ctx->sm_count * blocks_per_sm;
    if (blocks > cap) blocks = cap;
    if (ctx->grid_cap > 0 && blocks > (uint64_t)ctx->grid_cap) blocks = ctx->grid_cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}
inline int launch_check(zk_ctx* ctx) {
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        ctx->err = std::string("kernel launch: ") + cudaGetErrorString(e);
        return ZK_ERR_CUDA;
    }
    return ZK_OK;
}
// where the next round kernel publishes: the shared mailbox slot of this rank (sharded round, double-buffered by
// sequence parity) or the context's own mailbox
inline ReduceScratch reduce_scratch(zk_ctx* ctx, bool shared) {
    if (shared) {
        unsigned seq = ++ctx->xmail_seq;
        ctx->exchange_pending = true;
        return ReduceScratch{ctx->gacc, ctx->ticket, ctx->xmail_dev + (size_t)(seq & 1u) * ctx->world + ctx->rank, seq};
    }
    unsigned seq = ++ctx->mail_seq;
    ctx->exchange_pending = false;
    return ReduceScratch{ctx->gacc, ctx->ticket, ctx->mail_dev, seq};
}
inline int unsupported_pd(zk_ctx* ctx) {
    ctx->err = "unsupported (P, D): supported are (1,1) (1,2) (2,2) (3,2) (4,2) (1,3) (2,3), and (1,2) with one linear table";
    return ZK_ERR_ARG;
}
// resident blocks per SM for the round kernels (register-limited; see profiles/)
#ifndef ZK_ROUND_BPS
#define ZK_ROUND_BPS 4
#endif
inline int round_blocks_per_sm(int tables) { return (tables >= 4) ? 2 : ZK_ROUND_BPS; }

#define ZK_PD_CASES ZK_CASE(1, 1) ZK_CASE(1, 2) ZK_CASE(2, 2) ZK_CASE(1, 3) ZK_CASE(2, 3) ZK_CASE(3, 2) ZK_CASE(4, 2)

template <int FID> int launch_round_evals_pd(zk_ctx* ctx, const TablePtrs& tp, int P, int D, int nlin, uint64_t half, bool shared, bool skip1);
template <int FID> int launch_fold_evals_pd(zk_ctx* ctx, const TablePtrs& tp, int P, int D, int nlin, uint64_t q, const FoldTable& ft, bool skip1, bool shared);

#ifdef ZK_INSTANTIATE_ROUND_EVALS
// skip1 is honoured for the GKR phase shape only (one product + one linear table); see round_evals_skip1_supported
template <int FID> int launch_round_evals_pd(zk_ctx* ctx, const TablePtrs& tp, int P, int D, int nlin, uint64_t half, bool shared, bool skip1) {
    if (nlin != 0 && !(P == 1 && D == 2 && nlin == 1)) return unsupported_pd(ctx);
    if (skip1 && nlin != 1) return unsupported_pd(ctx);
    ReduceScratch rs = reduce_scratch(ctx, shared);
    int grid = launch_grid(ctx, half, round_blocks_per_sm(P * D + nlin));
    if (nlin == 1) {
        if (skip1) round_evals_kernel<FID, 1, 2, 1, true><<<grid, kThreads, 0, ctx->stream>>>(tp, half, rs);
        else round_evals_kernel<FID, 1, 2, 1><<<grid, kThreads, 0, ctx->stream>>>(tp, half, rs);
        return launch_check(ctx);
    }
#define ZK_CASE(PP, DD)                                                                    \
    if (P == PP && D == DD) {                                                              \
        round_evals_kernel<FID, PP, DD><<<grid, kThreads, 0, ctx->stream>>>(tp, half, rs); \
        return launch_check(ctx);                                                          \
    }
    ZK_PD_CASES
#undef ZK_CASE
    return unsupported_pd(ctx);
}
template int launch_round_evals_pd<ZK_INSTANTIATE_ROUND_EVALS>(zk_ctx*, const TablePtrs&, int, int, int, uint64_t, bool, bool);
#endif

#ifdef ZK_INSTANTIATE_FOLD_EVALS
template <int FID> int launch_fold_evals_pd(zk_ctx* ctx, const TablePtrs& tp, int P, int D, int nlin, uint64_t q, const FoldTable& ft, bool skip1, bool shared) {
    if (nlin != 0 && !(P == 1 && D == 2 && nlin == 1)) return unsupported_pd(ctx);
    ReduceScratch rs = reduce_scratch(ctx, shared);
    int grid = launch_grid(ctx, q, round_blocks_per_sm(P * D + nlin));
    if (nlin == 1) {
        if (skip1) fold_evals_kernel<FID, 1, 2, true, 1><<<grid, kThreads, 0, ctx->stream>>>(tp, q, ft, rs);
        else fold_evals_kernel<FID, 1, 2, false, 1><<<grid, kThreads, 0, ctx->stream>>>(tp, q, ft, rs);
        return launch_check(ctx);
    }
#define ZK_CASE(PP, DD)                                                                                     \
    if (P == PP && D == DD) {                                                                               \
        if (skip1) fold_evals_kernel<FID, PP, DD, true><<<grid, kThreads, 0, ctx->stream>>>(tp, q, ft, rs); \
        else fold_evals_kernel<FID, PP, DD, false><<<grid, kThreads, 0, ctx->stream>>>(tp, q, ft, rs);      \
        return launch_check(ctx);                                                                           \
    }
    ZK_PD_CASES
#undef ZK_CASE
    return unsupported_pd(ctx);
}
template int launch_fold_evals_pd<ZK_INSTANTIATE_FOLD_EVALS>(zk_ctx*, const TablePtrs&, int, int, int, uint64_t, const FoldTable&, bool, bool);
#endif

}  // namespace zk
