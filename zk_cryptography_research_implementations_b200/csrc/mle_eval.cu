// mle_eval.cu -- MultilinearPolynomial::evaluate (polynomials/src/multilinear/evaluation_form.rs:21-33).
//
// The reference binds the n variables one after the other: n partial_evaluate passes, N - 1 modular multiplications,
// ~3N elements of memory traffic and a clone of the table first.  All n challenges are known up front, so the same
// value is the inner product
//
//      f(r) = sum_i eq(r, i) T[i],      eq(r, i) = prod_v (i_v ? r_v : 1 - r_v),
//
// and eq(r, .) is an outer product: with the index split as i = (i_hi, i_lo), eq(r, i) = EH[i_hi] EL[i_lo].  One pass
// over the table, one UNREDUCED 256x256-bit multiply-accumulate per entry (64 IMAD.WIDE instead of the 7/8 x 84 of the
// three-variables-per-pass fold), no scratch table:
//
//      f(r) = sum_{i_hi} EH[i_hi] * ( sum_{i_lo} T[i_hi, i_lo] EL[i_lo] )
//
// A warp owns a row (fixed i_hi, 2^lo_bits consecutive entries, EL cache-resident): every lane accumulates its entries'
// products as a 17-limb integer, the warp's columns are added exactly (REDUX), ONE Montgomery reduction gives the row's
// inner sum u, and u * EH[i_hi] goes -- unreduced again -- into the warp's second-level accumulator.  The grid-wide finish
// is the exact column sum of the round kernels (kernels.cuh) with one last Montgomery reduction.  Every reduction returns
// the canonical representative, so the result equals the reference's n folds limb for limb (tests/test_gpu_parity*.py).
#include <string>
#include <vector>

#include "../../include/zk_sumcheck.h"
#include "engine.h"
#include "eq_tables.cuh"
#include "internal.h"

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
constexpr uint32_t kEvalLoBits = 11;   // EL has 2^11 entries = 64 KiB: stays in L1 beside the streamed table
inline uint32_t ilog2(uint64_t n) { uint32_t k = 0; while (n >>= 1) ++k; return k; }
inline int grid_of(const zk_ctx* ctx, uint64_t work, int bps) {
    uint64_t blocks = (work + kThreads - 1) / kThreads, cap = (uint64_t)ctx->sm_count * bps;
    if (blocks > cap) blocks = cap;
    if (ctx->grid_cap > 0 && blocks > (uint64_t)ctx->grid_cap) blocks = ctx->grid_cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

// what publish_round (kernels.cuh) needs: one evaluation, 17 columns
template <int FID> struct InnerAcc {
    static constexpr int NC = 17, NE = 1;
    uint32_t acc[17];
    __device__ __forceinline__ void columns(uint32_t (&col)[NC]) const {
#pragma unroll
        for (int k = 0; k < 17; ++k) col[k] = acc[k];
    }
    __device__ __forceinline__ static void finalize(Fe& out, int, const unsigned long long* tot) {
        uint32_t limbs[17];
        unsigned long long c = 0;
#pragma unroll
        for (int k = 0; k < 17; ++k) {
            c += tot[k];
            limbs[k] = (uint32_t)c;
            c >>= 32;
        }
        Fp<FID>::redc_wide(out, limbs);
    }
};

template <int FID>
__global__ void __launch_bounds__(kThreads, 2) mle_inner_kernel(const Fe* tab, const Fe* eq_hi, const Fe* eq_lo, uint32_t lo_bits, uint64_t n_rows, ReduceScratch rs) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = kThreads / 32;
    const uint64_t row_len = 1ull << lo_bits;
    InnerAcc<FID> second;
#pragma unroll
    for (int k = 0; k < 17; ++k) second.acc[k] = 0;
    for (uint64_t row = (uint64_t)blockIdx.x * wpb + warp; row < n_rows; row += (uint64_t)gridDim.x * wpb) {
        const Fe* rp = tab + (row << lo_bits);
        uint32_t acc[17];
#pragma unroll
        for (int k = 0; k < 17; ++k) acc[k] = 0;
        uint64_t j = lane;
        for (; j + 96 < row_len; j += 128) {   // four entries per lane in flight
            Fe t0 = ld256(rp + j), t1 = ld256(rp + j + 32), t2 = ld256(rp + j + 64), t3 = ld256(rp + j + 96);
            Fe e0 = ld256_ca(eq_lo + j), e1 = ld256_ca(eq_lo + j + 32), e2 = ld256_ca(eq_lo + j + 64), e3 = ld256_ca(eq_lo + j + 96);
            Fp<FID>::mul_acc(acc, t0, e0);
            Fp<FID>::mul_acc(acc, t1, e1);
            Fp<FID>::mul_acc(acc, t2, e2);
            Fp<FID>::mul_acc(acc, t3, e3);
        }
        for (; j < row_len; j += 32) {
            Fe t = ld256(rp + j), e = ld256_ca(eq_lo + j);
            Fp<FID>::mul_acc(acc, t, e);
        }
        // the warp's exact column totals (every lane gets them), carried into 17 limbs: < 2^11 products of canonical elements
        uint32_t limbs[17];
        unsigned long long c = 0;
#pragma unroll
        for (int k = 0; k < 17; ++k) {
            const unsigned lo = __reduce_add_sync(0xffffffffu, acc[k] & 0xffffu);
            const unsigned hi = __reduce_add_sync(0xffffffffu, acc[k] >> 16);
            c += (unsigned long long)lo + ((unsigned long long)hi << 16);
            limbs[k] = (uint32_t)c;
            c >>= 32;
        }
        Fe u, eh = ld256_ca(eq_hi + row);
        Fp<FID>::redc_wide(u, limbs);               // the row's inner sum, canonical Montgomery form
        Fp<FID>::mul_acc(second.acc, u, eh);        // every lane holds the same value; only lane 0's copy is counted below
    }
    if (lane != 0) {
#pragma unroll
        for (int k = 0; k < 17; ++k) second.acc[k] = 0;
    }
    publish_round(second, rs);
}
}  // namespace

// eq(r[0..k), .) for k variables into `out` (2^k entries), through scratch half tables when k is large
static int build_eq(zk_ctx* ctx, const HFe* r, uint32_t k, Fe* out, Fe* half_a, Fe* half_b) {
    const HostField& f = ctx->field;
    auto fill = [&](EqHalfArgs& a, int h, const HFe* rr, uint32_t bits, Fe* dst) {
        a.out[h] = dst;
        a.bits[h] = bits;
        for (uint32_t v = 0; v < bits; ++v) {
            HFe omr = f.sub(f.one(), rr[v]);
            memcpy(a.factors[h][2 * v].v, omr.l, 32);
            memcpy(a.factors[h][2 * v + 1].v, rr[v].l, 32);
        }
    };
    HFe one = f.one();
    EqHalfArgs a;
    memcpy(a.scale.v, one.l, 32);
    if (k <= (uint32_t)kEqHalfBits) {          // directly: one product of k factors per entry (half 1 is empty: one entry)
        fill(a, 0, r, k, out);
        fill(a, 1, r, 0, half_b);
        dim3 grid((unsigned)grid_of(ctx, 1ull << k, 4), 2);
        ZK_FID_SWITCH(ctx, (eq_halves_kernel<FID><<<grid, kThreads, 0, ctx->stream>>>(a)));
        ctx->launches++;
        ZK_CUDA(cudaGetLastError());
        return ZK_OK;
    }
    const uint32_t ka = k / 2, kb = k - ka;
    if (kb > (uint32_t)kEqHalfBits) return fail(ctx, ZK_ERR_ARG, "table too long for evaluate");
    fill(a, 0, r, ka, half_a);
    fill(a, 1, r + ka, kb, half_b);
    dim3 grid((unsigned)grid_of(ctx, 1ull << kb, 4), 2);
    ZK_FID_SWITCH(ctx, (eq_halves_kernel<FID><<<grid, kThreads, 0, ctx->stream>>>(a)));
    ZK_FID_SWITCH(ctx, (eq_outer2_kernel<FID, false><<<grid_of(ctx, 1ull << k, 4), kThreads, 0, ctx->stream>>>(out, half_a, half_b, nullptr, nullptr, kb, 1ull << k)));
    ctx->launches += 2;
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}

// evaluate at all n = log2(len) challenges, n >= 1: the inner-product form
static int evaluate_inner(zk_ctx* ctx, const zk_table* t, const uint64_t* values, uint32_t n, uint64_t out[4]) {
    const uint32_t kl = n < kEvalLoBits ? n : kEvalLoBits, kh = n - kl;
    // scratch: EL | EH | two half tables for a long EH
    const uint64_t n_lo = 1ull << kl, n_hi = 1ull << kh, n_half = 1ull << ((kh + 1) / 2);
    int rc = ensure_scratch(ctx, (size_t)(n_lo + n_hi + 2 * n_half + 2) * sizeof(Fe));
    if (rc) return rc;
    Fe* el = (Fe*)ctx->scratch;
    Fe* eh = el + n_lo;
    Fe* ha = eh + n_hi;
    Fe* hb = ha + n_half + 1;
    const HFe* r = reinterpret_cast<const HFe*>(values);
    if ((rc = build_eq(ctx, r + kh, kl, el, ha, hb))) return rc;     // the trailing variables index inside a row
    if ((rc = build_eq(ctx, r, kh, eh, ha, hb))) return rc;          // the leading variables pick the row (kh == 0: EH = [1])
    unsigned seq = ++ctx->mail_seq;
    ReduceScratch rs{ctx->gacc, ctx->ticket, ctx->mail_dev, seq};
    const uint64_t rows = n_hi;
    uint64_t blocks = (rows + (kThreads / 32) - 1) / (kThreads / 32), cap = (uint64_t)ctx->sm_count * 2;
    if (blocks > cap) blocks = cap;
    if (ctx->grid_cap > 0 && blocks > (uint64_t)ctx->grid_cap) blocks = ctx->grid_cap;
    ZK_FID_SWITCH(ctx, (mle_inner_kernel<FID><<<(int)blocks, kThreads, 0, ctx->stream>>>(t->d, eh, el, kl, rows, rs)));
    ctx->launches++;
    ZK_CUDA(cudaGetLastError());
    HFe res;
    rc = fetch_result(ctx, &res, 1);
    if (rc) return rc;
    memcpy(out, res.l, 32);
    return ZK_OK;
}

extern "C" int zk_mle_evaluate(zk_ctx* ctx, const zk_table* t, const uint64_t* values, uint32_t n_values, uint64_t out[4]) {
    uint32_t nvars = ilog2(t->len);
    if (n_values > nvars) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");  // fold of a 1-entry table
    const char* knob = getenv("ZKB200_EVAL_FOLDS");   // test hook: "1" forces the fold passes for a full evaluation too
    if (n_values == nvars && nvars >= 1 && !(knob && knob[0] == '1')) return evaluate_inner(ctx, t, values, nvars, out);
    // a prefix of the variables (the result is entry 0 of the partially folded table), or no variable at all:
    // bind three variables per pass (fold_multi_kernel), out of place
    const Fe* cur = t->d;
    uint64_t len = t->len;
    uint32_t done = 0;
    if (n_values > 0) {
        int rc = ensure_scratch(ctx, (size_t)(len / 2) * sizeof(Fe));
        if (rc) return rc;
    }
    while (done < n_values) {
        uint32_t k = n_values - done >= 3 ? 3 : n_values - done;
        FoldTables3 fts;
        for (uint32_t i = 0; i < k; ++i) {
            HFe rr;
            memcpy(rr.l, values + 4 * (done + i), 32);
            fts.t[i] = make_fold_table(ctx->field, rr);
        }
        uint64_t m = len >> k;
        Fe* dst = (Fe*)ctx->scratch;
        int grid = grid_of(ctx, m, 4);
        if (k == 3) { ZK_FID_SWITCH(ctx, (fold_multi_kernel<FID, 3><<<grid, kThreads, 0, ctx->stream>>>(cur, dst, m, fts))); }
        else if (k == 2) { ZK_FID_SWITCH(ctx, (fold_multi_kernel<FID, 2><<<grid, kThreads, 0, ctx->stream>>>(cur, dst, m, fts))); }
        else { ZK_FID_SWITCH(ctx, (fold_multi_kernel<FID, 1><<<grid, kThreads, 0, ctx->stream>>>(cur, dst, m, fts))); }
        ctx->launches++;
        ZK_CUDA(cudaGetLastError());
        cur = dst;
        len = m;
        done += k;
    }
    ZK_CUDA(cudaMemcpyAsync(out, cur, sizeof(Fe), cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}
