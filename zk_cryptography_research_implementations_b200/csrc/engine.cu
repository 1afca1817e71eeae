// engine.cu -- host side of libzkb200.so: context, tables, kernel dispatch, the round loops of the
// reference's provers, and the extern "C" boundary declared in include/zk_sumcheck.h.
//
// Mirrors (reference paths): polynomials/src/multilinear/evaluation_form.rs,
// polynomials/src/composed/{product,sum}_polynomial.rs,
// sumcheck_protocol/src/basic_sumcheck/prover.rs, sumcheck_protocol/src/gkr_sumcheck/sumcheck_gkr_protocol.rs.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <atomic>
#include <chrono>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/zk_sumcheck.h"
#include "engine.h"
#include "kernels.cuh"
#include "round_launch.cuh"
#include "devrounds_launch.cuh"
#include "internal.h"

using namespace zk;

// =================================================================================== helpers
#define ZK_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                   \
            return ZK_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

static inline bool is_pow2(uint64_t n) { return n && !(n & (n - 1)); }
static inline uint32_t ilog2(uint64_t n) { uint32_t k = 0; while (n >>= 1) ++k; return k; }

int fail(zk_ctx* ctx, int code, const char* msg) {
    ctx->err = msg;
    return code;
}

FoldTable make_fold_table(const HostField& f, const HFe& r_mont) {
    // tab[i] = r * 2^(32 i) mod p as PLAIN integers (see fp.cuh FoldScalar)
    FoldTable ft;
    HFe cur = f.from_mont(r_mont);
    HFe m232 = f.from_u64(1ull << 32);  // Montgomery form of 2^32: mul(x, m232) == x * 2^32 mod p
    for (int i = 0; i < 8; ++i) {
        memcpy(ft.w[i], cur.l, 32);
        cur = f.mul(cur, m232);
    }
    return ft;
}

static int grid_for(const zk_ctx* ctx, uint64_t work, int blocks_per_sm) {
    uint64_t blocks = (work + kThreads - 1) / kThreads;
    uint64_t cap = (uint64_t)ctx->sm_count * blocks_per_sm;
    if (blocks > cap) blocks = cap;
    if (ctx->grid_cap > 0 && blocks > (uint64_t)ctx->grid_cap) blocks = ctx->grid_cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

#define ZK_DISPATCH_FID(ctx, EXPR)                \
    switch ((ctx)->fid) {                         \
        case 0: { constexpr int FID = 0; EXPR; } break; \
        case 1: { constexpr int FID = 1; EXPR; } break; \
        case 2: { constexpr int FID = 2; EXPR; } break; \
        default: return fail(ctx, ZK_ERR_ARG, "unknown field id"); \
    }

static int post_launch(zk_ctx* ctx) {
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        ctx->err = std::string("kernel launch: ") + cudaGetErrorString(e);
        return ZK_ERR_CUDA;
    }
    return ZK_OK;
}

// round-kernel profiling (CUDA events on the launching stream)
static void prof_begin(zk_ctx* ctx) {
    if (!ctx->profiling) return;
    if (ctx->ev_used + 2 > ctx->events.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        ctx->events.push_back(a);
        ctx->events.push_back(b);
    }
    cudaEventRecord(ctx->events[ctx->ev_used], ctx->stream);
}
static void prof_end(zk_ctx* ctx, double bytes) {
    ctx->round_launches++;
    ctx->round_bytes += bytes;
    if (!ctx->profiling) return;
    cudaEventRecord(ctx->events[ctx->ev_used + 1], ctx->stream);
    ctx->ev_used += 2;
}
static void prof_collect(zk_ctx* ctx) {
    for (size_t i = 0; i + 1 < ctx->ev_used; i += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->events[i], ctx->events[i + 1]) == cudaSuccess) ctx->round_ms += ms;
    }
    ctx->ev_used = 0;
}

// =================================================================================== context
static int ctx_create(zk_ctx** out, int fid, int device, void* stream, bool own_stream) {
    if (!out || fid < 0 || fid >= ZKF_NUM_FIELDS) return ZK_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        fprintf(stderr, "zkb200: no CUDA device available -- this library has no CPU fallback\n");
        return ZK_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) return ZK_ERR_ARG;
    std::unique_ptr<zk_ctx> c(new zk_ctx(fid));
    zk_ctx* ctx = c.get();
    ctx->device = device;
    ZK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    ZK_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    if (own_stream) {
        ZK_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    } else {
        ctx->stream = (cudaStream_t)stream;
    }
    ctx->own_stream = own_stream;
    ctx->max_grid = ctx->sm_count * 8;
    if (const char* cap = getenv("ZKB200_GRID_CAP")) ctx->grid_cap = atoi(cap);
    ZK_CUDA(cudaMalloc(&ctx->gacc, (size_t)kMaxCols * sizeof(unsigned long long)));
    ZK_CUDA(cudaMemsetAsync(ctx->gacc, 0, (size_t)kMaxCols * sizeof(unsigned long long), ctx->stream));
    ZK_CUDA(cudaMalloc(&ctx->ticket, sizeof(unsigned)));
    ZK_CUDA(cudaMemsetAsync(ctx->ticket, 0, sizeof(unsigned), ctx->stream));
    ZK_CUDA(cudaHostAlloc(&ctx->mail_host, sizeof(Mailbox), cudaHostAllocMapped));
    memset(ctx->mail_host, 0, sizeof(Mailbox));
    ZK_CUDA(cudaHostGetDevicePointer((void**)&ctx->mail_dev, ctx->mail_host, 0));
    ZK_CUDA(cudaHostAlloc(&ctx->dev_host, sizeof(DevOut), cudaHostAllocMapped));
    memset(ctx->dev_host, 0, sizeof(DevOut));
    ZK_CUDA(cudaHostGetDevicePointer((void**)&ctx->dev_dev, ctx->dev_host, 0));
    ZK_CUDA(cudaMalloc(&ctx->dev_global, sizeof(DevGlobal)));
    ZK_CUDA(cudaMemsetAsync(ctx->dev_global, 0, sizeof(DevGlobal), ctx->stream));
    if (const char* tl = getenv("ZKB200_TAIL_LOG")) ctx->tail_log = atoi(tl);
    if (ctx->tail_log < 0 || ctx->tail_log > kDevMaxLog) ctx->tail_log = ctx->tail_log < 0 ? 0 : kDevMaxLog;
    {
        HFe cur = ctx->field.one(), m232 = ctx->field.from_u64(1ull << 32);
        for (int i = 0; i < 8; ++i) { ctx->pow32[i] = cur; cur = ctx->field.mul(cur, m232); }
    }
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = c.release();
    return ZK_OK;
}

extern "C" const char* zk_version(void) { return "zkb200 0.1 (sm_100a)"; }
extern "C" int zk_ctx_create(zk_ctx** out, int field_id, int device) { return ctx_create(out, field_id, device, nullptr, true); }
extern "C" int zk_ctx_create_on_stream(zk_ctx** out, int field_id, int device, void* stream) {
    return ctx_create(out, field_id, device, stream, false);
}
extern "C" void zk_ctx_destroy(zk_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    zk_comm_destroy(ctx);
    if (ctx->side_event) cudaEventDestroy(ctx->side_event);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->pool) cudaFree(ctx->pool);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    cudaFree(ctx->gacc);
    cudaFree(ctx->ticket);
    cudaFreeHost(ctx->mail_host);
    cudaFreeHost(ctx->dev_host);
    cudaFree(ctx->dev_global);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}
extern "C" const char* zk_last_error(const zk_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" int zk_ctx_synchronize(zk_ctx* ctx) {
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}
extern "C" int zk_ctx_set_profiling(zk_ctx* ctx, int on) { ctx->profiling = on != 0; return ZK_OK; }
extern "C" int zk_ctx_set_tail_log(zk_ctx* ctx, int tail_log) {
    if (tail_log < 0 || tail_log > kDevMaxLog) return fail(ctx, ZK_ERR_ARG, "tail_log must be in 0..32");
    ctx->tail_log = tail_log;
    return ZK_OK;
}
extern "C" int zk_ctx_get_tail_log(const zk_ctx* ctx) { return ctx->tail_log; }
extern "C" int zk_ctx_reset_stats(zk_ctx* ctx) {
    ctx->launches = ctx->round_launches = 0;
    ctx->round_ms = ctx->round_bytes = 0;
    ctx->ev_used = 0;
    return ZK_OK;
}
extern "C" int zk_ctx_get_stats(zk_ctx* ctx, uint64_t* launches, uint64_t* round_launches, double* round_ms, double* round_bytes) {
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    prof_collect(ctx);
    if (launches) *launches = ctx->launches;
    if (round_launches) *round_launches = ctx->round_launches;
    if (round_ms) *round_ms = ctx->round_ms;
    if (round_bytes) *round_bytes = ctx->round_bytes;
    return ZK_OK;
}

int ensure_scratch(zk_ctx* ctx, size_t bytes) {
    if (ctx->scratch_bytes >= bytes) return ZK_OK;
    if (ctx->scratch) {
        ZK_CUDA(cudaStreamSynchronize(ctx->stream));
        ZK_CUDA(cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
    }
    ZK_CUDA(cudaMalloc(&ctx->scratch, bytes));
    ctx->scratch_bytes = bytes;
    return ZK_OK;
}

int ensure_pool(zk_ctx* ctx, size_t bytes) {
    if (ctx->pool_bytes >= bytes) return ZK_OK;
    if (ctx->pool) {
        ZK_CUDA(cudaStreamSynchronize(ctx->stream));
        ZK_CUDA(cudaFree(ctx->pool));
        ctx->pool = nullptr;
        ctx->pool_bytes = 0;
    }
    ZK_CUDA(cudaMalloc(&ctx->pool, bytes));
    ctx->pool_bytes = bytes;
    return ZK_OK;
}

// =================================================================================== host field helpers
extern "C" int zk_fe_from_u64(int fid, uint64_t v, uint64_t out[4]) {
    if (fid < 0 || fid >= ZKF_NUM_FIELDS) return ZK_ERR_ARG;
    HostField f(fid);
    HFe r = f.from_u64(v);
    memcpy(out, r.l, 32);
    return ZK_OK;
}
#define ZK_HOST_UNARY(NAME, EXPR)                                              \
    extern "C" int NAME(int fid, const uint64_t in[4], uint64_t out[4]) {      \
        if (fid < 0 || fid >= ZKF_NUM_FIELDS) return ZK_ERR_ARG;               \
        HostField f(fid);                                                      \
        HFe a;                                                                 \
        memcpy(a.l, in, 32);                                                   \
        HFe r = EXPR;                                                          \
        memcpy(out, r.l, 32);                                                  \
        return ZK_OK;                                                          \
    }
ZK_HOST_UNARY(zk_fe_to_canonical, f.from_mont(a))
ZK_HOST_UNARY(zk_fe_from_canonical, f.to_mont(a))
#define ZK_HOST_BINARY(NAME, OP)                                                               \
    extern "C" int NAME(int fid, const uint64_t a_[4], const uint64_t b_[4], uint64_t out[4]) { \
        if (fid < 0 || fid >= ZKF_NUM_FIELDS) return ZK_ERR_ARG;                               \
        HostField f(fid);                                                                      \
        HFe a, b;                                                                              \
        memcpy(a.l, a_, 32);                                                                   \
        memcpy(b.l, b_, 32);                                                                   \
        HFe r = f.OP(a, b);                                                                    \
        memcpy(out, r.l, 32);                                                                  \
        return ZK_OK;                                                                          \
    }
ZK_HOST_BINARY(zk_fe_add, add)
ZK_HOST_BINARY(zk_fe_sub, sub)
ZK_HOST_BINARY(zk_fe_mul, mul)

// DenseUnivariatePolynomial::lagrange_interpolate on the nodes 0..n-1 (dense_univariate.rs:74-98) and
// ::evaluate (:57-68) -- the per-round host work of the product prover, exported for callers that keep
// the round loop (and the transcript) on their side.
extern "C" int zk_interpolate_evals(int fid, uint32_t n_evals, const uint64_t* evals, uint64_t* coeffs) {
    if (fid < 0 || fid >= ZKF_NUM_FIELDS || n_evals < 1 || n_evals > 16) return ZK_ERR_ARG;
    HostField f(fid);
    Interpolator ip(f, (int)n_evals - 1);
    ip.coefficients(reinterpret_cast<const HFe*>(evals), reinterpret_cast<HFe*>(coeffs));
    return ZK_OK;
}
extern "C" int zk_univariate_evaluate(int fid, const uint64_t* coeffs, uint32_t n, const uint64_t x[4], uint64_t out[4]) {
    if (fid < 0 || fid >= ZKF_NUM_FIELDS) return ZK_ERR_ARG;
    HostField f(fid);
    HFe xx;
    memcpy(xx.l, x, 32);
    HFe r = f.horner(reinterpret_cast<const HFe*>(coeffs), (int)n, xx);
    memcpy(out, r.l, 32);
    return ZK_OK;
}

// =================================================================================== transcript
extern "C" zk_transcript* zk_transcript_new(void) { return new zk_transcript(); }
extern "C" void zk_transcript_free(zk_transcript* t) { delete t; }
extern "C" void zk_transcript_append(zk_transcript* t, const uint8_t* data, size_t len) { t->t.append(data, len); }
extern "C" void zk_transcript_sample(zk_transcript* t, uint8_t out[32]) { t->t.sample(out); }
extern "C" void zk_transcript_challenge(zk_transcript* t, int fid, uint64_t out[4]) {
    HostField f(fid);
    HFe r = t->t.challenge(f);
    memcpy(out, r.l, 32);
}

// =================================================================================== tables
int table_alloc(zk_ctx* ctx, uint64_t n, zk_table** out) {
    std::unique_ptr<zk_table> t(new zk_table());
    t->len = n;
    t->cap = n;
    t->owned = true;
    ZK_CUDA(cudaSetDevice(ctx->device));
    ZK_CUDA(cudaMalloc(&t->d, (size_t)(n ? n : 1) * sizeof(Fe)));
    *out = t.release();
    return ZK_OK;
}
extern "C" int zk_table_upload(zk_ctx* ctx, const uint64_t* limbs, uint64_t n, zk_table** out) {
    if (!is_pow2(n)) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");
    int rc = table_alloc(ctx, n, out);
    if (rc) return rc;
    ZK_CUDA(cudaMemcpyAsync((*out)->d, limbs, (size_t)n * sizeof(Fe), cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}
extern "C" int zk_table_regenerate(zk_ctx* ctx, zk_table* t, uint64_t seed, uint64_t table_id, uint64_t n, uint64_t first, uint64_t step) {
    if (!is_pow2(n)) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");
    if (n > t->cap) return fail(ctx, ZK_ERR_ARG, "regenerate: table capacity too small");
    t->len = n;
    ZK_DISPATCH_FID(ctx, (generate_kernel<FID><<<grid_for(ctx, n, 8), kThreads, 0, ctx->stream>>>(t->d, n, seed, table_id, first, step)));
    return post_launch(ctx);
}
extern "C" int zk_table_generate(zk_ctx* ctx, uint64_t seed, uint64_t table_id, uint64_t n, uint64_t first, uint64_t step, zk_table** out) {
    if (!is_pow2(n)) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");
    int rc = table_alloc(ctx, n, out);
    if (rc) return rc;
    return zk_table_regenerate(ctx, *out, seed, table_id, n, first, step);
}
extern "C" int zk_table_wrap(zk_ctx* ctx, void* device_ptr, uint64_t n, zk_table** out) {
    if (!is_pow2(n)) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");
    if (((uintptr_t)device_ptr & 31) != 0) return fail(ctx, ZK_ERR_ARG, "table memory must be 32-byte aligned");
    zk_table* t = new zk_table();
    t->d = (Fe*)device_ptr;
    t->len = t->cap = n;
    t->owned = false;
    *out = t;
    return ZK_OK;
}
extern "C" int zk_table_clone(zk_ctx* ctx, const zk_table* src, zk_table** out) {
    int rc = table_alloc(ctx, src->len, out);
    if (rc) return rc;
    ZK_CUDA(cudaMemcpyAsync((*out)->d, src->d, (size_t)src->len * sizeof(Fe), cudaMemcpyDeviceToDevice, ctx->stream));
    return ZK_OK;
}
extern "C" int zk_table_download(zk_ctx* ctx, const zk_table* t, uint64_t* out_limbs) {
    ZK_CUDA(cudaMemcpyAsync(out_limbs, t->d, (size_t)t->len * sizeof(Fe), cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}
// asynchronous refill of an existing table from (pinned) host memory -- the per-step input copy of an
// end-to-end call; ordered on the context's stream, no synchronisation
extern "C" int zk_table_upload_into(zk_ctx* ctx, zk_table* t, const uint64_t* limbs, uint64_t n) {
    if (!is_pow2(n)) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");
    if (n > t->cap) return fail(ctx, ZK_ERR_ARG, "upload_into: table capacity too small");
    t->len = n;
    ZK_CUDA(cudaMemcpyAsync(t->d, limbs, (size_t)n * sizeof(Fe), cudaMemcpyHostToDevice, ctx->stream));
    return ZK_OK;
}
extern "C" int zk_pinned_alloc(size_t bytes, void** out) { return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? ZK_OK : ZK_ERR_CUDA; }
extern "C" void zk_pinned_free(void* p) { cudaFreeHost(p); }
extern "C" uint64_t zk_table_len(const zk_table* t) { return t->len; }
extern "C" void* zk_table_device_ptr(const zk_table* t) { return t->d; }
extern "C" void zk_table_free(zk_ctx* ctx, zk_table* t) {
    if (!t) return;
    if (t->owned && t->d) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        cudaFree(t->d);
    }
    delete t;
}

// =================================================================================== kernel launchers
static bool supported_pd(uint32_t P, uint32_t D);
extern const char* const kUnsupportedPdMsg;
namespace zk {

int wait_mailbox(zk_ctx* ctx, const volatile Mailbox* box, unsigned seq, bool own) { return wait_seq(ctx, &box->seq, seq, own); }

int wait_seq(zk_ctx* ctx, const volatile unsigned* word, unsigned seq, bool own) {
    auto t0 = std::chrono::steady_clock::now();
    for (uint64_t spins = 0;; ++spins) {
        if (*word == seq) break;
        if ((spins & 0x3ff) == 0x3ff) {
            cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) {
                ctx->err = std::string("round kernel failed: ") + cudaGetErrorString(q);
                return ZK_ERR_CUDA;
            }
            if (q == cudaSuccess && own && *word != seq) {
                // the stream drained: the write must have landed; re-read once before giving up
                if (*word == seq) break;
                return fail(ctx, ZK_ERR_CUDA, "round kernel finished without publishing its result");
            }
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120))
                return fail(ctx, ZK_ERR_CUDA, "timed out waiting for a round result (peer rank lost?)");
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return ZK_OK;
}

// wait for the round kernel (spin on the mailbox sequence number) and copy its NE published elements out
int fetch_result(zk_ctx* ctx, HFe* out, int ne) {
    int rc = wait_mailbox(ctx, ctx->mail_host, ctx->mail_seq, true);
    if (rc) return rc;
    memcpy(out, const_cast<const Fe*>(ctx->mail_host->vals), (size_t)ne * sizeof(Fe));
    return ZK_OK;
}

int launch_round_evals(zk_ctx* ctx, const TablePtrs& tp, int P, int D, uint64_t len, bool shared, int nlin, bool skip1) {
    prof_begin(ctx);
    int rc;
    ZK_DISPATCH_FID(ctx, rc = launch_round_evals_pd<FID>(ctx, tp, P, D, nlin, len / 2, shared, skip1));
    prof_end(ctx, 32.0 * (P * D + nlin) * (double)len);
    return rc;
}
// old length `len` (>= 4): folds to len/2 and evaluates the next round
int launch_fold_evals(zk_ctx* ctx, const TablePtrs& tp, int P, int D, uint64_t len, const FoldTable& ft, bool skip1, bool shared, int nlin) {
    prof_begin(ctx);
    int rc;
    ZK_DISPATCH_FID(ctx, rc = launch_fold_evals_pd<FID>(ctx, tp, P, D, nlin, len / 4, ft, skip1, shared));
    prof_end(ctx, 32.0 * (P * D + nlin) * 1.5 * (double)len);
    return rc;
}
int launch_fold0(zk_ctx* ctx, const TablePtrs& tp, int ntables, uint64_t len, const FoldTable& ft) {
    ZK_DISPATCH_FID(ctx, (fold0_kernel<FID><<<grid_for(ctx, len / 2, 4), kThreads, 0, ctx->stream>>>(tp, ntables, len / 2, ft)));
    return post_launch(ctx);
}

// A host-driven round costs a launch, a PCIe mailbox write, a host Keccak and the next launch (~16 us) around its kernel;
// in the persistent launch a round costs a grid barrier and the device transcript (~8 us), and the kernel's bandwidth
// efficiency on big tables is that of the per-round kernels (same loop body).  The hand-over pays once the rounds stop
// being pure bandwidth: tables x entries <= 2^tail_log (profiles/r02).
bool dev_rounds_apply(const zk_ctx* ctx, uint64_t len, int tables, uint32_t flags) {
    if ((flags & ZK_FLAG_HOST_ROUNDS) || ctx->tail_log <= 0 || len < 2) return false;
    uint64_t budget = (1ull << ctx->tail_log) / (uint64_t)(tables < 1 ? 1 : tables);
    return len <= budget;
}

int run_dev_rounds(zk_ctx* ctx, const TablePtrs& tp, int P, int D, int nlin, int mode, uint64_t len, const HFe* pending_r, HostTranscript& tr,
                   uint64_t* vals_out, uint64_t* chal_out, uint64_t* finals, uint32_t max_rounds, bool sharded, uint32_t* rounds_run) {
    const int NE = D + 1, T = P * D + nlin;
    if (P < 1 || D < 1 || NE > kMaxEvals || T > kMaxTables || !supported_pd((uint32_t)P, (uint32_t)D) || (nlin != 0 && !(P == 1 && D == 2 && nlin == 1)))
        return fail(ctx, ZK_ERR_ARG, kUnsupportedPdMsg);
    if (!is_pow2(len) || len < 2 || ilog2(len) > (uint32_t)kDevMaxLog) return fail(ctx, ZK_ERR_ARG, "internal: table too long for the device-resident rounds");
    if (sharded && (ctx->world < 2 || !ctx->peers_attached)) return fail(ctx, ZK_ERR_ARG, "internal: sharded device rounds without attached peers");
    DevArgs a;
    memset(&a, 0, sizeof a);
    a.tp = tp;
    a.log_len = ilog2(len);
    a.pending = pending_r ? 1u : 0u;
    a.mode = (uint32_t)mode;
    a.seq = ++ctx->dev_seq;
    const uint32_t all_rounds = a.log_len - (a.pending ? 1u : 0u);   // rounds until one entry is left
    a.max_rounds = max_rounds ? (max_rounds < all_rounds ? max_rounds : all_rounds) : all_rounds;
    if (a.max_rounds > (uint32_t)kDevMaxRounds) return fail(ctx, ZK_ERR_ARG, "internal: too many rounds for one launch");
    if (pending_r) a.ft = make_fold_table(ctx->field, *pending_r);
    static_assert(sizeof(a.interp) == sizeof(Fe) * kMaxEvals * kMaxEvals, "DevArgs::interp holds kMaxEvals^2 elements");
    memcpy(a.interp, interp_for(ctx, D).matrix(), (size_t)NE * NE * sizeof(Fe));   // NE <= kMaxEvals checked above
    for (int i = 0; i < NE * NE; ++i) {
        const HFe plain = ctx->field.from_mont(interp_for(ctx, D).matrix()[i]);
        memcpy(a.interp_plain[i].v, plain.l, 32);
    }
    memcpy(a.pow32, ctx->pow32, sizeof a.pow32);
    tr.export_state(a.sponge.s, &a.sponge.pos);
    a.out = ctx->dev_dev;
    a.g = ctx->dev_global;
    a.world = 1;
    a.final_fold = sharded ? 0u : 1u;
    if (sharded) {
        a.world = (uint32_t)ctx->world;
        a.rank = (uint32_t)ctx->rank;
        a.xseq = ctx->xseq;
        for (int q = 0; q < ctx->world; ++q) a.peers[q] = ctx->peer_slots[q];
    }
    if (ctx->dev_global_dirty) {   // a previous launch gave up: its barrier words are in an unknown state
        ZK_CUDA(cudaMemsetAsync(ctx->dev_global, 0, sizeof(DevGlobal), ctx->stream));
        ctx->dev_global_dirty = false;
    }
    // grid: the blocks that have work in the first round, capped by what the device keeps resident (the barrier spins)
    const int key = (P << 8) | (D << 4) | nlin;
    auto cap = ctx->dev_capacity.find(key);
    if (cap == ctx->dev_capacity.end()) {
        int blocks = 0, rc0;
        ZK_DISPATCH_FID(ctx, rc0 = launch_dev_rounds_pd<FID>(ctx, P, D, nlin, a, 0, &blocks));
        if (rc0) return rc0;
        cap = ctx->dev_capacity.emplace(key, blocks).first;
    }
    const uint64_t work = a.pending ? len / 4 : len / 2;
    uint64_t grid = (work + 31) / 32;   // the round's items go to warps, consecutive warps to different blocks (devrounds.cuh)
    if (grid > (uint64_t)cap->second) grid = cap->second;
    if (ctx->grid_cap > 0 && grid > (uint64_t)ctx->grid_cap) grid = ctx->grid_cap;
    if (grid < 1) grid = 1;
    // algorithmic bytes of the rounds the launch covers (same accounting as the per-round launches)
    double bytes = 0;
    {
        uint64_t l = len;
        bool pend = pending_r != nullptr;
        for (uint32_t k = 0; k < a.max_rounds; ++k) {
            bytes += pend ? 48.0 * T * (double)l : 32.0 * T * (double)l;
            if (pend) l /= 2;
            pend = true;
        }
    }
    prof_begin(ctx);
    int rc;
    ZK_DISPATCH_FID(ctx, rc = launch_dev_rounds_pd<FID>(ctx, P, D, nlin, a, (int)grid, nullptr));
    prof_end(ctx, bytes);
    if (rc) return rc;
    if (ctx->dev_hook) ctx->dev_hook(chal_out);   // the round loop is queued: the caller may queue side work behind / beside it
    rc = wait_seq(ctx, &ctx->dev_host->seq, a.seq, true);
    if (rc) { ctx->dev_global_dirty = true; return rc; }
    const DevOut* o = ctx->dev_host;
    if (o->status != kDevOk) {
        ctx->dev_global_dirty = true;
        static const char* what[] = {"", "device rounds: timed out waiting for the blocks of a round", "device rounds: timed out waiting for the leader's challenge",
                                     "device rounds: timed out waiting for a peer rank's partial evaluations"};
        return fail(ctx, ZK_ERR_CUDA, what[o->status < 4 ? o->status : 1]);
    }
    if (o->rounds != a.max_rounds) return fail(ctx, ZK_ERR_CUDA, "device rounds ran an unexpected number of rounds");
#ifdef ZK_DEV_TIMING
    if (getenv("ZKB200_DEV_TIMING")) {   // measurement builds: the leader's per-round timeline
        fprintf(stderr, "devrounds T=%d len=2^%u pending=%u grid=%d:", T, a.log_len, a.pending, (int)grid);
        for (uint32_t k = 0; k < o->rounds; ++k) {
            const unsigned long long* t = o->t_ns[k];
            const unsigned long long next = k + 1 < o->rounds ? o->t_ns[k + 1][0] : t[4];
            fprintf(stderr, " [%u: own %.1f wait %.1f evals %.1f transcript %.1f release %.1f]", k, (t[1] - t[0]) * 1e-3, (t[2] - t[1]) * 1e-3,
                    (t[3] - t[2]) * 1e-3, (t[4] - t[3]) * 1e-3, (next - t[4]) * 1e-3);
        }
        fprintf(stderr, "\n");
    }
#endif
    for (uint32_t k = 0; k < o->rounds; ++k) {
        memcpy(vals_out + (size_t)k * NE * 4, o->round_vals[k], (size_t)NE * sizeof(Fe));
        if (chal_out) memcpy(chal_out + (size_t)k * 4, &o->challenges[k], sizeof(Fe));
    }
    if (finals && a.final_fold && a.max_rounds == all_rounds) memcpy(finals, o->finals, (size_t)T * sizeof(Fe));
    tr.import_state(o->sponge.s, o->sponge.pos);
    if (sharded) ctx->xseq += o->rounds;
    if (rounds_run) *rounds_run = o->rounds;
    return ZK_OK;
}

}  // namespace zk

// =================================================================================== MLE operations
extern "C" int zk_mle_partial_evaluate(zk_ctx* ctx, zk_table* t, uint32_t var, const uint64_t r[4]) {
    if (t->len < 2) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");  // new(&[]) of the empty result
    uint32_t nvars = ilog2(t->len);
    if (var >= nvars) return fail(ctx, ZK_ERR_ASSERT, "attempt to subtract with overflow");     // evaluation_form.rs:82
    HFe rr;
    memcpy(rr.l, r, 32);
    FoldTable ft = make_fold_table(ctx->field, rr);
    if (var == 0) {
        TablePtrs tp{};
        tp.t[0] = t->d;
        int rc = launch_fold0(ctx, tp, 1, t->len, ft);
        if (rc) return rc;
        t->len /= 2;
        return ZK_OK;
    }
    uint64_t half = t->len / 2;
    int rc = ensure_scratch(ctx, (size_t)half * sizeof(Fe));
    if (rc) return rc;
    uint32_t power = nvars - 1 - var;
    ZK_DISPATCH_FID(ctx, (fold_var_kernel<FID><<<grid_for(ctx, half, 4), kThreads, 0, ctx->stream>>>(t->d, (Fe*)ctx->scratch, half, power, ft)));
    rc = post_launch(ctx);
    if (rc) return rc;
    ZK_CUDA(cudaMemcpyAsync(t->d, ctx->scratch, (size_t)half * sizeof(Fe), cudaMemcpyDeviceToDevice, ctx->stream));
    t->len = half;
    return ZK_OK;
}

extern "C" int zk_mle_to_bytes(zk_ctx* ctx, const zk_table* t, uint8_t* out_host) {
    // converted in chunks through the scratch buffer so the extra HBM stays bounded
    const uint64_t chunk = 1ull << 22;  // 128 MiB of bytes per chunk
    int rc = ensure_scratch(ctx, (size_t)(t->len < chunk ? t->len : chunk) * sizeof(Fe));
    if (rc) return rc;
    for (uint64_t off = 0; off < t->len; off += chunk) {
        uint64_t n = t->len - off < chunk ? t->len - off : chunk;
        ZK_DISPATCH_FID(ctx, (to_bytes_be_kernel<FID><<<grid_for(ctx, n, 8), kThreads, 0, ctx->stream>>>(t->d + off, (Fe*)ctx->scratch, n)));
        rc = post_launch(ctx);
        if (rc) return rc;
        ZK_CUDA(cudaMemcpyAsync(out_host + off * 32, ctx->scratch, (size_t)n * 32, cudaMemcpyDeviceToHost, ctx->stream));
        ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return ZK_OK;
}

extern "C" int zk_mle_scalar_mul(zk_ctx* ctx, const zk_table* t, const uint64_t s[4], zk_table** out) {
    int rc = table_alloc(ctx, t->len, out);
    if (rc) return rc;
    Fe sv;
    memcpy(sv.v, s, 32);
    ZK_DISPATCH_FID(ctx, (scale_kernel<FID><<<grid_for(ctx, t->len, 4), kThreads, 0, ctx->stream>>>(t->d, (*out)->d, t->len, sv)));
    return post_launch(ctx);
}
extern "C" int zk_mle_add(zk_ctx* ctx, const zk_table* a, const zk_table* b, zk_table** out) {
    if (a->len != b->len) return fail(ctx, ZK_ERR_ASSERT, "Polynomials must have same number of evaluations for addition");
    int rc = table_alloc(ctx, a->len, out);
    if (rc) return rc;
    ZK_DISPATCH_FID(ctx, (ew_kernel<FID, EW_ADD><<<grid_for(ctx, a->len, 4), kThreads, 0, ctx->stream>>>(a->d, b->d, (*out)->d, a->len)));
    return post_launch(ctx);
}
// out[b * n + c] = wb[b] (op) wc[c] into caller-provided memory (n * n elements)
int tensor_into(zk_ctx* ctx, const Fe* wb, const Fe* wc, uint64_t n, Fe* out, int op) {
    if (ilog2(n) > 31) return fail(ctx, ZK_ERR_ARG, "tensor too large");
    int grid = grid_for(ctx, n * n, 4);
    if (op == EW_ADD) { ZK_DISPATCH_FID(ctx, (tensor_kernel<FID, EW_ADD><<<grid, kThreads, 0, ctx->stream>>>(wb, wc, out, n, ilog2(n)))); }
    else { ZK_DISPATCH_FID(ctx, (tensor_kernel<FID, EW_MUL><<<grid, kThreads, 0, ctx->stream>>>(wb, wc, out, n, ilog2(n)))); }
    return post_launch(ctx);
}
static int tensor_op(zk_ctx* ctx, const zk_table* wb, const zk_table* wc, zk_table** out, int op) {
    if (wb->len != wc->len) return fail(ctx, ZK_ERR_ASSERT, "Different polynomial length");
    uint64_t n = wb->len;
    if (ilog2(n) > 31) return fail(ctx, ZK_ERR_ARG, "tensor too large");
    int rc = table_alloc(ctx, n * n, out);
    if (rc) return rc;
    return tensor_into(ctx, wb->d, wc->d, n, (*out)->d, op);
}
extern "C" int zk_mle_tensor_add(zk_ctx* ctx, const zk_table* wb, const zk_table* wc, zk_table** out) { return tensor_op(ctx, wb, wc, out, EW_ADD); }
extern "C" int zk_mle_tensor_mul(zk_ctx* ctx, const zk_table* wb, const zk_table* wc, zk_table** out) { return tensor_op(ctx, wb, wc, out, EW_MUL); }

extern "C" int zk_sum_halves(zk_ctx* ctx, const zk_table* t, uint64_t out[8]) {
    if (t->len < 2) {  // split_at(0): left empty, right = the single entry (prover.rs:79-80)
        memset(out, 0, 32);
        ZK_CUDA(cudaMemcpyAsync(out + 4, t->d, sizeof(Fe), cudaMemcpyDeviceToHost, ctx->stream));
        ZK_CUDA(cudaStreamSynchronize(ctx->stream));
        return ZK_OK;
    }
    TablePtrs tp{};
    tp.t[0] = t->d;
    int rc = launch_round_evals(ctx, tp, 1, 1, t->len);
    if (rc) return rc;
    return fetch_result(ctx, (HFe*)out, 2);
}

// =================================================================================== SumPolynomial
// the (P, D) shapes the round kernels and the device tail are instantiated for (round_launch.cuh ZK_PD_CASES); anything
// else is refused at the boundary, before any buffer sized by kMaxEvals / kMaxTables is filled
static bool supported_pd(uint32_t P, uint32_t D) {
#define ZK_CASE(PP, DD) if (P == PP && D == DD) return true;
    ZK_PD_CASES
#undef ZK_CASE
    return false;
}
static_assert(kMaxEvals >= 4, "ZK_PD_CASES goes up to D = 3");
const char* const kUnsupportedPdMsg = "unsupported (P, D): supported are (1,1) (1,2) (2,2) (3,2) (4,2) (1,3) (2,3)";

extern "C" int zk_sumpoly_create(zk_ctx* ctx, zk_table* const* tables, uint32_t P, uint32_t D, zk_sumpoly** out) {
    if (!supported_pd(P, D)) return fail(ctx, ZK_ERR_ARG, kUnsupportedPdMsg);
    for (uint32_t i = 0; i < P * D; ++i)
        if (tables[i]->len != tables[0]->len) return fail(ctx, ZK_ERR_ASSERT, "different number of variables");
    for (uint32_t i = 0; i < P * D; ++i)   // folds are in place: one table in two slots would be folded twice per round (and freed twice)
        for (uint32_t j = 0; j < i; ++j)
            if (tables[i] == tables[j] || tables[i]->d == tables[j]->d) return fail(ctx, ZK_ERR_ARG, "the tables of a sumpoly must be distinct (clone a repeated factor)");
    zk_sumpoly* sp = new zk_sumpoly();
    sp->P = P;
    sp->D = D;
    sp->len = tables[0]->len;
    sp->tabs.assign(tables, tables + P * D);
    *out = sp;
    return ZK_OK;
}
extern "C" void zk_sumpoly_free(zk_ctx* ctx, zk_sumpoly* sp) {
    if (!sp) return;
    for (zk_table* t : sp->tabs) zk_table_free(ctx, t);
    delete sp;
}
extern "C" uint64_t zk_sumpoly_len(const zk_sumpoly* sp) { return sp->tabs[0]->len; }
extern "C" zk_table* zk_sumpoly_table(const zk_sumpoly* sp, uint32_t i) { return i < sp->tabs.size() ? sp->tabs[i] : nullptr; }

TablePtrs ptrs_of(const zk_sumpoly* sp) {
    TablePtrs tp{};
    for (size_t i = 0; i < sp->tabs.size(); ++i) tp.t[i] = sp->tabs[i]->d;
    return tp;
}
void set_len(zk_sumpoly* sp, uint64_t len) {
    sp->len = len;
    for (zk_table* t : sp->tabs) t->len = len;
}
// the tables may have been refilled (zk_table_regenerate) since the last operation
int sync_len(zk_ctx* ctx, zk_sumpoly* sp) {
    for (zk_table* t : sp->tabs)
        if (t->len != sp->tabs[0]->len) return fail(ctx, ZK_ERR_ASSERT, "different number of variables");
    sp->len = sp->tabs[0]->len;
    return ZK_OK;
}

extern "C" int zk_sumpoly_reduce(zk_ctx* ctx, const zk_sumpoly* sp_, zk_table** out) {
    zk_sumpoly* sp = const_cast<zk_sumpoly*>(sp_);
    if (int rc0 = sync_len(ctx, sp)) return rc0;
    if (sp->P < 2) return fail(ctx, ZK_ERR_ASSERT, "more than one product polynomial required for add operation");
    if (sp->D < 2) return fail(ctx, ZK_ERR_ASSERT, "more than one polynomial required for mul operation");
    int rc = table_alloc(ctx, sp->len, out);
    if (rc) return rc;
    ZK_DISPATCH_FID(ctx, (sumpoly_reduce_kernel<FID><<<grid_for(ctx, sp->len, 4), kThreads, 0, ctx->stream>>>(ptrs_of(sp), sp->P, sp->D, (*out)->d, sp->len)));
    return post_launch(ctx);
}

extern "C" int zk_sumcheck_round_evals(zk_ctx* ctx, zk_sumpoly* sp, uint64_t* evals) {
    if (int rc0 = sync_len(ctx, sp)) return rc0;
    if (sp->len < 2) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");
    int rc = launch_round_evals(ctx, ptrs_of(sp), sp->P, sp->D, sp->len);
    if (rc) return rc;
    return fetch_result(ctx, (HFe*)evals, sp->D + 1);
}

extern "C" int zk_sumcheck_fold_and_evals(zk_ctx* ctx, zk_sumpoly* sp, const uint64_t r[4], uint64_t* evals) {
    if (int rc0 = sync_len(ctx, sp)) return rc0;
    if (sp->len < 2) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");
    HFe rr;
    memcpy(rr.l, r, 32);
    FoldTable ft = make_fold_table(ctx->field, rr);
    TablePtrs tp = ptrs_of(sp);
    if (evals && sp->len >= 4) {
        int rc = launch_fold_evals(ctx, tp, sp->P, sp->D, sp->len, ft, false);
        if (rc) return rc;
        set_len(sp, sp->len / 2);
        return fetch_result(ctx, (HFe*)evals, sp->D + 1);
    }
    if (evals) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");  // a 1-entry table has no next round
    int rc = launch_fold0(ctx, tp, (int)sp->tabs.size(), sp->len, ft);
    if (rc) return rc;
    set_len(sp, sp->len / 2);
    return ZK_OK;
}

// =================================================================================== one-shot provers
const Interpolator& interp_for(zk_ctx* ctx, int degree) {
    auto it = ctx->interps.find(degree);
    if (it == ctx->interps.end()) it = ctx->interps.emplace(degree, Interpolator(ctx->field, degree)).first;
    return it->second;
}

// sumcheck_gkr_protocol::prove -- sumcheck_gkr_protocol.rs:24-67
extern "C" int zk_prove_product(zk_ctx* ctx, zk_sumpoly* sp, const uint64_t claimed_sum[4], zk_transcript* tr,
                                uint64_t* coeffs_out, uint64_t* challenges_out, uint64_t* final_values, uint32_t flags) {
    if (int rc0 = sync_len(ctx, sp)) return rc0;
    if (!is_pow2(sp->len)) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");
    const HostField& f = ctx->field;
    const int D = sp->D, P = sp->P, NE = D + 1, NL = (int)sp->nlin, T = P * D + NL;
    const uint32_t n = ilog2(sp->len);
    const Interpolator& ip = interp_for(ctx, D);
    HFe claim;
    memcpy(claim.l, claimed_sum, 32);
    if (!(flags & ZK_FLAG_NO_CLAIM_ABSORB)) tr->t.append_be(f, claim);           // :35
    TablePtrs tp = ptrs_of(sp);
    HFe evals[kMaxEvals], coeffs[kMaxEvals], r = f.zero();
    HFe running = claim;  // s_{k-1}(r_{k-1}); only trusted from round 1 on
    // the caller vouches for claimed_sum (the GKR layer prover computed it): round 0 can derive s(1) like every later round
    const bool trusted0 = (flags & ZK_FLAG_TRUSTED_CLAIM) && !(flags & ZK_FLAG_DIRECT_S1) && round_evals_skip1_supported(P, D, NL);
    for (uint32_t k = 0; k < n; ++k) {                                           // :37
        const bool skip1 = (k > 0 || trusted0) && !(flags & ZK_FLAG_DIRECT_S1);
        int rc;
        if (dev_rounds_apply(ctx, sp->len, T, flags)) {   // rounds k..n-1 and the last fold in one launch, transcript on the device
            rc = run_dev_rounds(ctx, tp, P, D, NL, kDevProduct, sp->len, k > 0 ? &r : nullptr, tr->t,
                                coeffs_out + (size_t)k * NE * 4, challenges_out + (size_t)k * 4, final_values);
            if (rc) return rc;
            set_len(sp, 1);
            return ZK_OK;
        }
        if (k == 0) {
            rc = launch_round_evals(ctx, tp, P, D, sp->len, false, NL, skip1);   // :41 generate_round_univariate
        } else {
            rc = launch_fold_evals(ctx, tp, P, D, sp->len, make_fold_table(f, r), skip1, false, NL);   // :57 fused with :41
            set_len(sp, sp->len / 2);
        }
        if (rc) return rc;
        rc = fetch_result(ctx, evals, NE);
        if (rc) return rc;
        if (skip1) evals[1] = f.sub(running, evals[0]);                          // s(0) + s(1) == previous s(r)
        ip.coefficients(evals, coeffs);                                          // :46-50 lagrange_interpolate
        uint8_t bytes[32 * kMaxEvals];
        for (int i = 0; i < NE; ++i) f.to_bytes_le(coeffs[i], bytes + 32 * i);  // :145-150 little-endian
        tr->t.append(bytes, 32 * NE);                                            // :52
        r = tr->t.challenge(f);                                                  // :55
        running = f.horner(coeffs, NE, r);
        memcpy(coeffs_out + (size_t)k * NE * 4, coeffs, 32 * NE);
        memcpy(challenges_out + (size_t)k * 4, r.l, 32);                         // :59
    }
    if (n > 0) {                                                                 // :57 last partial_evaluate
        int rc = launch_fold0(ctx, tp, T, sp->len, make_fold_table(f, r));
        if (rc) return rc;
        set_len(sp, sp->len / 2);
    }
    if (final_values) {
        for (int t = 0; t < T; ++t)
            ZK_CUDA(cudaMemcpyAsync(final_values + 4 * t, sp->tabs[t]->d, sizeof(Fe), cudaMemcpyDeviceToHost, ctx->stream));
    }
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

// Prover::init + Prover::prove -- prover.rs:22-33,35-71
extern "C" int zk_prove_basic_device(zk_ctx* ctx, zk_table* t, uint64_t claimed_sum[4], uint64_t* round_polys,
                                     uint64_t* challenges, uint64_t final_value[4], uint32_t flags) {
    if (!is_pow2(t->len)) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");
    const HostField& f = ctx->field;
    const uint32_t n = ilog2(t->len);
    HostTranscript tr;
    TablePtrs tp{};
    tp.t[0] = t->d;
    HFe evals[2], r = f.zero(), claimed;
    int rc;
    // round 0's two half sums also give the claimed sum of init (prover.rs:28): sum = left + right
    if (n == 0) {
        ZK_CUDA(cudaMemcpyAsync(claimed.l, t->d, sizeof(Fe), cudaMemcpyDeviceToHost, ctx->stream));
        ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
        rc = launch_round_evals(ctx, tp, 1, 1, t->len);
        if (rc) return rc;
        rc = fetch_result(ctx, evals, 2);
        if (rc) return rc;
        claimed = f.add(evals[0], evals[1]);
    }
    if (!(flags & ZK_FLAG_SKIP_ABSORB)) {                                        // :38-39 convert_to_bytes + append
        // stream the big-endian bytes through pinned chunks into the host sponge
        const uint64_t chunk = 1ull << 19;  // 16 MiB
        rc = ensure_scratch(ctx, (size_t)(t->len < chunk ? t->len : chunk) * sizeof(Fe));
        if (rc) return rc;
        if (!ctx->pinned) ZK_CUDA(cudaHostAlloc(&ctx->pinned, (size_t)chunk * 32, cudaHostAllocDefault));
        for (uint64_t off = 0; off < t->len; off += chunk) {
            uint64_t m = t->len - off < chunk ? t->len - off : chunk;
            ZK_DISPATCH_FID(ctx, (to_bytes_be_kernel<FID><<<grid_for(ctx, m, 8), kThreads, 0, ctx->stream>>>(t->d + off, (Fe*)ctx->scratch, m)));
            rc = post_launch(ctx);
            if (rc) return rc;
            ZK_CUDA(cudaMemcpyAsync(ctx->pinned, ctx->scratch, (size_t)m * 32, cudaMemcpyDeviceToHost, ctx->stream));
            ZK_CUDA(cudaStreamSynchronize(ctx->stream));
            tr.append((const uint8_t*)ctx->pinned, (size_t)m * 32);
        }
    }
    tr.append_be(f, claimed);                                                    // :40-41
    memcpy(claimed_sum, claimed.l, 32);
    for (uint32_t k = 0; k < n; ++k) {                                           // :46
        if (k > 0 && dev_rounds_apply(ctx, t->len, 1, flags)) {   // rounds k..n-1 and the last fold in one launch
            uint64_t* chal = challenges ? challenges + (size_t)k * 4 : nullptr;
            rc = run_dev_rounds(ctx, tp, 1, 1, 0, kDevPlain, t->len, &r, tr, round_polys + (size_t)k * 8, chal, final_value);
            if (rc) return rc;
            t->len = 1;
            return ZK_OK;
        }
        if (k > 0) {
            rc = launch_fold_evals(ctx, tp, 1, 1, t->len, make_fold_table(f, r), false);  // :61-63 fused with :50
            t->len /= 2;
            if (rc) return rc;
            rc = fetch_result(ctx, evals, 2);
            if (rc) return rc;
        }
        uint8_t bytes[64];
        f.to_bytes_be(evals[0], bytes);
        f.to_bytes_be(evals[1], bytes + 32);
        tr.append(bytes, 64);                                                    // :51-55
        r = tr.challenge(f);                                                     // :58
        memcpy(round_polys + (size_t)k * 8, evals, 64);
        if (challenges) memcpy(challenges + (size_t)k * 4, r.l, 32);
    }
    if (n > 0) {
        rc = launch_fold0(ctx, tp, 1, t->len, make_fold_table(f, r));            // :61-63 of the last round
        if (rc) return rc;
        t->len /= 2;
    }
    if (final_value) ZK_CUDA(cudaMemcpyAsync(final_value, t->d, sizeof(Fe), cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

extern "C" int zk_prove_basic(zk_ctx* ctx, const uint64_t* host_table, uint64_t n, uint64_t claimed_sum[4], uint64_t* round_polys,
                              uint64_t* challenges, uint64_t final_value[4], uint32_t flags) {
    zk_table* t = nullptr;
    int rc = zk_table_upload(ctx, host_table, n, &t);
    if (rc) return rc;
    rc = zk_prove_basic_device(ctx, t, claimed_sum, round_polys, challenges, final_value, flags);
    zk_table_free(ctx, t);
    return rc;
}

extern "C" int zk_prove_product_host(zk_ctx* ctx, const uint64_t* host_tables, uint32_t P, uint32_t D, uint64_t n,
                                     const uint64_t claimed_sum[4], zk_transcript* tr, uint64_t* coeffs, uint64_t* challenges,
                                     uint64_t* final_values, uint32_t flags) {
    if (!supported_pd(P, D)) return fail(ctx, ZK_ERR_ARG, kUnsupportedPdMsg);
    if (!is_pow2(n)) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");
    std::vector<zk_table*> tabs(P * D, nullptr);
    int rc = ZK_OK;
    for (uint32_t i = 0; i < P * D && rc == ZK_OK; ++i) {
        rc = table_alloc(ctx, n, &tabs[i]);
        if (rc == ZK_OK) {
            cudaError_t e = cudaMemcpyAsync(tabs[i]->d, host_tables + (size_t)i * n * 4, (size_t)n * sizeof(Fe), cudaMemcpyHostToDevice, ctx->stream);
            if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = ZK_ERR_CUDA; }
        }
    }
    zk_sumpoly* sp = nullptr;
    if (rc == ZK_OK) rc = zk_sumpoly_create(ctx, tabs.data(), P, D, &sp);
    if (rc != ZK_OK) {
        for (zk_table* t : tabs) zk_table_free(ctx, t);
        return rc;
    }
    rc = zk_prove_product(ctx, sp, claimed_sum, tr, coeffs, challenges, final_values, flags);
    zk_sumpoly_free(ctx, sp);
    return rc;
}

// =================================================================================== arithmetic probe
// Times `iters` x 4 field operations per thread on a full grid (blocks_per_sm x SMs x 256 threads);
// returns operations per second.  kind: 0 mont_mul, 1 fold-by-scalar, 2 unreduced multiply-accumulate.
// (The raw pipe-rate probes of round 1 -- DFMA, IMAD.WIDE forms, column accumulators -- live in experiments/.)
extern "C" int zk_arith_probe(zk_ctx* ctx, int kind, uint32_t iters, int blocks_per_sm, double* ops_per_s, double* ms_out) {
    if (kind < 0 || kind > 2 || blocks_per_sm < 1 || blocks_per_sm > 8) return fail(ctx, ZK_ERR_ARG, "bad probe arguments");
    int grid = ctx->sm_count * blocks_per_sm;
    int rc = ensure_scratch(ctx, (size_t)grid * kThreads * sizeof(Fe));
    if (rc) return rc;
    FoldTable ft = make_fold_table(ctx->field, ctx->field.from_u64(0x123456789abcdefull));
    cudaEvent_t e0, e1;
    ZK_CUDA(cudaEventCreate(&e0));
    ZK_CUDA(cudaEventCreate(&e1));
    for (int pass = 0; pass < 2; ++pass) {  // first pass warms up
        ZK_CUDA(cudaEventRecord(e0, ctx->stream));
        if (kind == 0) { ZK_DISPATCH_FID(ctx, (arith_probe_kernel<FID, 0><<<grid, kThreads, 0, ctx->stream>>>((Fe*)ctx->scratch, iters, ft))); }
        else if (kind == 1) { ZK_DISPATCH_FID(ctx, (arith_probe_kernel<FID, 1><<<grid, kThreads, 0, ctx->stream>>>((Fe*)ctx->scratch, iters, ft))); }
        else { ZK_DISPATCH_FID(ctx, (arith_probe_kernel<FID, 2><<<grid, kThreads, 0, ctx->stream>>>((Fe*)ctx->scratch, iters, ft))); }
        rc = post_launch(ctx);
        if (rc) return rc;
        ZK_CUDA(cudaEventRecord(e1, ctx->stream));
        ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    float ms = 0;
    ZK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_out) *ms_out = ms;
    if (ops_per_s) *ops_per_s = (double)grid * kThreads * 4.0 * iters / (ms * 1e-3);
    return ZK_OK;
}
