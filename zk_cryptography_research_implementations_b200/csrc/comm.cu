// comm.cu -- one process per GPU: tables sharded across ranks, one tiny exchange per round.
//
// Partition: with G = 2^g ranks, rank q holds the entries whose LOW g index bits are q
// (local j <-> global j*G + q).  The prover binds variables MSB first, so every fold pairs
// (j, j + M/2) inside one rank: folds never communicate.  Per round each rank's kernel produces
// partial evaluations of the round polynomial; an NCCL all-gather over NVLink moves the G x (d+1)
// elements, every rank adds them in the field and runs the same host transcript, so the challenge
// needs no broadcast.  When the local tables shrink to `collapse_len` entries they are all-gathered
// and interleaved, and the remaining rounds run redundantly on every rank with no communication.
//
// NCCL is resolved with dlopen (the copy torch already loaded), so the library still loads on a
// CPU-only box; nothing here runs without a GPU.
#include <dlfcn.h>
#include <fcntl.h>
#include <nccl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <string>
#include <vector>

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

namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool load() {
        if (handle) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { error = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
#define ZK_SYM(field, name)                                                         \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, name));                 \
    if (!field) { error = std::string("dlsym ") + name; handle = nullptr; return false; }
        ZK_SYM(GetUniqueId, "ncclGetUniqueId")
        ZK_SYM(CommInitRank, "ncclCommInitRank")
        ZK_SYM(CommDestroy, "ncclCommDestroy")
        ZK_SYM(AllGather, "ncclAllGather")
        ZK_SYM(GroupStart, "ncclGroupStart")
        ZK_SYM(GroupEnd, "ncclGroupEnd")
        ZK_SYM(GetErrorString, "ncclGetErrorString")
#undef ZK_SYM
        return true;
    }
};
NcclApi g_nccl;

#define ZK_NCCL(call)                                                                   \
    do {                                                                                \
        ncclResult_t r__ = (call);                                                      \
        if (r__ != ncclSuccess) {                                                       \
            ctx->err = std::string(#call) + ": " + g_nccl.GetErrorString(r__);          \
            return ZK_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

inline bool is_pow2(uint64_t n) { return n && !(n & (n - 1)); }
inline uint32_t ilog2(uint64_t n) { uint32_t k = 0; while (n >>= 1) ++k; return k; }

// gathered[q][j] (rank-major, m entries per rank) -> out[j * G + q]: the sharded variables become
// the low index bits of one table again
__global__ void __launch_bounds__(kThreads) interleave_kernel(const Fe* gathered, Fe* out, uint64_t m, uint32_t G) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, total = m * G;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        uint64_t j = i / G, q = i % G;
        st256(out + i, ld256(gathered + q * m + j));
    }
}
}  // namespace

// In-kernel exchange of the device-resident rounds (devrounds.cuh): every rank owns a PeerSlot[2][world] array in its
// HBM; the other ranks map it (cudaIpc, peer access over NVLink) and their round-loop kernels write their partial
// evaluations straight into it.  Collective over the communicator; leaves peers_attached false (and the host-mailbox
// exchange in charge) if any rank cannot map any peer.  ZKB200_PEER_EXCHANGE=0 skips it.
static int attach_peers(zk_ctx* ctx) {
    const int G = ctx->world;
    ctx->peers_attached = false;
    if (G < 2 || G > kMaxRanks) return ZK_OK;
    if (const char* knob = getenv("ZKB200_PEER_EXCHANGE"))
        if (knob[0] == '0') return ZK_OK;
    PeerSlot* own = nullptr;
    ZK_CUDA(cudaMalloc(&own, sizeof(PeerSlot) * 2 * (size_t)G));
    ZK_CUDA(cudaMemset(own, 0, sizeof(PeerSlot) * 2 * (size_t)G));
    ctx->peer_slots[ctx->rank] = own;
    cudaIpcMemHandle_t mine;
    int ok = cudaIpcGetMemHandle(&mine, own) == cudaSuccess ? 1 : 0;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    // all-gather {handle, ok} through NCCL (device staging buffers)
    const size_t rec = 64 + 8;
    uint8_t *d_send = nullptr, *d_recv = nullptr;
    ZK_CUDA(cudaMalloc(&d_send, rec));
    ZK_CUDA(cudaMalloc(&d_recv, rec * G));
    std::vector<uint8_t> h_send(rec, 0), h_recv(rec * G, 0);
    auto gather = [&]() -> int {
        ZK_CUDA(cudaMemcpyAsync(d_send, h_send.data(), rec, cudaMemcpyHostToDevice, ctx->stream));
        ZK_NCCL(g_nccl.AllGather(d_send, d_recv, rec, ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
        ZK_CUDA(cudaMemcpyAsync(h_recv.data(), d_recv, rec * G, cudaMemcpyDeviceToHost, ctx->stream));
        ZK_CUDA(cudaStreamSynchronize(ctx->stream));
        return ZK_OK;
    };
    memcpy(h_send.data(), &mine, 64);
    h_send[64] = (uint8_t)ok;
    int rc = gather();
    if (rc) return rc;
    for (int q = 0; q < G; ++q) ok = ok && h_recv[rec * q + 64];
    if (ok) {
        for (int q = 0; q < G && ok; ++q) {
            if (q == ctx->rank) continue;
            cudaIpcMemHandle_t h;
            memcpy(&h, h_recv.data() + rec * q, 64);
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();   // clear: falling back is not an error
                ok = 0;
            } else {
                ctx->peer_slots[q] = (PeerSlot*)p;
            }
        }
    }
    // every rank must have mapped every peer, or nobody uses the path
    h_send[64] = (uint8_t)ok;
    rc = gather();
    cudaFree(d_send);
    cudaFree(d_recv);
    if (rc) return rc;
    for (int q = 0; q < G; ++q) ok = ok && h_recv[rec * q + 64];
    ctx->peers_attached = ok != 0;
    return ZK_OK;
}
static void detach_peers(zk_ctx* ctx) {
    for (int q = 0; q < kMaxRanks; ++q) {
        if (!ctx->peer_slots[q]) continue;
        if (q == ctx->rank) cudaFree(ctx->peer_slots[q]);
        else cudaIpcCloseMemHandle(ctx->peer_slots[q]);
        ctx->peer_slots[q] = nullptr;
    }
    ctx->peers_attached = false;
}
// All ranks agree on the exchange sequence number a sharded prove starts from (a rank that failed half-way through an
// earlier prove would otherwise wait for numbers its peers never send): one 4-byte all-gather per prove.
static int agree_xseq(zk_ctx* ctx) {
    const int G = ctx->world;
    ZK_CUDA(cudaMemcpyAsync(ctx->xchg_send, &ctx->xseq, sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
    ZK_NCCL(g_nccl.AllGather(ctx->xchg_send, ctx->xchg_recv, sizeof(unsigned), ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    ZK_CUDA(cudaMemcpyAsync(ctx->xchg_host, ctx->xchg_recv, sizeof(unsigned) * G, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    unsigned mx = 0;
    for (int q = 0; q < G; ++q) mx = std::max(mx, reinterpret_cast<const unsigned*>(ctx->xchg_host)[q]);
    ctx->xseq = (mx + 128u) & ~63u;
    return ZK_OK;
}
// rounds the sharded prover runs before the collapse, counted from a round that sees local tables of `len` entries:
// round 0 evaluates without folding, every later round folds first and stays sharded while len/2 > collapse_len
static uint32_t sharded_rounds_from(uint64_t len, bool pending, uint64_t collapse_len) {
    uint32_t r = 0;
    if (!pending) { r = 1; }          // this round only evaluates; the next one is the first to fold `len`
    while (len / 2 > collapse_len && len >= 4) { ++r; len /= 2; }
    return r;
}

extern "C" int zk_comm_unique_id(uint8_t out[128]) {
    if (!g_nccl.load()) return ZK_ERR_CUDA;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return ZK_ERR_CUDA;
    static_assert(sizeof(id) == 128, "ncclUniqueId size");
    memcpy(out, &id, 128);
    return ZK_OK;
}

extern "C" int zk_comm_init(zk_ctx* ctx, int rank, int world, const uint8_t id_bytes[128]) {
    if (world < 1 || !is_pow2((uint64_t)world) || rank < 0 || rank >= world) return fail(ctx, ZK_ERR_ARG, "world must be a power of two, 0 <= rank < world");
    if (ctx->nccl_comm) return fail(ctx, ZK_ERR_ARG, "communicator already initialised");
    if (!g_nccl.load()) { ctx->err = g_nccl.error; return ZK_ERR_CUDA; }
    ZK_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, 128);
    ncclComm_t comm;
    ZK_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    ZK_CUDA(cudaMalloc(&ctx->xchg_send, kMaxEvals * sizeof(Fe)));
    ZK_CUDA(cudaMalloc(&ctx->xchg_recv, (size_t)world * kMaxEvals * sizeof(Fe)));
    ZK_CUDA(cudaHostAlloc(&ctx->xchg_host, (size_t)world * kMaxEvals * sizeof(Fe), cudaHostAllocDefault));
    return attach_peers(ctx);
}
// 1 if the per-round exchange of the sharded provers runs inside the round-loop kernels over peer memory (NVLink)
extern "C" int zk_comm_peer_exchange(const zk_ctx* ctx) { return ctx->peers_attached ? 1 : 0; }

extern "C" int zk_comm_destroy(zk_ctx* ctx) {
    if (!ctx->nccl_comm) return ZK_OK;
    cudaStreamSynchronize(ctx->stream);
    detach_peers(ctx);
    g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    if (ctx->xmail_host) {
        cudaHostUnregister(ctx->xmail_host);
        munmap(ctx->xmail_host, ctx->xmail_bytes);
        ctx->xmail_host = ctx->xmail_dev = nullptr;
    }
    cudaFree(ctx->xchg_send);
    cudaFree(ctx->xchg_recv);
    cudaFreeHost(ctx->xchg_host);
    ctx->xchg_send = ctx->xchg_recv = ctx->xchg_host = nullptr;
    ctx->world = 1;
    ctx->rank = 0;
    return ZK_OK;
}

extern "C" int zk_comm_rank(const zk_ctx* ctx) { return ctx->rank; }
extern "C" int zk_comm_world(const zk_ctx* ctx) { return ctx->world; }

// Attach the shared round mailboxes: a POSIX shared-memory segment holding Mailbox[2][world], mapped by every
// rank process and registered with CUDA so each rank's round kernel can publish straight into it.  Rank 0
// passes create = 1 (and may shm_unlink the name once every rank has attached).
extern "C" int zk_comm_attach_mailboxes(zk_ctx* ctx, const char* shm_name, int create) {
    if (ctx->world < 2) return fail(ctx, ZK_ERR_ARG, "zk_comm_init first");
    if (ctx->xmail_host) return fail(ctx, ZK_ERR_ARG, "mailboxes already attached");
    size_t bytes = sizeof(Mailbox) * 2 * (size_t)ctx->world;
    bytes = (bytes + 4095) & ~(size_t)4095;
    int fd = shm_open(shm_name, create ? (O_CREAT | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return fail(ctx, ZK_ERR_ARG, "shm_open failed");
    if (create && ftruncate(fd, (off_t)bytes) != 0) { close(fd); return fail(ctx, ZK_ERR_ARG, "ftruncate failed"); }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return fail(ctx, ZK_ERR_ARG, "mmap failed");
    if (create) memset(p, 0, bytes);
    ZK_CUDA(cudaSetDevice(ctx->device));
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable);
    if (e != cudaSuccess) { munmap(p, bytes); ctx->err = std::string("cudaHostRegister: ") + cudaGetErrorString(e); return ZK_ERR_CUDA; }
    ctx->xmail_host = (Mailbox*)p;
    ctx->xmail_bytes = bytes;
    ZK_CUDA(cudaHostGetDevicePointer((void**)&ctx->xmail_dev, p, 0));
    ctx->xmail_seq = 0;
    return ZK_OK;
}
extern "C" int zk_comm_unlink_mailboxes(const char* shm_name) { return shm_unlink(shm_name) == 0 ? ZK_OK : ZK_ERR_ARG; }

// All ranks: sum over ranks of the `ne` elements the last round kernel published.  Field addition is not an
// NCCL reduction op, so the raw elements are exchanged and added mod p on the host by every rank.
//  * shared mailboxes attached (default): every rank's kernel wrote into its slot of the shared segment; spin
//    until all G slots carry this round's sequence number -- no collective, no copy, no stream sync;
//  * otherwise: ncclAllGather of the (d+1) x 32 bytes over NVLink + D2H.
static int exchange_sum(zk_ctx* ctx, HFe* vals, int ne) {
    const HostField& f = ctx->field;
    const int G = ctx->world;
    const size_t bytes = (size_t)ne * sizeof(Fe);
    for (int e = 0; e < ne; ++e) vals[e] = f.zero();
    if (ctx->exchange_pending) {
        const unsigned seq = ctx->xmail_seq;
        const Mailbox* row = ctx->xmail_host + (size_t)(seq & 1u) * G;
        for (int q = 0; q < G; ++q) {
            int rc = wait_mailbox(ctx, row + q, seq, q == ctx->rank);
            if (rc) return rc;
            const HFe* v = reinterpret_cast<const HFe*>(const_cast<const Fe*>(row[q].vals));
            for (int e = 0; e < ne; ++e) vals[e] = f.add(vals[e], v[e]);
        }
        return ZK_OK;
    }
    // the kernel published into this context's own mailbox; stage a device copy for NCCL on the same stream
    ZK_CUDA(cudaMemcpyAsync(ctx->xchg_send, ctx->mail_dev->vals, bytes, cudaMemcpyDefault, ctx->stream));
    ZK_NCCL(g_nccl.AllGather(ctx->xchg_send, ctx->xchg_recv, bytes, ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    ZK_CUDA(cudaMemcpyAsync(ctx->xchg_host, ctx->xchg_recv, bytes * G, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    const HFe* all = reinterpret_cast<const HFe*>(ctx->xchg_host);
    for (int e = 0; e < ne; ++e)
        for (int q = 0; q < G; ++q) vals[e] = f.add(vals[e], all[(size_t)q * ne + e]);
    return ZK_OK;
}

// recv[q * bytes ..] = rank q's `bytes` bytes of `send` (host buffers; staged through the device for NCCL).
// `recv` must hold ctx->world * bytes.
namespace zk {
int allgather_host_bytes(zk_ctx* ctx, const void* send, size_t bytes, void* recv) {
    const int G = ctx->world;
    if (G == 1) {
        memcpy(recv, send, bytes);
        return ZK_OK;
    }
    if (!ctx->nccl_comm) return fail(ctx, ZK_ERR_ARG, "zk_comm_init has not been called");
    int rc = ensure_scratch(ctx, bytes * (size_t)(G + 1));
    if (rc) return rc;
    uint8_t* d_send = (uint8_t*)ctx->scratch;
    uint8_t* d_recv = d_send + bytes;
    ZK_CUDA(cudaMemcpyAsync(d_send, send, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ZK_NCCL(g_nccl.AllGather(d_send, d_recv, bytes, ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    ZK_CUDA(cudaMemcpyAsync(recv, d_recv, bytes * G, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}
}  // namespace zk

// all-gather every table of the sumpoly and interleave: afterwards each rank holds the full tables
static int collapse(zk_ctx* ctx, zk_sumpoly* sp) {
    const int G = ctx->world;
    const uint64_t m = sp->len;
    int rc = ensure_scratch(ctx, (size_t)G * m * sizeof(Fe));
    if (rc) return rc;
    for (zk_table* t : sp->tabs) {
        ZK_NCCL(g_nccl.AllGather(t->d, ctx->scratch, (size_t)m * sizeof(Fe), ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
        if (t->cap < m * G) {
            if (!t->owned) return fail(ctx, ZK_ERR_ARG, "collapse: wrapped table too small to hold the gathered table");
            ZK_CUDA(cudaStreamSynchronize(ctx->stream));
            ZK_CUDA(cudaFree(t->d));
            ZK_CUDA(cudaMalloc(&t->d, (size_t)m * G * sizeof(Fe)));
            t->cap = m * G;
        }
        uint64_t blocks = (m * G + kThreads - 1) / kThreads;
        if (blocks > (uint64_t)ctx->sm_count * 4) blocks = (uint64_t)ctx->sm_count * 4;
        interleave_kernel<<<(int)blocks, kThreads, 0, ctx->stream>>>((const Fe*)ctx->scratch, t->d, m, (uint32_t)G);
        ctx->launches++;
        ZK_CUDA(cudaGetLastError());
    }
    set_len(sp, m * G);
    return ZK_OK;
}

// sumcheck_gkr_protocol::prove over tables sharded across the communicator's ranks.
// `sp` holds this rank's shard (global length = world * local length).  Outputs as zk_prove_product;
// every rank returns the same proof.
extern "C" int zk_prove_product_sharded(zk_ctx* ctx, zk_sumpoly* sp, const uint64_t claimed_sum[4], zk_transcript* tr,
                                        uint64_t* coeffs_out, uint64_t* challenges_out, uint64_t* final_values,
                                        uint32_t flags, uint64_t collapse_len) {
    if (int rc0 = sync_len(ctx, sp)) return rc0;
    if (ctx->world == 1) return zk_prove_product(ctx, sp, claimed_sum, tr, coeffs_out, challenges_out, final_values, flags);
    if (!ctx->nccl_comm) return fail(ctx, ZK_ERR_ARG, "zk_comm_init has not been called");
    if (!is_pow2(sp->len)) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");
    if (collapse_len < 1) collapse_len = 1;
    const HostField& f = ctx->field;
    const int D = sp->D, P = sp->P, NE = D + 1, G = ctx->world, NL = (int)sp->nlin, T = P * D + NL;
    const uint32_t n = ilog2(sp->len) + ilog2((uint64_t)G);
    const Interpolator& ip = interp_for(ctx, D);
    HFe claim;
    memcpy(claim.l, claimed_sum, 32);
    if (!(flags & ZK_FLAG_NO_CLAIM_ABSORB)) tr->t.append_be(f, claim);
    HFe evals[kMaxEvals], coeffs[kMaxEvals], r = f.zero(), running = claim;
    bool sharded = true;
    const bool dev_exchange = ctx->peers_attached && !(flags & (ZK_FLAG_NCCL_EXCHANGE | ZK_FLAG_HOST_ROUNDS));
    if (dev_exchange) {
        int rc0 = agree_xseq(ctx);
        if (rc0) return rc0;
    }
    const bool trusted0 = (flags & ZK_FLAG_TRUSTED_CLAIM) && !(flags & ZK_FLAG_DIRECT_S1) && round_evals_skip1_supported(P, D, NL);
    for (uint32_t k = 0; k < n; ++k) {
        const bool skip1 = (k > 0 || trusted0) && !(flags & ZK_FLAG_DIRECT_S1);
        int rc;
        bool need_plain_evals = (k == 0);
        // the remaining SHARDED rounds in one persistent launch per rank: partial evaluations go from kernel to kernel
        // over peer memory, every rank runs the transcript on its GPU (devrounds.cuh); the collapse follows on the host
        if (sharded && dev_exchange && dev_rounds_apply(ctx, sp->len, T, flags) && (k == 0 ? sp->len > collapse_len : sp->len / 2 > collapse_len)) {
            const uint32_t want = sharded_rounds_from(sp->len, k > 0, collapse_len);
            uint32_t ran = 0;
            rc = run_dev_rounds(ctx, ptrs_of(sp), P, D, NL, kDevProduct, sp->len, k > 0 ? &r : nullptr, tr->t,
                                coeffs_out + (size_t)k * NE * 4, challenges_out + (size_t)k * 4, nullptr, want, true, &ran);
            if (rc) return rc;
            // the tables were folded by every challenge but the last one
            const uint32_t folds = ran - (k == 0 ? 1u : 0u);
            set_len(sp, sp->len >> folds);
            k += ran - 1;
            memcpy(r.l, challenges_out + (size_t)k * 4, 32);
            running = f.horner(reinterpret_cast<const HFe*>(coeffs_out + (size_t)k * NE * 4), NE, r);
            continue;
        }
        if (k > 0 && sharded && sp->len / 2 <= collapse_len) {
            // fold by r_{k-1} locally, then gather: the remaining rounds run on the full table
            rc = launch_fold0(ctx, ptrs_of(sp), T, sp->len, make_fold_table(f, r));
            if (rc) return rc;
            set_len(sp, sp->len / 2);
            rc = collapse(ctx, sp);
            if (rc) return rc;
            sharded = false;
            need_plain_evals = true;
        }
        if (k == 0 && sp->len <= collapse_len) {   // tiny input: gather straight away
            rc = collapse(ctx, sp);
            if (rc) return rc;
            sharded = false;
        }
        TablePtrs tp = ptrs_of(sp);
        if (!sharded && dev_rounds_apply(ctx, sp->len, T, flags)) {   // collapsed and small: the rest in one launch per rank
            rc = run_dev_rounds(ctx, tp, P, D, NL, kDevProduct, sp->len, need_plain_evals ? nullptr : &r, tr->t,
                                coeffs_out + (size_t)k * NE * 4, challenges_out + (size_t)k * 4, final_values);
            if (rc) return rc;
            set_len(sp, 1);
            return ZK_OK;
        }
        const bool shared = sharded && ctx->xmail_host != nullptr && !(flags & ZK_FLAG_NCCL_EXCHANGE);
        // a round that only evaluates (round 0, or the first one after the collapse): s(1) can be left out whenever the
        // running claim is trustworthy -- always after round 0, in round 0 only on the caller's word -- and a kernel exists
        const bool skip_now = skip1 && round_evals_skip1_supported(P, D, NL);
        if (need_plain_evals) {
            rc = launch_round_evals(ctx, tp, P, D, sp->len, shared, NL, skip_now);
        } else {
            rc = launch_fold_evals(ctx, tp, P, D, sp->len, make_fold_table(f, r), skip1, shared, NL);
            set_len(sp, sp->len / 2);
        }
        if (rc) return rc;
        if (sharded) rc = exchange_sum(ctx, evals, NE);
        else rc = fetch_result(ctx, evals, NE);
        if (rc) return rc;
        if (need_plain_evals ? skip_now : skip1) evals[1] = f.sub(running, evals[0]);
        ip.coefficients(evals, coeffs);
        uint8_t bytes[32 * kMaxEvals];
        for (int i = 0; i < NE; ++i) f.to_bytes_le(coeffs[i], bytes + 32 * i);
        tr->t.append(bytes, 32 * NE);
        r = tr->t.challenge(f);
        running = f.horner(coeffs, NE, r);
        memcpy(coeffs_out + (size_t)k * NE * 4, coeffs, 32 * NE);
        memcpy(challenges_out + (size_t)k * 4, r.l, 32);
    }
    if (sharded) return fail(ctx, ZK_ERR_ARG, "internal: tables still sharded after the last round");
    int rc = launch_fold0(ctx, ptrs_of(sp), T, sp->len, make_fold_table(f, r));
    if (rc) return rc;
    set_len(sp, sp->len / 2);
    if (final_values) {
        for (int t = 0; t < T; ++t)
            ZK_CUDA(cudaMemcpyAsync(final_values + 4 * t, sp->tabs[t]->d, sizeof(Fe), cudaMemcpyDeviceToHost, ctx->stream));
    }
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

// basic_sumcheck Prover::prove (prover.rs:35-71) over a table sharded across the ranks (world 1: the whole table).
// The reference first absorbs the whole polynomial (prover.rs:38-39); with the table spread over several GPUs that
// absorb is the caller's: `tr` arrives with the table bytes already absorbed (every rank must pass an identical
// transcript).  From there on: claimed sum, n rounds of [sum left, sum right] -> absorb (big-endian) -> challenge ->
// fold, exactly as the single-GPU prover; identical outputs on every rank.
extern "C" int zk_prove_basic_sharded(zk_ctx* ctx, zk_table* local, zk_transcript* tr, uint64_t claimed_sum[4],
                                      uint64_t* round_polys, uint64_t* challenges, uint64_t final_value[4], uint32_t flags,
                                      uint64_t collapse_len) {
    if (ctx->world > 1 && !ctx->nccl_comm) return fail(ctx, ZK_ERR_ARG, "zk_comm_init has not been called");
    if (!is_pow2(local->len)) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");
    if (collapse_len < 1) collapse_len = 1;
    const HostField& f = ctx->field;
    const int G = ctx->world;
    const uint32_t n = ilog2(local->len) + ilog2((uint64_t)G);
    zk_sumpoly sp;   // a one-table view so that the shared helpers (collapse, set_len) apply
    sp.P = 1;
    sp.D = 1;
    sp.len = local->len;
    sp.tabs.assign(1, local);
    HFe evals[2], r = f.zero();
    bool sharded = G > 1;
    const bool dev_exchange = G > 1 && ctx->peers_attached && !(flags & (ZK_FLAG_NCCL_EXCHANGE | ZK_FLAG_HOST_ROUNDS | ZK_FLAG_HOST_EXCHANGE));
    if (dev_exchange) {
        int rc0 = agree_xseq(ctx);
        if (rc0) return rc0;
    }
    if (n == 0) {
        ZK_CUDA(cudaMemcpyAsync(claimed_sum, local->d, sizeof(Fe), cudaMemcpyDeviceToHost, ctx->stream));
        ZK_CUDA(cudaStreamSynchronize(ctx->stream));
        HFe c;
        memcpy(c.l, claimed_sum, 32);
        tr->t.append_be(f, c);
        if (final_value) memcpy(final_value, claimed_sum, 32);
        return ZK_OK;
    }
    for (uint32_t k = 0; k < n; ++k) {
        int rc;
        bool plain = (k == 0);
        if (k > 0 && sharded && sp.len / 2 <= collapse_len) {
            rc = launch_fold0(ctx, ptrs_of(&sp), 1, sp.len, make_fold_table(f, r));
            if (rc) return rc;
            set_len(&sp, sp.len / 2);
            if ((rc = collapse(ctx, &sp))) return rc;
            sharded = false;
            plain = true;
        }
        if (k == 0 && sharded && sp.len <= collapse_len) {
            if ((rc = collapse(ctx, &sp))) return rc;
            sharded = false;
        }
        // sharded rounds k >= 1 in one persistent launch per rank (round 0 stays on the host: its sums are the claimed sum)
        if (k > 0 && sharded && dev_exchange && dev_rounds_apply(ctx, sp.len, 1, flags) && sp.len / 2 > collapse_len) {
            const uint32_t want = sharded_rounds_from(sp.len, true, collapse_len);
            uint32_t ran = 0;
            rc = run_dev_rounds(ctx, ptrs_of(&sp), 1, 1, 0, kDevPlain, sp.len, &r, tr->t, round_polys + (size_t)k * 8,
                                challenges ? challenges + (size_t)k * 4 : nullptr, nullptr, want, true, &ran);
            if (rc) return rc;
            set_len(&sp, sp.len >> ran);
            k += ran - 1;
            memcpy(r.l, &ctx->dev_host->challenges[ran - 1], 32);
            continue;
        }
        TablePtrs tp = ptrs_of(&sp);
        if (k > 0 && !sharded && dev_rounds_apply(ctx, sp.len, 1, flags)) {   // the rest in one launch (the claimed sum is in)
            rc = run_dev_rounds(ctx, tp, 1, 1, 0, kDevPlain, sp.len, plain ? nullptr : &r, tr->t, round_polys + (size_t)k * 8,
                                challenges ? challenges + (size_t)k * 4 : nullptr, final_value);
            if (rc) return rc;
            set_len(&sp, 1);
            return ZK_OK;
        }
        const bool shared = sharded && ctx->xmail_host != nullptr && !(flags & ZK_FLAG_NCCL_EXCHANGE);
        if (plain) {
            rc = launch_round_evals(ctx, tp, 1, 1, sp.len, shared);                    // prover.rs:50
        } else {
            rc = launch_fold_evals(ctx, tp, 1, 1, sp.len, make_fold_table(f, r), false, shared);   // :61-63 fused with :50
            set_len(&sp, sp.len / 2);
        }
        if (rc) return rc;
        rc = sharded ? exchange_sum(ctx, evals, 2) : fetch_result(ctx, evals, 2);
        if (rc) return rc;
        if (k == 0) {                                                                  // init's claimed sum, prover.rs:28,40-41
            HFe claimed = f.add(evals[0], evals[1]);
            memcpy(claimed_sum, claimed.l, 32);
            tr->t.append_be(f, claimed);
        }
        uint8_t bytes[64];
        f.to_bytes_be(evals[0], bytes);
        f.to_bytes_be(evals[1], bytes + 32);
        tr->t.append(bytes, 64);                                                       // :51-55
        r = tr->t.challenge(f);                                                        // :58
        memcpy(round_polys + (size_t)k * 8, evals, 64);
        if (challenges) memcpy(challenges + (size_t)k * 4, r.l, 32);
    }
    if (sharded) return fail(ctx, ZK_ERR_ARG, "internal: table still sharded after the last round");
    int rc = launch_fold0(ctx, ptrs_of(&sp), 1, sp.len, make_fold_table(f, r));
    if (rc) return rc;
    set_len(&sp, sp.len / 2);
    if (final_value) ZK_CUDA(cudaMemcpyAsync(final_value, local->d, sizeof(Fe), cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

// MultilinearPolynomial::evaluate over a sharded table: each rank binds the leading n-g variables of
// its shard (no communication), the G survivors are gathered and the last g variables are bound on
// the host.  `values` holds all n challenges; every rank returns the same element.
extern "C" int zk_mle_evaluate_sharded(zk_ctx* ctx, const zk_table* local, const uint64_t* values, uint32_t n_values, uint64_t out[4]) {
    if (ctx->world == 1) return zk_mle_evaluate(ctx, local, values, n_values, out);
    if (!ctx->nccl_comm) return fail(ctx, ZK_ERR_ARG, "zk_comm_init has not been called");
    const int G = ctx->world;
    const uint32_t g = ilog2((uint64_t)G), nl = ilog2(zk_table_len(local));
    if (n_values != nl + g) return fail(ctx, ZK_ERR_ARG, "sharded evaluate needs exactly log2(global length) values");
    HFe mine;
    int rc = zk_mle_evaluate(ctx, local, values, nl, mine.l);
    if (rc) return rc;
    ZK_CUDA(cudaMemcpyAsync(ctx->xchg_send, mine.l, sizeof(Fe), cudaMemcpyHostToDevice, ctx->stream));
    ZK_NCCL(g_nccl.AllGather(ctx->xchg_send, ctx->xchg_recv, sizeof(Fe), ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    ZK_CUDA(cudaMemcpyAsync(ctx->xchg_host, ctx->xchg_recv, sizeof(Fe) * G, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    // the gathered vector is the table over the g low variables (entry q = shard q); bind them MSB first
    const HostField& f = ctx->field;
    std::vector<HFe> cur(reinterpret_cast<const HFe*>(ctx->xchg_host), reinterpret_cast<const HFe*>(ctx->xchg_host) + G);
    for (uint32_t i = 0; i < g; ++i) {
        HFe rr;
        memcpy(rr.l, values + 4 * (nl + i), 32);
        size_t half = cur.size() / 2;
        for (size_t j = 0; j < half; ++j) cur[j] = f.add(cur[j], f.mul(rr, f.sub(cur[j + half], cur[j])));
        cur.resize(half);
    }
    memcpy(out, cur[0].l, 32);
    return ZK_OK;
}
