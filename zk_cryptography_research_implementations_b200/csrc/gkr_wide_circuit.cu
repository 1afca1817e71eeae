// gkr_wide_circuit.cu -- construction of the device-resident layered circuit of the wide GKR prover.
//
// The reference builds each layer's wiring inside `prove` (circuit/src/arithmetic_circuit.rs:126-163 add_i_and_mul_i_mle,
// called from gkr/src/gkr_protocol.rs:58) as dense 2^(3i+2) indicator tables.  Here the wiring stays a gate list, and
// what the prover needs from it -- the gates grouped by left input, by right input and by output (three CSR orderings
// per layer) -- is built ON THE GPU from the caller's flat arrays: one upload of the layer's (left, right, out, op),
// a range check, a duplicate check (the dense indicator stores `= one`, so a repeated gate would make prover and
// wiring predicate disagree -- refused), and per ordering a histogram (64-bit atomics), an exclusive scan and a scatter.
// Nothing is sorted on the host.  The order of the gates inside one bucket is whatever the scatter's atomics produce;
// every consumer sums a bucket with exact (canonical) field additions, so the proof does not depend on it.
#include <algorithm>
#include <string>
#include <vector>

#include "gkr_wide.h"
#include "scan.cuh"

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
using zk::scan::u64;
using zk::scan::kScanChunk;

inline int blocks_for(const zk_ctx* ctx, uint64_t work, int bps = 8) {
    uint64_t blocks = (work + kThreads - 1) / kThreads, cap = (uint64_t)ctx->sm_count * bps;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

// flags[0] |= 1: an index does not fit its layer; |= 2: an operator other than 0 / 1
__global__ void __launch_bounds__(kThreads) validate_gates_kernel(const uint32_t* left, const uint32_t* right, const uint32_t* out, const uint8_t* op,
                                                                  uint64_t n, uint64_t n_in, uint64_t n_out, unsigned* flags) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned bad = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (left[i] >= n_in || right[i] >= n_in || out[i] >= n_out) bad |= 1u;
        if (op[i] > 1) bad |= 2u;
    }
    if (bad) atomicOr(flags, bad);
}

// duplicate detection: open-addressing table of gate numbers keyed by a 64-bit mix of (out, left, right, op); two
// gates that meet in a slot are compared field by field.  slots: power-of-two count >= 2 n, zero = empty.
__device__ __forceinline__ u64 mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(kThreads) find_duplicates_kernel(const uint32_t* left, const uint32_t* right, const uint32_t* out, const uint8_t* op,
                                                                   uint64_t n, unsigned* slots, uint64_t mask, unsigned* flags) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t l = left[i], r = right[i], o = out[i], p = op[i];
        u64 h = mix64(((u64)l << 32 | r) ^ mix64(((u64)o << 8) | p)) & mask;
        for (;;) {
            const unsigned old = atomicCAS(slots + h, 0u, (unsigned)i + 1u);
            if (old == 0u) break;
            const uint64_t j = old - 1u;
            if (left[j] == l && right[j] == r && out[j] == o && op[j] == p) {
                atomicOr(flags, 4u);
                break;
            }
            h = (h + 1) & mask;
        }
    }
}

// off[key + 1] += 1 per gate (off zeroed before)
__global__ void __launch_bounds__(kThreads) histogram_kernel(const uint32_t* key, uint64_t n, uint64_t n_keys, u64* off) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (key[i] < n_keys) atomicAdd(off + key[i] + 1, 1ull);   // an out-of-range key was flagged by validate_gates_kernel: never an array position
}

// gate i goes to the next free place of its bucket (cursor starts as a copy of off[0..n_keys))
__global__ void __launch_bounds__(kThreads) scatter_gates_kernel(const uint32_t* key, const uint32_t* x, const uint32_t* y, const uint8_t* op, uint64_t n,
                                                                 uint64_t n_keys, u64* cursor, uint32_t* sx, uint32_t* sy, uint8_t* so) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (key[i] >= n_keys) continue;
        const u64 p = atomicAdd(cursor + key[i], 1ull);
        sx[p] = x[i];
        sy[p] = y[i];
        so[p] = op[i];
    }
}

struct Staging {   // one layer's flat gate arrays on the device + scratch of the builders
    uint32_t *left = nullptr, *right = nullptr, *out = nullptr;
    uint8_t* op = nullptr;
    u64 *cursor = nullptr, *chunk_tot = nullptr;
    unsigned *slots = nullptr, *flags = nullptr;
    ~Staging() {
        cudaFree(left); cudaFree(right); cudaFree(out); cudaFree(op);
        cudaFree(cursor); cudaFree(chunk_tot); cudaFree(slots); cudaFree(flags);
    }
};

// `out`'s arrays are already carved out of the circuit's pools (zk_wide_circuit_create)
int build_csr_device(zk_ctx* ctx, Staging& st, uint64_t n_keys, uint64_t n, const uint32_t* key, const uint32_t* x, const uint32_t* y, GateCsr* out) {
    u64* off = reinterpret_cast<u64*>(out->off);
    ZK_CUDA(cudaMemsetAsync(off, 0, (n_keys + 1) * sizeof(u64), ctx->stream));
    if (n) histogram_kernel<<<blocks_for(ctx, n), kThreads, 0, ctx->stream>>>(key, n, n_keys, off);
    zk::scan::inclusive_scan(ctx->stream, off, n_keys + 1, st.chunk_tot);
    ZK_CUDA(cudaMemcpyAsync(st.cursor, off, n_keys * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
    if (n) scatter_gates_kernel<<<blocks_for(ctx, n), kThreads, 0, ctx->stream>>>(key, x, y, st.op, n, n_keys, st.cursor, out->x, out->y, out->op);
    ctx->launches += 5;
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}
}  // namespace

extern "C" int zk_wide_circuit_create(zk_ctx* ctx, uint32_t n_layers, const uint32_t* layer_bits, const uint64_t* layer_off,
                                      const uint32_t* left, const uint32_t* right, const uint32_t* out, const uint8_t* op,
                                      zk_wide_circuit** result) {
    if (n_layers == 0) return fail(ctx, ZK_ERR_ARG, "circuit has no layers");
    for (uint32_t i = 0; i <= n_layers; ++i)
        if (layer_bits[i] > 30) return fail(ctx, ZK_ERR_ARG, "layer wider than 2^30");
    for (uint32_t i = 1; i <= n_layers; ++i)
        if (layer_bits[i] == 0) return fail(ctx, ZK_ERR_ARG, "every layer must read at least two wires");
    uint64_t max_gates = 0, max_keys = 0;
    for (uint32_t li = 0; li < n_layers; ++li) {
        if (layer_off[li + 1] < layer_off[li]) return fail(ctx, ZK_ERR_ARG, "layer_off must be non-decreasing");
        max_gates = std::max<uint64_t>(max_gates, layer_off[li + 1] - layer_off[li]);
    }
    if (max_gates >= (1ull << 32) - 1) return fail(ctx, ZK_ERR_ARG, "more than 2^32 - 2 gates in one layer");
    ZK_CUDA(cudaSetDevice(ctx->device));
    zk_wide_circuit* wc = new zk_wide_circuit();
    wc->L = n_layers;
    wc->device = ctx->device;
    wc->bits.assign(layer_bits, layer_bits + n_layers + 1);
    // a single output is the reference's padded output layer [out, 0] with ONE challenge r_a (gkr_protocol.rs:43-51):
    // a one-bit layer whose index 1 no gate drives
    if (wc->bits[0] == 0) wc->bits[0] = 1;
    for (uint32_t b : wc->bits) max_keys = std::max<uint64_t>(max_keys, 1ull << b);
    wc->layers.resize(n_layers);
    int rc = ZK_OK;
    {   // the CSR arrays of all layers and orderings come out of four pools: four allocations per circuit, not 12 per layer
        uint64_t n_off = 0, n_g = 0;
        for (uint32_t li = 0; li < n_layers; ++li) {
            const uint64_t n = layer_off[li + 1] - layer_off[li];
            n_off += 2 * ((1ull << wc->bits[li + 1]) + 1) + (1ull << wc->bits[li]) + 1;
            n_g += 3 * (n ? n : 1);
        }
        cudaError_t e = cudaMalloc(&wc->pool_off, n_off * sizeof(uint64_t));
        if (e == cudaSuccess) e = cudaMalloc(&wc->pool_x, n_g * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&wc->pool_y, n_g * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&wc->pool_op, n_g);
        if (e != cudaSuccess) {
            ctx->err = std::string("cudaMalloc (circuit): ") + cudaGetErrorString(e);
            zk_wide_circuit_free(ctx, wc);
            return ZK_ERR_CUDA;
        }
        uint64_t o = 0, g = 0;
        for (uint32_t li = 0; li < n_layers; ++li) {
            const uint64_t n = layer_off[li + 1] - layer_off[li], cap = n ? n : 1;
            const uint64_t keys[3] = {1ull << wc->bits[li + 1], 1ull << wc->bits[li + 1], 1ull << wc->bits[li]};
            GateCsr* csr[3] = {&wc->layers[li].by_left, &wc->layers[li].by_right, &wc->layers[li].by_out};
            for (int k = 0; k < 3; ++k) {
                csr[k]->off = wc->pool_off + o;
                csr[k]->x = wc->pool_x + g;
                csr[k]->y = wc->pool_y + g;
                csr[k]->op = wc->pool_op + g;
                o += keys[k] + 1;
                g += cap;
            }
        }
    }
    {
        Staging st;
        const uint64_t cap = max_gates ? max_gates : 1;
        uint64_t n_slots = 2;
        while (n_slots < 2 * cap) n_slots <<= 1;
        cudaError_t e = cudaMalloc(&st.left, cap * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&st.right, cap * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&st.out, cap * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&st.op, cap);
        if (e == cudaSuccess) e = cudaMalloc(&st.cursor, (max_keys + 1) * sizeof(u64));
        if (e == cudaSuccess) e = cudaMalloc(&st.chunk_tot, ((max_keys + 1 + kScanChunk - 1) / kScanChunk + 1) * sizeof(u64));
        if (e == cudaSuccess) e = cudaMalloc(&st.slots, n_slots * sizeof(unsigned));
        if (e == cudaSuccess) e = cudaMalloc(&st.flags, sizeof(unsigned));
        if (e != cudaSuccess) {
            ctx->err = std::string("cudaMalloc (circuit staging): ") + cudaGetErrorString(e);
            zk_wide_circuit_free(ctx, wc);
            return ZK_ERR_CUDA;
        }
        cudaMemsetAsync(st.flags, 0, sizeof(unsigned), ctx->stream);
        for (uint32_t li = 0; li < n_layers && rc == ZK_OK; ++li) {
            const uint64_t g0 = layer_off[li], n = layer_off[li + 1] - g0;
            // the caller's declared width: a padded single output still only has index 0 driven
            const uint64_t n_out = 1ull << wc->bits[li], n_in = 1ull << wc->bits[li + 1];
            const uint64_t n_out_valid = (li == 0 && layer_bits[0] == 0) ? 1 : n_out;
            WideLayer& wl = wc->layers[li];
            wl.n_gates = n;
            auto up = [&](void* dst, const void* src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream); };
            e = up(st.left, left + g0, n * sizeof(uint32_t));
            if (e == cudaSuccess) e = up(st.right, right + g0, n * sizeof(uint32_t));
            if (e == cudaSuccess) e = up(st.out, out + g0, n * sizeof(uint32_t));
            if (e == cudaSuccess) e = up(st.op, op + g0, n);
            if (e != cudaSuccess) { ctx->err = std::string("cudaMemcpyAsync (gates): ") + cudaGetErrorString(e); rc = ZK_ERR_CUDA; break; }
            if (n) {
                // every index is range-checked; the builders below never use an out-of-range index as an array position
                // (they skip it), so the verdict can be read once, after the last layer, without a sync per layer
                validate_gates_kernel<<<blocks_for(ctx, n), kThreads, 0, ctx->stream>>>(st.left, st.right, st.out, st.op, n, n_in, n_out_valid, st.flags);
                uint64_t slots_n = 2;
                while (slots_n < 2 * n) slots_n <<= 1;
                cudaMemsetAsync(st.slots, 0, slots_n * sizeof(unsigned), ctx->stream);
                find_duplicates_kernel<<<blocks_for(ctx, n), kThreads, 0, ctx->stream>>>(st.left, st.right, st.out, st.op, n, st.slots, slots_n - 1, st.flags);
                ctx->launches += 2;
            }
            rc = build_csr_device(ctx, st, n_in, n, st.left, st.out, st.right, &wl.by_left);
            if (!rc) rc = build_csr_device(ctx, st, n_in, n, st.right, st.out, st.left, &wl.by_right);
            if (!rc) rc = build_csr_device(ctx, st, n_out, n, st.out, st.left, st.right, &wl.by_out);
            if (rc) break;
        }
        if (rc == ZK_OK) {
            unsigned flags = 0;
            cudaMemcpyAsync(&flags, st.flags, sizeof flags, cudaMemcpyDeviceToHost, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, ZK_ERR_CUDA, "circuit construction kernels failed");
            else if (flags & 1u) rc = fail(ctx, ZK_ERR_ARG, "gate index does not fit its layer width");
            else if (flags & 2u) rc = fail(ctx, ZK_ERR_ARG, "gate operator must be 0 (add) or 1 (mul)");
            else if (flags & 4u) rc = fail(ctx, ZK_ERR_ARG, "duplicate gate: the wide prover needs a duplicate-free gate list (the reference's dense wiring tables store `= one`)");
        }
    }
    if (rc) { zk_wide_circuit_free(ctx, wc); return rc; }
    {   // workspace
        uint32_t maxbits = 0;
        for (uint32_t b : wc->bits) maxbits = std::max(maxbits, b);
        const uint64_t maxn = 1ull << maxbits;
        wc->W.resize(n_layers + 1);
        cudaError_t e = cudaSuccess;
        for (uint32_t li = 0; li <= n_layers && e == cudaSuccess; ++li) e = wc->W[li].alloc(1ull << wc->bits[li]);
        DevBuf* bufs[] = {&wc->wtab, &wc->eqa, &wc->h1, &wc->h2, &wc->Wc};
        for (DevBuf* b : bufs)
            if (e == cudaSuccess) e = b->alloc(maxn);
        DevBuf* halves[] = {&wc->half_hi, &wc->half_lo, &wc->half_hi2, &wc->half_lo2, &wc->pre_eh};   // half_hi doubles as the slice scratch of evaluate
        for (DevBuf* b : halves)
            if (e == cudaSuccess) e = b->alloc(1ull << 16);
        if (e == cudaSuccess) e = wc->pre_pg.alloc(max_gates ? max_gates : 1);
        if (e != cudaSuccess) {
            ctx->err = std::string("cudaMalloc (GKR workspace): ") + cudaGetErrorString(e);
            zk_wide_circuit_free(ctx, wc);
            return ZK_ERR_CUDA;
        }
    }
    *result = wc;
    return ZK_OK;
}

extern "C" void zk_wide_circuit_free(zk_ctx* ctx, zk_wide_circuit* wc) {
    if (!wc) return;
    cudaSetDevice(wc->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(wc->pool_off);
    cudaFree(wc->pool_x);
    cudaFree(wc->pool_y);
    cudaFree(wc->pool_op);
    delete wc;
}

extern "C" uint64_t zk_wide_circuit_total_rounds(const zk_wide_circuit* wc) {
    uint64_t s = 0;
    for (uint32_t li = 0; li < wc->L; ++li) s += 2ull * wc->bits[li + 1];
    return s;
}
extern "C" uint32_t zk_wide_circuit_output_bits(const zk_wide_circuit* wc) { return wc->bits[0]; }
