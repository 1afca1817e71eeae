// gkr.cu -- the GKR layer prover (gkr/src/gkr_protocol.rs:26-143, gkr/src/utils.rs:8-82) on the GPU.
//
// Per layer i the reference builds f(b,c) = add_i(b,c) (W(b)+W(c)) + mul_i(b,c) (W(b) W(c)) as a 2x2
// SumPolynomial over 4^(i+1) entries and runs the product sumcheck on it.  Here:
//   * the circuit is evaluated on the host (arithmetic_circuit.rs:65-109; O(gates), the input model);
//   * add_i(b,c) / mul_i(b,c) with the output variables `a` already bound are NOT obtained by folding
//     dense 2^(3i+2) tables (utils.rs:23-68, gkr_protocol.rs:60-72): the bound value of a wiring
//     indicator at (b,c) is  sum over gates (out,b,c) of  w(out),  w(a) = eq(r_a, a)  for layer 0 and
//     w(a) = alpha eq(r_b, a) + beta eq(r_c, a)  afterwards -- the same field elements, computed from the
//     gate list (host, O(gates + 2^i)) and scattered into zeroed device tables;
//   * W(b)+W(c), W(b)W(c) are built by the tensor kernels (evaluation_form.rs:108-143) and the sumcheck
//     is the fused <P=2,D=2> round kernel; W(r_b), W(r_c) use the multi-level evaluate kernels.
// The transcript flow (absorb W_0, r_a, per-layer sumcheck, absorb W(r_b) -> alpha, W(r_c) -> beta) is the
// reference's, on the host.
#include <map>
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
inline bool is_pow2(uint64_t n) { return n && !(n & (n - 1)); }

// num_of_layer_variables split into (a bits, b/c bits) -- arithmetic_circuit.rs:166-178 (layer 0 prints
// a as the single digit "0": one a-bit)
inline uint32_t a_bits(uint32_t layer) { return layer == 0 ? 1 : layer; }
inline uint32_t bc_bits(uint32_t layer) { return layer + 1; }

// eq(r, .) over k variables, variable 0 = most significant index bit
std::vector<HFe> eq_table(const HostField& f, const HFe* r, uint32_t k) {
    std::vector<HFe> t(1, f.one());
    for (uint32_t v = 0; v < k; ++v) {
        std::vector<HFe> nxt(t.size() * 2);
        HFe one_minus = f.sub(f.one(), r[v]);
        for (size_t j = 0; j < t.size(); ++j) {
            nxt[2 * j] = f.mul(t[j], one_minus);
            nxt[2 * j + 1] = f.mul(t[j], r[v]);
        }
        t.swap(nxt);
    }
    return t;
}

__global__ void __launch_bounds__(kThreads) scatter_kernel(Fe* table, const uint64_t* pos, const Fe* val, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) st256(table + pos[i], val[i]);
}

// zero a table and scatter (position -> value) pairs with unique positions
int fill_sparse(zk_ctx* ctx, zk_table* t, const std::map<uint64_t, HFe>& entries) {
    ZK_CUDA(cudaMemsetAsync(t->d, 0, (size_t)t->len * sizeof(Fe), ctx->stream));
    if (entries.empty()) return ZK_OK;
    std::vector<uint64_t> pos;
    std::vector<HFe> val;
    pos.reserve(entries.size());
    val.reserve(entries.size());
    for (const auto& kv : entries) { pos.push_back(kv.first); val.push_back(kv.second); }
    size_t n = pos.size();
    size_t bytes = n * (sizeof(uint64_t) + sizeof(Fe)) + 64;
    int rc = ensure_scratch(ctx, bytes);
    if (rc) return rc;
    Fe* dval = (Fe*)ctx->scratch;
    uint64_t* dpos = (uint64_t*)((char*)ctx->scratch + n * sizeof(Fe));
    ZK_CUDA(cudaMemcpyAsync(dval, val.data(), n * sizeof(Fe), cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(cudaMemcpyAsync(dpos, pos.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    uint64_t blocks = (n + kThreads - 1) / kThreads;
    if (blocks > (uint64_t)ctx->sm_count * 4) blocks = (uint64_t)ctx->sm_count * 4;
    scatter_kernel<<<(int)blocks, kThreads, 0, ctx->stream>>>(t->d, dpos, dval, n);
    ctx->launches++;
    ZK_CUDA(cudaGetLastError());
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));   // host vectors go out of scope
    return ZK_OK;
}
}  // namespace

extern "C" uint64_t zk_gkr_total_rounds(uint32_t n_layers) {
    uint64_t s = 0;
    for (uint32_t i = 0; i < n_layers; ++i) s += 2 * bc_bits(i);
    return s;
}

// Circuit::evaluate -- arithmetic_circuit.rs:65-109 (host).  sizes: n_layers+1 entries; values: the layer
// value vectors concatenated, output layer first, inputs last.
extern "C" int zk_circuit_evaluate(int fid, const zk_circuit_desc* c, const uint64_t* inputs, uint64_t n_inputs,
                                   uint64_t* sizes, uint64_t* values, uint64_t values_cap) {
    if (fid < 0 || fid >= ZKF_NUM_FIELDS) return ZK_ERR_ARG;
    HostField f(fid);
    const uint32_t L = c->n_layers;
    sizes[L] = n_inputs;
    uint64_t total = n_inputs;
    for (uint32_t li = 0; li < L; ++li) {
        uint64_t mx = 0;
        for (uint64_t g = c->layer_off[li]; g < c->layer_off[li + 1]; ++g)
            if (c->out[g] > mx) mx = c->out[g];
        sizes[li] = mx + 1;                                                   // :73-80
        total += sizes[li];
    }
    if (total > values_cap) return ZK_ERR_ARG;
    std::vector<uint64_t> off(L + 2, 0);
    for (uint32_t i = 0; i <= L; ++i) off[i + 1] = off[i] + sizes[i];
    memcpy(values + 4 * off[L], inputs, n_inputs * 32);
    for (uint32_t li = L; li-- > 0;) {                                        // :72 layers walked input side first
        const HFe* in = reinterpret_cast<const HFe*>(values + 4 * off[li + 1]);
        HFe* out = reinterpret_cast<HFe*>(values + 4 * off[li]);
        for (uint64_t i = 0; i < sizes[li]; ++i) out[i] = f.zero();
        for (uint64_t g = c->layer_off[li]; g < c->layer_off[li + 1]; ++g) {
            if (c->left[g] >= sizes[li + 1] || c->right[g] >= sizes[li + 1]) return ZK_ERR_ASSERT;   // index out of bounds
            HFe v = c->op[g] == 0 ? f.add(in[c->left[g]], in[c->right[g]]) : f.mul(in[c->left[g]], in[c->right[g]]);
            out[c->out[g]] = f.add(out[c->out[g]], v);                        // :96 accumulates
        }
    }
    return ZK_OK;
}

// gkr_protocol::prove -- gkr_protocol.rs:26-143.  Flat proof layout (include/zk_sumcheck.h):
// output (n_output elements), claimed_sum, layer_claims[L], coeffs[rounds x 3], challenges[rounds],
// wb[L-1], wc[L-1].
extern "C" int zk_gkr_prove(zk_ctx* ctx, const zk_circuit_desc* c, const uint64_t* inputs, uint64_t n_inputs,
                            uint64_t* output, uint64_t output_cap, uint64_t* n_output, uint64_t claimed_sum[4],
                            uint64_t* layer_claims, uint64_t* coeffs_out, uint64_t* challenges_out, uint64_t* wb_out,
                            uint64_t* wc_out) {
    const HostField& f = ctx->field;
    const uint32_t L = c->n_layers;
    if (L == 0) return fail(ctx, ZK_ERR_ARG, "circuit has no layers");
    // ---- circuit.evaluate(inputs) :27
    std::vector<uint64_t> sizes(L + 1);
    uint64_t cap = n_inputs;
    for (uint32_t li = 0; li < L; ++li) {
        uint64_t mx = 0;
        for (uint64_t g = c->layer_off[li]; g < c->layer_off[li + 1]; ++g)
            if (c->out[g] > mx) mx = c->out[g];
        cap += mx + 1;
    }
    std::vector<uint64_t> values(cap * 4);
    int rc = zk_circuit_evaluate(ctx->fid, c, inputs, n_inputs, sizes.data(), values.data(), cap);
    if (rc == ZK_ERR_ASSERT) return fail(ctx, rc, "index out of bounds");
    if (rc) return fail(ctx, rc, "circuit evaluation failed");
    std::vector<uint64_t> off(L + 2, 0);
    for (uint32_t i = 0; i <= L; ++i) off[i + 1] = off[i] + sizes[i];
    for (uint32_t i = 0; i <= L; ++i)
        if (!is_pow2(sizes[i])) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");   // w_i_polynomial -> new
    if (sizes[0] > output_cap) return fail(ctx, ZK_ERR_ARG, "output buffer too small");
    memcpy(output, values.data(), sizes[0] * 32);
    *n_output = sizes[0];

    // ---- layer 0 :39-51
    HostTranscript tr;
    std::vector<HFe> w0(reinterpret_cast<HFe*>(values.data()), reinterpret_cast<HFe*>(values.data()) + sizes[0]);
    if (w0.size() == 1) w0.push_back(f.zero());                                // :43-47
    for (const HFe& x : w0) tr.append_be(f, x);                                // :49 convert_to_bytes
    HFe ra = tr.challenge(f);                                                  // :50
    // the reference binds ONE output variable (evaluate(&[r_a]), :51): its circuits have at most two outputs
    if (w0.size() != 2) return fail(ctx, ZK_ERR_ARG, "reference-shaped circuits have at most two outputs (layer 0 has one output bit)");
    HFe claim = f.add(w0[0], f.mul(ra, f.sub(w0[1], w0[0])));                  // :51 W_0(r_a)

    // one pool for the four 4^(i+1)-entry tables and W of the widest layer
    const uint64_t nbc_max = 1ull << (2 * bc_bits(L - 1)), w_max = 1ull << bc_bits(L - 1);
    ZK_CUDA(cudaSetDevice(ctx->device));
    if ((rc = ensure_pool(ctx, (size_t)(4 * nbc_max + w_max) * sizeof(Fe)))) return rc;   // kept in the context across proves
    Fe* pool = (Fe*)ctx->pool;
    zk_table view[5];
    for (int i = 0; i < 5; ++i) {
        view[i].d = pool + (size_t)i * nbc_max;
        view[i].owned = false;
    }

    HFe alpha = f.zero(), beta = f.zero();
    std::vector<HFe> rb, rcv;
    uint64_t round_off = 0;
    for (uint32_t li = 0; li < L; ++li) {                                      // :57
        const uint32_t ab = a_bits(li), bcb = bc_bits(li), rounds = 2 * bcb;
        const uint64_t wlen = sizes[li + 1], nbc = 1ull << rounds;
        if (wlen != (1ull << bcb)) return fail(ctx, ZK_ERR_ASSERT, "different number of variables");   // SumPolynomial::new
        // ---- weights of the output variable: layer 0 eq(r_a, .), later alpha eq(r_b, .) + beta eq(r_c, .)
        std::vector<HFe> w;
        if (li == 0) {
            w = eq_table(f, &ra, 1);                                           // :60-72
        } else {
            if (rb.size() != ab) return fail(ctx, ZK_ERR_ARG, "internal: challenge split does not match the layer");
            std::vector<HFe> eb = eq_table(f, rb.data(), ab), ec = eq_table(f, rcv.data(), ab);
            w.resize(eb.size());
            for (size_t a = 0; a < w.size(); ++a) w[a] = f.add(f.mul(alpha, eb[a]), f.mul(beta, ec[a]));   // utils.rs:59-66
        }
        // ---- add_i(b,c), mul_i(b,c) with `a` bound: scatter over the (deduplicated) gate list
        std::map<uint64_t, HFe> add_e, mul_e;
        {
            std::map<std::pair<uint64_t, uint64_t>, bool> seen_add, seen_mul;   // (a, bc): the dense tables hold `= one`, not `+=`
            for (uint64_t g = c->layer_off[li]; g < c->layer_off[li + 1]; ++g) {
                uint64_t a = c->out[g], b = c->left[g], cc = c->right[g];
                if (a >= (1ull << ab) || b >= (1ull << bcb) || cc >= (1ull << bcb))
                    return fail(ctx, ZK_ERR_ARG, "gate index does not fit the reference's layer shape");
                uint64_t bc = (b << bcb) | cc;
                auto& seen = c->op[g] == 0 ? seen_add : seen_mul;
                auto& ent = c->op[g] == 0 ? add_e : mul_e;
                if (seen.count({a, bc})) continue;
                seen[{a, bc}] = true;
                auto it = ent.find(bc);
                if (it == ent.end()) ent[bc] = w[a];
                else it->second = f.add(it->second, w[a]);
            }
        }
        // ---- the four tables of compute_fbc_polynomial (utils.rs:8-21)
        zk_table *t_add = &view[0], *t_wadd = &view[1], *t_mul = &view[2], *t_wmul = &view[3], *t_w = &view[4];
        for (int i = 0; i < 4; ++i) view[i].len = view[i].cap = nbc;
        t_w->len = t_w->cap = wlen;
        if ((rc = fill_sparse(ctx, t_add, add_e))) return rc;
        if ((rc = fill_sparse(ctx, t_mul, mul_e))) return rc;
        ZK_CUDA(cudaMemcpyAsync(t_w->d, values.data() + 4 * off[li + 1], wlen * sizeof(Fe), cudaMemcpyHostToDevice, ctx->stream));   // :88-89 W_b = W_c
        if ((rc = tensor_into(ctx, t_w->d, t_w->d, wlen, t_wadd->d, EW_ADD))) return rc;
        if ((rc = tensor_into(ctx, t_w->d, t_w->d, wlen, t_wmul->d, EW_MUL))) return rc;
        zk_sumpoly sp_obj;
        sp_obj.P = 2;
        sp_obj.D = 2;
        sp_obj.len = nbc;
        sp_obj.tabs = {t_add, t_wadd, t_mul, t_wmul};
        zk_sumpoly* sp = &sp_obj;
        // ---- sumcheck_prove(f_bc, claimed_sum, &mut transcript) :99
        memcpy(layer_claims + 4 * li, claim.l, 32);
        zk_transcript wrap;
        wrap.t = tr;
        uint64_t* chal = challenges_out + 4 * round_off;
        rc = zk_prove_product(ctx, sp, claim.l, &wrap, coeffs_out + 12 * round_off, chal, nullptr, 0);
        tr = wrap.t;
        if (rc) return rc;
        if (li + 1 < L) {                                                      // :109
            HFe wbv, wcv;
            if ((rc = zk_mle_evaluate(ctx, t_w, chal, bcb, wbv.l))) return rc;                                           // utils.rs:70-82
            if ((rc = zk_mle_evaluate(ctx, t_w, chal + 4 * bcb, bcb, wcv.l))) return rc;
            memcpy(wb_out + 4 * li, wbv.l, 32);
            memcpy(wc_out + 4 * li, wcv.l, 32);
            rb.assign(reinterpret_cast<HFe*>(chal), reinterpret_cast<HFe*>(chal) + bcb);                                 // :120-123
            rcv.assign(reinterpret_cast<HFe*>(chal) + bcb, reinterpret_cast<HFe*>(chal) + 2 * bcb);
            tr.append_be(f, wbv);
            alpha = tr.challenge(f);                                           // :125-126
            tr.append_be(f, wcv);
            beta = tr.challenge(f);                                            // :128-129
            claim = f.add(f.mul(alpha, wbv), f.mul(beta, wcv));                // :132
        }
        round_off += rounds;
    }
    memcpy(claimed_sum, claim.l, 32);
    return ZK_OK;
}
