// verify.cu -- the verifiers of the reference, over the same kernels.
//
//   basic_sumcheck::Verifier::verify        sumcheck_protocol/src/basic_sumcheck/verifier.rs:23-71
//   gkr_sumcheck::verify                    sumcheck_protocol/src/gkr_sumcheck/sumcheck_gkr_protocol.rs:69-106
//   gkr_protocol::verify + claim helpers    gkr/src/gkr_protocol.rs:146-236, gkr/src/utils.rs:84-135
//
// The transcript replay and the per-round checks are host work (a few field operations per round); what
// is data-parallel is the final oracle check of the plain sumcheck -- `initial_polynomial.evaluate(&challenges)`
// over the whole table (verifier.rs:67) -- and the table absorb, both done by the prover's kernels
// (fold_multi_kernel, to_bytes_be_kernel).  The GKR verifier's add_i / mul_i evaluations are computed from the
// gate list as  sum_g w(out_g) eq(r_b, left_g) eq(r_c, right_g)  instead of folding dense 2^(3i+2) tables.
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
inline uint32_t ilog2(uint64_t n) { uint32_t k = 0; while (n >>= 1) ++k; return k; }
inline uint32_t a_bits(uint32_t layer) { return layer == 0 ? 1 : layer; }
inline uint32_t bc_bits(uint32_t layer) { return layer + 1; }

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
// MultilinearPolynomial::evaluate on the host for the tiny tables of the protocol (round polynomials, outputs)
HFe host_mle_evaluate(const HostField& f, std::vector<HFe> cur, const HFe* rs, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) {
        size_t half = cur.size() / 2;
        for (size_t j = 0; j < half; ++j) cur[j] = f.add(cur[j], f.mul(rs[i], f.sub(cur[j + half], cur[j])));
        cur.resize(half ? half : 1);
    }
    return cur[0];
}
}  // namespace

// gkr_sumcheck::verify -- sumcheck_gkr_protocol.rs:69-106 (host only; needs no context)
extern "C" int zk_verify_product(int fid, const uint64_t claimed_sum[4], const uint64_t* coeffs, uint32_t n_rounds, uint32_t D,
                                 zk_transcript* tr, uint64_t* challenges, uint64_t last_claimed_sum[4], int* is_proof_valid) {
    if (fid < 0 || fid >= ZKF_NUM_FIELDS || D < 1 || D + 1 > (uint32_t)kMaxEvals) return ZK_ERR_ARG;
    HostField f(fid);
    HFe cur;
    memcpy(cur.l, claimed_sum, 32);
    tr->t.append_be(f, cur);                                                     // :73
    const HFe* c = reinterpret_cast<const HFe*>(coeffs);
    *is_proof_valid = 1;
    for (uint32_t k = 0; k < n_rounds; ++k, c += D + 1) {
        HFe e0 = f.horner(c, D + 1, f.zero()), e1 = f.horner(c, D + 1, f.one()); // :81-82
        if (f.add(e0, e1) != cur) {                                              // :84-90
            *is_proof_valid = 0;
            memcpy(last_claimed_sum, cur.l, 32);
            return ZK_OK;
        }
        uint8_t bytes[32 * kMaxEvals];
        for (uint32_t i = 0; i <= D; ++i) f.to_bytes_le(c[i], bytes + 32 * i);
        tr->t.append(bytes, 32 * (D + 1));                                       // :92
        HFe r = tr->t.challenge(f);                                              // :94
        cur = f.horner(c, D + 1, r);                                             // :96
        if (challenges) memcpy(challenges + 4 * k, r.l, 32);                     // :98
    }
    memcpy(last_claimed_sum, cur.l, 32);
    return ZK_OK;
}

// Verifier::verify -- verifier.rs:23-71.  `table` is the proof's initial_polynomial, resident on the GPU.
extern "C" int zk_verify_basic(zk_ctx* ctx, const zk_table* table, const uint64_t claimed_sum[4], const uint64_t* round_polys,
                               uint32_t n_rounds, int* ok) {
    const HostField& f = ctx->field;
    *ok = 0;
    const uint64_t len = zk_table_len(table);
    if (!is_pow2(len)) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");
    if (n_rounds != ilog2(len)) return ZK_OK;                                    // :26-30
    HFe claim;
    memcpy(claim.l, claimed_sum, 32);
    HostTranscript tr;
    {   // :34-35 convert_to_bytes + append, streamed from the GPU converter
        std::vector<uint8_t> bytes((size_t)len * 32);
        int rc = zk_mle_to_bytes(ctx, table, bytes.data());
        if (rc) return rc;
        tr.append(bytes.data(), bytes.size());
    }
    tr.append_be(f, claim);                                                      // :36-37
    std::vector<HFe> chal(n_rounds);
    const HFe* rp = reinterpret_cast<const HFe*>(round_polys);
    for (uint32_t i = 0; i < n_rounds; ++i, rp += 2) {
        // a 2-entry table evaluated at x is rp[0] + x (rp[1] - rp[0]): at 0 and 1 the entries themselves (:48-52)
        if (f.add(rp[0], rp[1]) != claim) return ZK_OK;
        tr.append_be(f, rp[0]);
        tr.append_be(f, rp[1]);                                                  // :58-59
        chal[i] = tr.challenge(f);                                               // :61
        claim = f.add(rp[0], f.mul(chal[i], f.sub(rp[1], rp[0])));               // :64
    }
    HFe fin;
    int rc = zk_mle_evaluate(ctx, table, n_rounds ? chal[0].l : nullptr, n_rounds, fin.l);   // :67 the oracle check, on the GPU
    if (rc) return rc;
    *ok = (fin == claim) ? 1 : 0;                                                // :70
    return ZK_OK;
}

// gkr_protocol::verify -- gkr_protocol.rs:146-236 for reference-shaped circuits; proof laid out as zk_gkr_prove writes it.
extern "C" int zk_gkr_verify(zk_ctx* ctx, const zk_circuit_desc* c, const uint64_t* output, uint64_t n_output,
                             const uint64_t* layer_claims, const uint64_t* coeffs, const uint64_t* wb, const uint64_t* wc,
                             const uint64_t* inputs, uint64_t n_inputs, int* ok) {
    const HostField& f = ctx->field;
    const uint32_t L = c->n_layers;
    *ok = 0;
    if (L == 0) return fail(ctx, ZK_ERR_ARG, "circuit has no layers");
    HostTranscript tr;
    std::vector<HFe> w0(reinterpret_cast<const HFe*>(output), reinterpret_cast<const HFe*>(output) + n_output);
    if (w0.size() == 1) w0.push_back(f.zero());                                  // :153-159
    if (!is_pow2(w0.size())) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");
    for (const HFe& x : w0) tr.append_be(f, x);                                  // :161
    HFe ra = tr.challenge(f);
    HFe claimed = host_mle_evaluate(f, w0, &ra, 1);                              // :164 (entry 0 after one fold)
    HFe alpha = f.zero(), beta = f.zero();
    std::vector<HFe> prev;
    uint64_t round_off = 0;
    for (uint32_t li = 0; li < L; ++li) {
        const uint32_t ab = a_bits(li), bcb = bc_bits(li), rounds = 2 * bcb;
        HFe layer_claim;
        memcpy(layer_claim.l, layer_claims + 4 * li, 32);
        if (claimed != layer_claim) return ZK_OK;                                // :167-169
        std::vector<HFe> chal(rounds);
        HFe last;
        int valid = 0;
        zk_transcript wrap;
        wrap.t = tr;
        int rc = zk_verify_product(ctx->fid, layer_claim.l, coeffs + 12 * round_off, rounds, 2, &wrap, chal[0].l, last.l, &valid);
        tr = wrap.t;
        if (rc) return rc;
        if (!valid) return ZK_OK;                                                // :172-176
        HFe wbv, wcv;
        if (li + 1 < L) {                                                        // :183-187
            memcpy(wbv.l, wb + 4 * li, 32);
            memcpy(wcv.l, wc + 4 * li, 32);
        } else {                                                                 // :188-194 the verifier knows the inputs
            if (!is_pow2(n_inputs)) return fail(ctx, ZK_ERR_ASSERT, "Evaluated values must be a power of 2");
            if (n_inputs != (1ull << bcb)) return ZK_OK;   // the reference would read past / short of the table here
            zk_table* t = nullptr;
            rc = zk_table_upload(ctx, inputs, n_inputs, &t);
            if (rc) return rc;
            rc = zk_mle_evaluate(ctx, t, chal[0].l, bcb, wbv.l);                 // utils.rs:70-82 evaluate_wb_wc
            if (!rc) rc = zk_mle_evaluate(ctx, t, chal[bcb].l, bcb, wcv.l);
            zk_table_free(ctx, t);
            if (rc) return rc;
        }
        // add_i(r), mul_i(r) at the sumcheck point with `a` bound as the prover bound it (utils.rs:84-135)
        std::vector<HFe> w;
        if (li == 0) {
            w = eq_table(f, &ra, 1);
        } else {
            const uint32_t pm = (uint32_t)prev.size() / 2;
            if (pm != ab) return fail(ctx, ZK_ERR_ARG, "internal: challenge split does not match the layer");
            std::vector<HFe> eb = eq_table(f, prev.data(), ab), ec = eq_table(f, prev.data() + pm, ab);
            w.resize(eb.size());
            for (size_t a = 0; a < w.size(); ++a) w[a] = f.add(f.mul(alpha, eb[a]), f.mul(beta, ec[a]));
        }
        std::vector<HFe> eqb = eq_table(f, chal.data(), bcb), eqc = eq_table(f, chal.data() + bcb, bcb);
        HFe add_r = f.zero(), mul_r = f.zero();
        std::map<std::pair<uint64_t, uint64_t>, bool> seen_add, seen_mul;         // dense indicators hold `= one`
        for (uint64_t g = c->layer_off[li]; g < c->layer_off[li + 1]; ++g) {
            uint64_t a = c->out[g], b = c->left[g], cc = c->right[g];
            if (a >= (1ull << ab) || b >= (1ull << bcb) || cc >= (1ull << bcb))
                return fail(ctx, ZK_ERR_ARG, "gate index does not fit the reference's layer shape");
            auto& seen = c->op[g] == 0 ? seen_add : seen_mul;
            if (seen.count({a, (b << bcb) | cc})) continue;
            seen[{a, (b << bcb) | cc}] = true;
            HFe term = f.mul(w[a], f.mul(eqb[b], eqc[cc]));
            if (c->op[g] == 0) add_r = f.add(add_r, term);
            else mul_r = f.add(mul_r, term);
        }
        HFe expected = f.add(f.mul(add_r, f.add(wbv, wcv)), f.mul(mul_r, f.mul(wbv, wcv)));   // utils.rs:110,134
        if (expected != last) return ZK_OK;                                      // :220-222
        prev = chal;                                                             // :224
        tr.append_be(f, wbv);
        alpha = tr.challenge(f);                                                 // :226-227
        tr.append_be(f, wcv);
        beta = tr.challenge(f);                                                  // :229-230
        claimed = f.add(f.mul(alpha, wbv), f.mul(beta, wcv));                    // :232
        round_off += rounds;
    }
    *ok = 1;
    return ZK_OK;
}
