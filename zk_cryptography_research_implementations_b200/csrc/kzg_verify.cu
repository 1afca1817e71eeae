// kzg_verify.cu -- the verifier's half of the multilinear KZG, host only (no context, no GPU): the G2 side of the trusted
// setup (multilinear_kzg/src/trusted_setup.rs:65-78 compute_g2_powers_of_tau) and MultilinearKZG::verify
// (multilinear_kzg/src/multilinear_kzg.rs:132-159) with the pairing of host_pairing.h.  The prover never calls this.
#include <vector>

#include "../../include/zk_sumcheck.h"
#include "host_field.h"
#include "host_pairing.h"

using namespace zk;

namespace {
const HostField& fr() {
    static const HostField f(2 /* BLS12_381_FR */);
    return f;
}
void canonical(const uint64_t mont[4], uint64_t out[4]) {
    HFe a;
    memcpy(a.l, mont, 32);
    const HFe c = fr().from_mont(a);
    memcpy(out, c.l, 32);
}
HG1Affine g1_at(const uint64_t* p) {
    HG1Affine r;
    memcpy(&r, p, sizeof r);
    return r;
}
HG2Affine g2_at(const uint64_t* p) {
    HG2Affine r;
    memcpy(&r, p, sizeof r);
    return r;
}
}  // namespace

// g2_powers_of_tau[i] = tau_i * G2 (trusted_setup.rs:65-78); out: n points of 24 limbs (x.c0, x.c1, y.c0, y.c1)
extern "C" int zk_kzg_g2_powers_of_tau(const uint64_t* taus, uint32_t n, uint64_t* out) {
    static_assert(sizeof(HG2Affine) == 24 * sizeof(uint64_t), "G2 affine layout");
    if (n == 0) return ZK_ERR_ASSERT;   // "requires at least one variable"
    const HG2Affine g2 = HostG2::generator();
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t k[4];
        canonical(taus + 4 * i, k);
        const HG2Affine p = HostG2::mul(g2, k);
        memcpy(out + 24 * i, &p, sizeof p);
    }
    return ZK_OK;
}

// MultilinearKZG::verify (multilinear_kzg.rs:132-159).  ZK_ERR_ASSERT: "Number of opening values must match number of proofs";
// ZK_ERR_ARG: a point off its curve.  The reference zips the setup's g2 powers with the proofs (:148-154): n_g2 entries are used.
extern "C" int zk_kzg_verify(const uint64_t* g2_powers_of_tau, uint32_t n_g2, const uint64_t commitment[12], const uint64_t* opening_values,
                             uint32_t n_opening, const uint64_t evaluation[4], const uint64_t* proofs, uint32_t n_proofs, int* ok) {
    *ok = 0;
    if (n_opening != n_proofs) return ZK_ERR_ASSERT;
    if (n_g2 > n_proofs) return ZK_ERR_ARG;                      // the reference would index past `proofs`
    const HG1Affine c = g1_at(commitment);
    if (!HostG1::on_curve(c)) return ZK_ERR_ARG;
    // lhs point: C - v G
    uint64_t v[4];
    canonical(evaluation, v);
    const HG1Xyzz vg = HostG1::mul(HostG1::from_affine(HostG1::generator()), v, 256);
    const HG1Affine lhs = HostG1::to_affine(HostG1::add(HostG1::from_affine(c), HostG1::neg(vg)));
    std::vector<HG1Affine> ps;
    std::vector<HG2Affine> qs;
    ps.push_back(lhs);
    qs.push_back(HostG2::generator());
    const HG2Affine g2 = HostG2::generator();
    for (uint32_t i = 0; i < n_g2; ++i) {
        const HG1Affine q = g1_at(proofs + 12 * i);
        const HG2Affine tau = g2_at(g2_powers_of_tau + 24 * i);
        if (!HostG1::on_curve(q) || !HostG2::on_curve(tau)) return ZK_ERR_ARG;
        uint64_t r[4];
        canonical(opening_values + 4 * i, r);
        const HG2Affine rhs = HostG2::add_affine(tau, HostG2::neg(HostG2::mul(g2, r)));   // tau_i G2 - r_i G2
        HG1Affine nq = q;                                                                 // e(-Q_i, .) on the left
        nq.y = HostFq::neg(q.y);
        if (q.is_inf()) nq = q;
        ps.push_back(nq);
        qs.push_back(rhs);
    }
    *ok = HostPairing::product_is_one(ps, qs) ? 1 : 0;
    return ZK_OK;
}

// prod_i e(g1[i], g2[i]) == 1 (one product of Miller loops, one final exponentiation): the primitive zk_kzg_verify is built on,
// exported for callers that batch several checks into one.  ZK_ERR_ARG: a point off its curve.
extern "C" int zk_pairing_product_is_one(const uint64_t* g1_points, const uint64_t* g2_points, uint32_t n, int* ok) {
    *ok = 0;
    std::vector<HG1Affine> ps;
    std::vector<HG2Affine> qs;
    for (uint32_t i = 0; i < n; ++i) {
        const HG1Affine p = g1_at(g1_points + 12 * i);
        const HG2Affine q = g2_at(g2_points + 24 * i);
        if (!HostG1::on_curve(p) || !HostG2::on_curve(q)) return ZK_ERR_ARG;
        ps.push_back(p);
        qs.push_back(q);
    }
    *ok = HostPairing::product_is_one(ps, qs) ? 1 : 0;
    return ZK_OK;
}

// host helpers for callers that hold points as plain integers (tests, bindings)
extern "C" int zk_g1_is_on_curve(const uint64_t p[12]) { return HostG1::on_curve(g1_at(p)) ? 1 : 0; }
extern "C" void zk_g1_generator(uint64_t out[12]) {
    const HG1Affine g = HostG1::generator();
    memcpy(out, &g, sizeof g);
}
extern "C" void zk_g2_generator(uint64_t out[24]) {
    const HG2Affine g = HostG2::generator();
    memcpy(out, &g, sizeof g);
}
