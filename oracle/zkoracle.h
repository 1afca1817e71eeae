/*
 * zkoracle.h -- CPU restatement of the reference's sumcheck / GKR prover path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the CUDA library under
 * zk_cryptography_research_implementations_b200/) may include, link or call this.
 * It is used by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
 * `--impl reference` leg, always as the checker or the timed CPU baseline.
 *
 * The reference (casweeney/zk-cryptography-research-implementations) is Rust on
 * arkworks 0.5.0 and cannot be compiled in this image (no cargo/rustc).  Its field
 * arithmetic lives in third-party crates that are NOT vendored under /root/reference:
 *   ark-ff 0.5.0, ark-bn254 0.5.0, ark-bls12-381 0.5.0 (Cargo.lock:45-105),
 *   sha3 0.10.8 / keccak 0.1.5 (Cargo.lock:566-572,859-866).
 * Their published algorithms are restated here: 4x64-bit-limb Montgomery prime
 * fields with R = 2^256, Keccak-256 (original 0x01 padding, rate 136).
 *
 * PARITY STATUS: the small-integer known-answer tests of the reference
 * (evaluation_form.rs:179-220, product_polynomial.rs:107-173, sum_polynomial.rs:117-245,
 * dense_univariate.rs:206-261, sumcheck_gkr_protocol.rs:164-186, protocol.rs:10-26,
 * prover.rs:101-107, arithmetic_circuit.rs:219-384) are all reproduced (tests/test_oracle_*.py).
 * Everything that flows through the Fiat-Shamir transcript (challenges, later round
 * polynomials, proofs) is "parity unpinned" BY THE REFERENCE ITSELF -- it only asserts
 * verify()==true.  Those are pinned here by public Keccak-256 vectors, by an independent
 * Python big-integer model (oracle/pyoracle.py), by the prove->verify round trips and by
 * the survey-derived vectors of SURVEY.md appendix B (tests/golden/).
 *
 * Element layout everywhere: 4 x uint64 little-endian limbs, MONTGOMERY form, canonical (< p)
 * -- exactly arkworks' Fp<MontBackend<_,4>,4>.
 */
#ifndef ZK_ORACLE_H
#define ZK_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ZKO_BN254_FQ = 0, ZKO_BN254_FR = 1, ZKO_BLS12_381_FR = 2 };

/* ---- all-core mode for the timed CPU baseline (default 1 thread = the reference's behaviour; see zkoracle.c) ---- */
void zko_set_threads(int n);
int  zko_get_threads(void);
int  zko_openmp_enabled(void);    /* 0: built without OpenMP, zko_set_threads has no effect */

/* ---- field (ark-ff 0.5.0 MontBackend restated) ---- */
void zko_fe_from_u64(int fid, uint64_t v, uint64_t out[4]);
void zko_fe_from_canonical(int fid, const uint64_t in[4], uint64_t out[4]); /* in < p, plain -> Montgomery */
void zko_fe_to_canonical(int fid, const uint64_t in[4], uint64_t out[4]);   /* Montgomery -> plain       */
void zko_fe_from_le_bytes_mod_order(int fid, const uint8_t *bytes, size_t len, uint64_t out[4]);
void zko_fe_to_bytes_be(int fid, const uint64_t in[4], uint8_t out[32]);
void zko_fe_to_bytes_le(int fid, const uint64_t in[4], uint8_t out[32]);
void zko_fe_add(int fid, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);
void zko_fe_sub(int fid, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);
void zko_fe_mul(int fid, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);
void zko_fe_inv(int fid, const uint64_t a[4], uint64_t out[4]);
void zko_fe_sum(int fid, const uint64_t *v, uint64_t n, uint64_t out[4]);

/* ---- Keccak-256 + transcript (transcripts/src/fiat_shamir/fiat_shamir_transcript.rs) ---- */
void zko_keccak256(const uint8_t *data, size_t len, uint8_t out[32]);
typedef struct zko_transcript zko_transcript;
zko_transcript *zko_transcript_new(void);
zko_transcript *zko_transcript_clone(const zko_transcript *t);
void zko_transcript_free(zko_transcript *t);
void zko_transcript_append(zko_transcript *t, const uint8_t *data, size_t len);
void zko_transcript_sample(zko_transcript *t, uint8_t out[32]);
void zko_transcript_challenge(zko_transcript *t, int fid, uint64_t out[4]);

/* ---- MultilinearPolynomial (polynomials/src/multilinear/evaluation_form.rs) ---- */
int  zko_mle_partial_evaluate(int fid, const uint64_t *in, uint64_t len, uint32_t var, const uint64_t r[4], uint64_t *out);
int  zko_mle_evaluate(int fid, const uint64_t *in, uint64_t len, const uint64_t *rs, uint32_t nr, uint64_t out[4]);
void zko_mle_to_bytes(int fid, const uint64_t *in, uint64_t len, uint8_t *out /* 32*len */);
void zko_mle_scalar_mul(int fid, const uint64_t *in, uint64_t len, const uint64_t s[4], uint64_t *out);
void zko_mle_tensor_add(int fid, const uint64_t *wb, const uint64_t *wc, uint64_t len, uint64_t *out /* len*len */);
void zko_mle_tensor_mul(int fid, const uint64_t *wb, const uint64_t *wc, uint64_t len, uint64_t *out /* len*len */);
void zko_mle_add(int fid, const uint64_t *a, const uint64_t *b, uint64_t len, uint64_t *out);

/* ---- Product/SumPolynomial element-wise reduce (composed/{product,sum}_polynomial.rs) ----
 * tables: P*D tables of `len` elements, table (p,d) at tables + (p*D+d)*len*4.            */
void zko_sumpoly_reduce(int fid, const uint64_t *tables, uint32_t P, uint32_t D, uint64_t len, uint64_t *out /* len */);

/* ---- DenseUnivariatePolynomial (polynomials/src/univariate/dense_univariate.rs) ---- */
void zko_univariate_evaluate(int fid, const uint64_t *coeffs, uint32_t n, const uint64_t x[4], uint64_t out[4]);
void zko_lagrange_interpolate(int fid, const uint64_t *xs, const uint64_t *ys, uint32_t n, uint64_t *out_coeffs /* n */);

/* ---- basic (plain) sumcheck (sumcheck_protocol/src/basic_sumcheck/{prover,verifier}.rs) ---- */
void zko_split_and_sum(int fid, const uint64_t *in, uint64_t len, uint64_t out[8]);
/* round_polys: n x 2 elements; challenges (extra, not in the reference proof): n elements; final_eval: 1 element */
int  zko_basic_prove(int fid, const uint64_t *table, uint64_t len, uint64_t claimed_sum[4],
                     uint64_t *round_polys, uint64_t *challenges, uint64_t final_eval[4]);
int  zko_basic_verify(int fid, const uint64_t *table, uint64_t len, const uint64_t claimed_sum[4],
                      const uint64_t *round_polys, uint32_t n_rounds);

/* ---- product sumcheck (sumcheck_protocol/src/gkr_sumcheck/sumcheck_gkr_protocol.rs) ---- */
void zko_generate_round_univariate(int fid, const uint64_t *tables, uint32_t P, uint32_t D, uint64_t len,
                                   uint64_t *out_evals /* D+1 */);
/* coeffs: n x (D+1); challenges: n.  final_tables (may be NULL): P*D elements left after the last fold. */
int  zko_product_prove(int fid, const uint64_t *tables, uint32_t P, uint32_t D, uint64_t len,
                       const uint64_t claimed_sum[4], zko_transcript *t,
                       uint64_t *coeffs, uint64_t *challenges, uint64_t *final_tables);
int  zko_product_verify(int fid, const uint64_t claimed_sum[4], const uint64_t *coeffs, uint32_t n_rounds, uint32_t D,
                        zko_transcript *t, uint64_t *challenges, uint64_t last_claim[4]);

/* ---- circuit (circuit/src/arithmetic_circuit.rs) ----
 * Layers are given output-first (layer 0 = output layer) as in the reference.  Gates of layer i are
 * entries [layer_off[i], layer_off[i+1]) of left/right/out/op (op: 0 = Add, 1 = Mul).            */
typedef struct {
    uint32_t n_layers;
    const uint64_t *layer_off; /* n_layers + 1 */
    const uint32_t *left, *right, *out;
    const uint8_t *op;
} zko_circuit;
/* sizes[i] = number of values of layer i (i = 0..n_layers; layer n_layers = inputs); values concatenated. */
int  zko_circuit_evaluate(int fid, const zko_circuit *c, const uint64_t *inputs, uint64_t n_inputs,
                          uint64_t *sizes /* n_layers+1 */, uint64_t *values /* caller-sized */, uint64_t values_cap);
uint32_t zko_num_of_layer_variables(uint32_t layer_index);
uint64_t zko_gate_position(uint32_t layer_index, uint64_t a, uint64_t b, uint64_t c);
void zko_add_i_mul_i(int fid, const zko_circuit *c, uint32_t layer, uint64_t *add_i, uint64_t *mul_i);

/* ---- GKR (gkr/src/{gkr_protocol,utils}.rs) ----
 * Proof layout (flat): output (n_out elems, unpadded), claimed_sum, per layer i: sumcheck claimed_sum,
 * 2(i+1) rounds (layer 0: 2 rounds) x 3 coeffs, the rounds' challenges; wb/wc evaluations for layers 0..L-2. */
typedef struct {
    uint64_t *output;      uint64_t n_output;
    uint64_t claimed_sum[4];
    uint64_t *layer_claims;   /* L elements */
    uint64_t *coeffs;         /* sum_i rounds_i * 3 elements */
    uint64_t *challenges;     /* sum_i rounds_i elements */
    uint64_t *wb, *wc;        /* L-1 elements each */
} zko_gkr_proof;
uint64_t zko_gkr_total_rounds(uint32_t n_layers);
int  zko_gkr_prove(int fid, const zko_circuit *c, const uint64_t *inputs, uint64_t n_inputs, zko_gkr_proof *proof);
int  zko_gkr_verify(int fid, const zko_circuit *c, const zko_gkr_proof *proof, const uint64_t *inputs, uint64_t n_inputs);

/* ---- GKR over layers of explicit width, add_i / mul_i evaluated from the gate list (O(gates) per round, no dense
 * 2^(3i+2) tables): gkr_protocol.rs:26-143 / :146-236 for circuits the reference's dense storage cannot hold.
 * layer_bits[li] = log2(#values of layer li), li = 0..n_layers; rounds of layer li = 2 * layer_bits[li + 1];
 * the output claim binds layer_bits[0] successive challenges (one in the reference's shape; layer_bits[0] == 0 is the
 * reference's padded single output).  Identical to zko_gkr_prove on reference-shaped circuits (tests/test_oracle.py).
 * proof->output must hold 2^layer_bits[0] elements.  Gate lists duplicate-free. */
int  zko_gkr_prove_sparse(int fid, uint32_t n_layers, const uint32_t *layer_bits, const uint64_t *layer_off,
                          const uint32_t *left, const uint32_t *right, const uint32_t *out, const uint8_t *op,
                          const uint64_t *inputs, uint64_t n_inputs, zko_gkr_proof *proof);
int  zko_gkr_verify_sparse(int fid, uint32_t n_layers, const uint32_t *layer_bits, const uint64_t *layer_off,
                           const uint32_t *left, const uint32_t *right, const uint32_t *out, const uint8_t *op,
                           const zko_gkr_proof *proof, const uint64_t *inputs, uint64_t n_inputs);

/* ---- the bench workload's seeded tables, as the CUDA generator makes them ---- */
void zko_table_generate(int fid, uint64_t seed, uint64_t table_id, uint64_t n, uint64_t first, uint64_t step, uint64_t *out);

/* ---- multilinear KZG over BLS12-381 G1 (multilinear_kzg/src/{trusted_setup,multilinear_kzg}.rs; zkoracle_kzg.c) ----
 * Points: affine (x, y) as 6 + 6 uint64 little-endian limbs, Montgomery form (R = 2^384), infinity = all zero.
 * Scalars: BLS12-381 Fr elements (4 limbs, Montgomery).  zko_set_threads applies to the sums over the setup. */
void zko_g1_generator(uint64_t out[12]);
void zko_fq_from_canonical(const uint64_t in[6], uint64_t out[6]);
void zko_fq_to_canonical(const uint64_t in[6], uint64_t out[6]);
void zko_fq_op(int op, const uint64_t a[6], const uint64_t b[6], uint64_t out[6]);   /* 0 add, 1 sub, 2 Montgomery mul */
int  zko_g1_is_on_curve(const uint64_t p[12]);
void zko_g1_add(const uint64_t a[12], const uint64_t b[12], uint64_t out[12]);
void zko_g1_neg(const uint64_t a[12], uint64_t out[12]);
void zko_g1_mul(const uint64_t p[12], const uint64_t k_canonical[4], uint64_t out[12]);   /* mul_bigint */
int  zko_kzg_setup_g1(const uint64_t *taus, uint32_t n, uint64_t *g1_out /* 12 * 2^n */);
int  zko_kzg_commit(const uint64_t *vals, uint64_t len, const uint64_t *g1, uint64_t g1_len, uint64_t out[12]);
int  zko_kzg_open(const uint64_t *vals, uint32_t nvars, const uint64_t *g1, uint64_t g1_len, const uint64_t *opening,
                  uint32_t n_opening, uint64_t eval[4], uint64_t *proofs /* 12 * nvars */);
int  zko_kzg_verify_trapdoor(const uint64_t *taus, uint32_t n, const uint64_t commitment[12], const uint64_t *opening,
                             const uint64_t eval[4], const uint64_t *proofs);

#ifdef __cplusplus
}
#endif
#endif
