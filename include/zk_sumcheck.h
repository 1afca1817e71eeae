/*
 * zk_sumcheck.h -- C ABI of the B200 sumcheck / GKR prover library (libzkb200.so).
 *
 * The reference (casweeney/zk-cryptography-research-implementations) is a pure-Rust workspace with
 * no FFI; the drop-in boundary is its public generic API.  Each entry point below names the
 * reference item it replaces (paths relative to the reference root).  A Rust `-sys` shim binds
 * these unchanged: arkworks' `Fp<MontBackend<_,4>,4>` is `[u64;4]` in Montgomery form, which is
 * exactly the element layout used here, so `&[F]` crosses as `*const u64`.
 *
 * Conventions
 *   - every field element in or out: 4 x uint64_t little-endian limbs, Montgomery form, canonical
 *   - status: 0 = ok; ZK_ERR_ASSERT = the reference would have panicked (zk_last_error() returns the
 *     reference's panic text); ZK_ERR_CUDA = CUDA/NCCL failure; ZK_ERR_ARG = bad argument.
 *     Nothing unwinds across the boundary.
 *   - a zk_ctx is bound to one GPU and one stream and is not thread-safe (one per host thread /
 *     one per rank); there is NO CPU fallback: without a usable GPU zk_ctx_create fails.
 */
#ifndef ZK_SUMCHECK_H
#define ZK_SUMCHECK_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ZK_BN254_FQ = 0, ZK_BN254_FR = 1, ZK_BLS12_381_FR = 2 };
enum { ZK_OK = 0, ZK_ERR_ASSERT = -1, ZK_ERR_CUDA = -2, ZK_ERR_ARG = -3 };
/* flags for the one-shot provers */
enum {
    ZK_FLAG_DIRECT_S1 = 1,   /* compute s(1) in the kernel every round instead of claim - s(0) */
    ZK_FLAG_SKIP_ABSORB = 2,   /* MEASUREMENT ONLY.  zk_prove_basic_device / zk_prove_basic: leave the 32*N-byte table absorb of
                                  prover.rs:38-39 out of the (internal) transcript, so that the n rounds can be timed without the
                                  serial host Keccak.  The resulting proof does NOT match the reference or verify; real proofs
                                  pass 0.  (zk_prove_basic_sharded takes the caller's transcript instead, which carries the
                                  absorb.)  zk_gkr_prove_wide*: do not absorb the output layer (same caveat). */
    ZK_FLAG_NO_CLAIM_ABSORB = 8, /* zk_prove_product[_sharded]: do not absorb claimed_sum first (continuation of a sumcheck whose
                                  earlier rounds ran in a previous call -- the two phases of a sparse GKR layer) */
    ZK_FLAG_NCCL_EXCHANGE = 4, /* sharded provers: exchange the per-round partials with ncclAllGather even if the
                                  shared mailboxes are attached (for comparison) */
    ZK_FLAG_HOST_EXCHANGE = 64, /* sharded provers: exchange the per-round partials through the host (shared mailboxes, or
                                  ncclAllGather when they are not attached) even if the ranks' exchange slots are peer-mapped */
    ZK_FLAG_TRUSTED_CLAIM = 32, /* zk_prove_product[_sharded]: the caller guarantees that claimed_sum IS the sum of the polynomial over
                                  the hypercube (the GKR layer prover computed it): round 0 may then derive s(1) = claimed_sum - s(0)
                                  like every later round instead of summing it.  The reference always sums (sumcheck_gkr_protocol.rs:
                                  127-137); with a true claim the proof is identical, with a false one it would differ -- hence opt-in */
    ZK_FLAG_HOST_ROUNDS = 16   /* keep every round on the host-driven path: one kernel + one host Fiat-Shamir step per
                                  round.  Default: once tables x entries <= 2^tail_log (zk_ctx_set_tail_log, default 20)
                                  ONE persistent launch runs all remaining rounds with the transcript on the device
                                  (transcripts/.../fiat_shamir_transcript.rs:12-43 restated in csrc/dev_transcript.cuh,
                                  the round loop in csrc/devrounds.cuh); the proof is identical either way */
};

typedef struct zk_ctx zk_ctx;
typedef struct zk_table zk_table;
typedef struct zk_sumpoly zk_sumpoly;
typedef struct zk_transcript zk_transcript;
typedef struct zk_wide_circuit zk_wide_circuit;

/* ---- library / context ---- */
const char *zk_version(void);
int  zk_ctx_create(zk_ctx **out, int field_id, int device);
/* same, launching on a stream the caller owns (e.g. torch's current stream); stream = cudaStream_t */
int  zk_ctx_create_on_stream(zk_ctx **out, int field_id, int device, void *stream);
void zk_ctx_destroy(zk_ctx *);
const char *zk_last_error(const zk_ctx *);
int  zk_ctx_synchronize(zk_ctx *);
/* kernel accounting for bench.py: launches since the last reset; with profiling on, the summed
 * CUDA-event time and algorithmic bytes of the round kernels (round_evals / fold_evals). */
int  zk_ctx_set_profiling(zk_ctx *, int on);
int  zk_ctx_reset_stats(zk_ctx *);
int  zk_ctx_get_stats(zk_ctx *, uint64_t *launches, uint64_t *round_launches, double *round_ms, double *round_bytes);
/* Device-resident rounds: a sumcheck whose tables together hold at most 2^tail_log entries (one table: 2^tail_log; the
 * three tables of a GKR phase: 2^(tail_log-2) each, rounded down) finishes in ONE persistent cooperative launch that runs
 * every remaining round -- sums, fold, grid barrier, Lagrange coefficients (dense_univariate.rs:74-127), transcript absorb
 * and challenge (fiat_shamir_transcript.rs:22-43) -- on the GPU; blocks leave the loop as the tables shrink, so the last
 * rounds run on one block.  Sharded provers exchange the per-round partial evaluations between the ranks' kernels over
 * peer memory (zk_comm_peer_exchange).  Default 20 (env ZKB200_TAIL_LOG); 0 keeps every round host-driven; at most 32.
 * Proofs are bit-identical for every setting. */
int  zk_ctx_set_tail_log(zk_ctx *, int tail_log);
int  zk_ctx_get_tail_log(const zk_ctx *);

/* ---- field helpers on the host (ark-ff: F::from(u64), into_bigint, from_le_bytes_mod_order) ---- */
int  zk_fe_from_u64(int field_id, uint64_t v, uint64_t out[4]);
int  zk_fe_to_canonical(int field_id, const uint64_t in[4], uint64_t out[4]);
int  zk_fe_from_canonical(int field_id, const uint64_t in[4], uint64_t out[4]);
int  zk_fe_add(int field_id, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);
int  zk_fe_sub(int field_id, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);
int  zk_fe_mul(int field_id, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);

/* DenseUnivariatePolynomial::lagrange_interpolate on x = 0..n_evals-1 (polynomials/src/univariate/dense_univariate.rs:74-98)
 * and ::evaluate (:57-68) */
int  zk_interpolate_evals(int field_id, uint32_t n_evals, const uint64_t *evals, uint64_t *coeffs);
int  zk_univariate_evaluate(int field_id, const uint64_t *coeffs, uint32_t n, const uint64_t x[4], uint64_t out[4]);

/* ---- transcript: transcripts/src/fiat_shamir/fiat_shamir_transcript.rs:5-43 (host, Keccak-256) ---- */
zk_transcript *zk_transcript_new(void);                                              /* Transcript::new :12-16 */
void zk_transcript_free(zk_transcript *);
void zk_transcript_append(zk_transcript *, const uint8_t *data, size_t len);         /* append :22-24 */
void zk_transcript_sample(zk_transcript *, uint8_t out[32]);                         /* sample_random_challenge :29-36 */
void zk_transcript_challenge(zk_transcript *, int field_id, uint64_t out[4]);        /* random_challenge_as_field_element :38-43 */

/* ---- tables: MultilinearPolynomial<F> (polynomials/src/multilinear/evaluation_form.rs:7-18) ---- */
/* new(&[F]) :12-18 -- copies n elements from the host; n must be a power of two
 * ("Evaluated values must be a power of 2") */
int  zk_table_upload(zk_ctx *, const uint64_t *mont_limbs, uint64_t n, zk_table **out);
/* synthetic table generated on the device: local entry j = global entry first + j*step */
int  zk_table_generate(zk_ctx *, uint64_t seed, uint64_t table_id, uint64_t n, uint64_t first, uint64_t step, zk_table **out);
/* refill an existing table in place with the same generator (bench: restore inputs between steps) */
int  zk_table_regenerate(zk_ctx *, zk_table *, uint64_t seed, uint64_t table_id, uint64_t n, uint64_t first, uint64_t step);
/* wrap caller-owned device memory (e.g. a torch tensor) without copying */
int  zk_table_wrap(zk_ctx *, void *device_ptr, uint64_t n, zk_table **out);
int  zk_table_clone(zk_ctx *, const zk_table *, zk_table **out);                      /* Clone */
int  zk_table_download(zk_ctx *, const zk_table *, uint64_t *out_limbs);              /* .evaluated_values */
/* refill an existing table from host memory, asynchronously on the context's stream (end-to-end input copy) */
int  zk_table_upload_into(zk_ctx *, zk_table *, const uint64_t *mont_limbs, uint64_t n);
int  zk_pinned_alloc(size_t bytes, void **out);     /* page-locked host memory for the copies above */
void zk_pinned_free(void *p);
uint64_t zk_table_len(const zk_table *);
void *zk_table_device_ptr(const zk_table *);
void zk_table_free(zk_ctx *, zk_table *);

/* partial_evaluate(&Vec<F>, evaluating_variable, value) :61-106 -- halves the table (in place for var 0) */
int  zk_mle_partial_evaluate(zk_ctx *, zk_table *t, uint32_t var, const uint64_t r[4]);
/* evaluate(&self, &[F]) :21-33 -- non-destructive; n_values may be < log2(len) (returns entry 0) */
int  zk_mle_evaluate(zk_ctx *, const zk_table *t, const uint64_t *values, uint32_t n_values, uint64_t out[4]);
/* convert_to_bytes :35-43 -- 32*len bytes, big-endian canonical, converted on the GPU */
int  zk_mle_to_bytes(zk_ctx *, const zk_table *t, uint8_t *out_host);
/* scalar_mul :49-57, add_polynomials :145-163, polynomial_tensor_add/mul :108-143 -- new tables */
int  zk_mle_scalar_mul(zk_ctx *, const zk_table *t, const uint64_t s[4], zk_table **out);
int  zk_mle_add(zk_ctx *, const zk_table *a, const zk_table *b, zk_table **out);
int  zk_mle_tensor_add(zk_ctx *, const zk_table *wb, const zk_table *wc, zk_table **out);
int  zk_mle_tensor_mul(zk_ctx *, const zk_table *wb, const zk_table *wc, zk_table **out);
/* split_polynomial_and_sum_each (sumcheck_protocol/src/basic_sumcheck/prover.rs:74-89): out = [sum left, sum right] */
int  zk_sum_halves(zk_ctx *, const zk_table *t, uint64_t out[8]);

/* ---- SumPolynomial / ProductPolynomial (polynomials/src/composed/{sum,product}_polynomial.rs) ----
 * tables[p*D + d] is factor d of product p; the sumpoly takes ownership of the tables (on success only), which must be
 * pairwise distinct (folds are in place).  "different number of variables" if the lengths differ.  (P, D) must be one of
 * (1,1) (1,2) (2,2) (3,2) (4,2) (1,3) (2,3). */
int  zk_sumpoly_create(zk_ctx *, zk_table *const *tables, uint32_t P, uint32_t D, zk_sumpoly **out);
void zk_sumpoly_free(zk_ctx *, zk_sumpoly *);
uint64_t zk_sumpoly_len(const zk_sumpoly *);
zk_table *zk_sumpoly_table(const zk_sumpoly *, uint32_t index);
/* add_polynomials_element_wise (sum_polynomial.rs:57-76; product_polynomial.rs:58-73) -> new table */
int  zk_sumpoly_reduce(zk_ctx *, const zk_sumpoly *, zk_table **out);
/* generate_round_univariate (sumcheck_gkr_protocol.rs:113-143): evals[X] = sum_x f(X, x), X = 0..D */
int  zk_sumcheck_round_evals(zk_ctx *, zk_sumpoly *, uint64_t *evals /* 4*(D+1) */);
/* SumPolynomial::partial_evaluate(0, r) (sum_polynomial.rs:40-53) fused with the NEXT round's
 * generate_round_univariate: one pass over the tables.  evals == NULL: fold only. */
int  zk_sumcheck_fold_and_evals(zk_ctx *, zk_sumpoly *, const uint64_t r[4], uint64_t *evals);

/* ---- one-shot provers (host transcript inside, tables stay on the GPU) ---- */
/* sumcheck_gkr_protocol::prove (sumcheck_gkr_protocol.rs:24-67).  Consumes the sumpoly's tables (folded in
 * place down to one entry each).  coeffs: n*(D+1) elements (round polynomials, coefficient form);
 * challenges: n elements; final_values (may be NULL): P*D elements = the tables after the last fold. */
int  zk_prove_product(zk_ctx *, zk_sumpoly *, const uint64_t claimed_sum[4], zk_transcript *,
                      uint64_t *coeffs, uint64_t *challenges, uint64_t *final_values, uint32_t flags);
/* basic_sumcheck Prover::init + Prover::prove (prover.rs:22-71) on a device table (consumed).
 * claimed_sum: out; round_polys: n*2 elements ([sum left, sum right] per round); challenges (extra,
 * may be NULL): n elements; final_value (may be NULL): the table after the last fold. */
int  zk_prove_basic_device(zk_ctx *, zk_table *t, uint64_t claimed_sum[4], uint64_t *round_polys,
                           uint64_t *challenges, uint64_t final_value[4], uint32_t flags);
/* same from a HOST table (uploads it first): the end-to-end call a reference user makes */
int  zk_prove_basic(zk_ctx *, const uint64_t *host_table, uint64_t n, uint64_t claimed_sum[4], uint64_t *round_polys,
                    uint64_t *challenges, uint64_t final_value[4], uint32_t flags);
/* product sumcheck from HOST tables laid out [P][D][n] (uploads, proves, frees) */
int  zk_prove_product_host(zk_ctx *, const uint64_t *host_tables, uint32_t P, uint32_t D, uint64_t n,
                           const uint64_t claimed_sum[4], zk_transcript *, uint64_t *coeffs, uint64_t *challenges,
                           uint64_t *final_values, uint32_t flags);

/* ---- verifiers: same kernels, transcript replay on the host ----
 * gkr_sumcheck::verify (sumcheck_gkr_protocol.rs:69-106); host only.  challenges (may be NULL): n_rounds elements. */
int  zk_verify_product(int field_id, const uint64_t claimed_sum[4], const uint64_t *coeffs, uint32_t n_rounds, uint32_t D,
                       zk_transcript *, uint64_t *challenges, uint64_t last_claimed_sum[4], int *is_proof_valid);
/* basic_sumcheck Verifier::verify (verifier.rs:23-71): the final `initial_polynomial.evaluate(&challenges)` and the table
 * absorb run on the GPU.  *ok = 1 iff the reference would return true. */
int  zk_verify_basic(zk_ctx *, const zk_table *initial_polynomial, const uint64_t claimed_sum[4], const uint64_t *round_polys,
                     uint32_t n_rounds, int *ok);

/* ---- circuit + GKR (circuit/src/arithmetic_circuit.rs, gkr/src/gkr_protocol.rs) ----
 * Layers output-first as in the reference; gates of layer i are entries [layer_off[i], layer_off[i+1]) of
 * left/right/out/op (op 0 = Add, 1 = Mul). */
typedef struct {
    uint32_t n_layers;
    const uint64_t *layer_off;   /* n_layers + 1 */
    const uint32_t *left, *right, *out;
    const uint8_t *op;
} zk_circuit_desc;
/* Circuit::evaluate (arithmetic_circuit.rs:65-109), host: sizes[n_layers+1], values concatenated output-first */
int  zk_circuit_evaluate(int field_id, const zk_circuit_desc *, const uint64_t *inputs, uint64_t n_inputs,
                         uint64_t *sizes, uint64_t *values, uint64_t values_cap);
uint64_t zk_gkr_total_rounds(uint32_t n_layers);   /* sum over layers of 2(i+1) */
/* gkr_protocol::prove (gkr_protocol.rs:26-143) for reference-shaped circuits (layer i: i output bits -- one at
 * layer 0 -- and i+1 bits per input index).  Proof{circuit_output, claimed_sum, sumcheck_proofs, wb/wc_evaluations}
 * flattened: layer_claims[L], coeffs[rounds*3], challenges[rounds], wb[L-1], wc[L-1]. */
int  zk_gkr_prove(zk_ctx *, const zk_circuit_desc *, const uint64_t *inputs, uint64_t n_inputs,
                  uint64_t *output, uint64_t output_cap, uint64_t *n_output, uint64_t claimed_sum[4],
                  uint64_t *layer_claims, uint64_t *coeffs, uint64_t *challenges, uint64_t *wb, uint64_t *wc);

/* gkr_protocol::verify (gkr_protocol.rs:146-236) of a proof laid out as zk_gkr_prove writes it */
int  zk_gkr_verify(zk_ctx *, const zk_circuit_desc *, const uint64_t *output, uint64_t n_output, const uint64_t *layer_claims,
                   const uint64_t *coeffs, const uint64_t *wb, const uint64_t *wc, const uint64_t *inputs, uint64_t n_inputs,
                   int *ok);

/* ---- GKR for wide layers: sparse two-phase layer sumcheck over 2^m-entry tables (m = log2 width of the layer below)
 * instead of the reference's dense 2^(3i+2) / 4^(i+1) tables; identical round polynomials on reference-shaped
 * circuits.  layer_bits[li] = log2(#values of layer li), li = 0..n_layers (last = inputs); gates as in
 * zk_circuit_desc, duplicate-free (checked: ZK_ERR_ARG "duplicate gate"; indices are range-checked first).  The
 * circuit lives on the GPU as three CSR orderings per layer, built there (histogram, scan, scatter): the step the reference
 * performs inside prove (circuit/src/arithmetic_circuit.rs:126-163 via gkr_protocol.rs:58). */
int  zk_wide_circuit_create(zk_ctx *, uint32_t n_layers, const uint32_t *layer_bits, const uint64_t *layer_off,
                            const uint32_t *left, const uint32_t *right, const uint32_t *out, const uint8_t *op,
                            zk_wide_circuit **result);
void zk_wide_circuit_free(zk_ctx *, zk_wide_circuit *);
uint64_t zk_wide_circuit_total_rounds(const zk_wide_circuit *);   /* sum over layers of 2 * layer_bits[li+1] */
/* log2 of the output layer as the prover sees it: layer_bits[0], except that a single output (layer_bits[0] == 0) is the
 * reference's padded [out, 0] layer (gkr_protocol.rs:43-51) and reports 1 -- `output` buffers hold 2^this elements */
uint32_t zk_wide_circuit_output_bits(const zk_wide_circuit *);
/* gkr_protocol::prove (gkr_protocol.rs:26-143).  output (may be NULL): 2^layer_bits[0] elements; the rest as
 * zk_gkr_prove.  The output claim binds layer_bits[0] successive challenges (one in the reference's shape).
 * ZK_FLAG_SKIP_ABSORB: do not absorb the output layer into the transcript. */
int  zk_gkr_prove_wide(zk_ctx *, const zk_wide_circuit *, const uint64_t *inputs, uint64_t n_inputs, uint64_t *output,
                       uint64_t *claimed_sum, uint64_t *layer_claims, uint64_t *coeffs, uint64_t *challenges,
                       uint64_t *wb, uint64_t *wc, uint32_t flags);
/* same with the input layer already resident in HBM (a table of 2^layer_bits[n_layers] entries; it is not modified) */
int  zk_gkr_prove_wide_device(zk_ctx *, const zk_wide_circuit *, const zk_table *inputs, uint64_t *output,
                              uint64_t *claimed_sum, uint64_t *layer_claims, uint64_t *coeffs, uint64_t *challenges,
                              uint64_t *wb, uint64_t *wc, uint32_t flags);

/* gkr_protocol::verify (gkr_protocol.rs:146-236, claim helpers gkr/src/utils.rs:84-135) of a proof laid out as
 * zk_gkr_prove_wide writes it (output: 2^layer_bits[0] elements).  add_i / mul_i at the sumcheck point are evaluated from
 * the gate list on the GPU (eq tables), W(u) / W(v) of the input layer by the evaluate kernels.  `flags` as given to the
 * prover.  *ok = 1 iff the reference's verifier would return true. */
int  zk_gkr_verify_wide(zk_ctx *, const zk_wide_circuit *, const uint64_t *output, const uint64_t *layer_claims,
                        const uint64_t *coeffs, const uint64_t *wb, const uint64_t *wc, const uint64_t *inputs,
                        uint64_t n_inputs, uint32_t flags, int *ok);
int  zk_gkr_verify_wide_device(zk_ctx *, const zk_wide_circuit *, const uint64_t *output, const uint64_t *layer_claims,
                               const uint64_t *coeffs, const uint64_t *wb, const uint64_t *wc, const zk_table *inputs,
                               uint32_t flags, int *ok);
/* verify_succinct (gkr/src/succinct_gkr_protocol.rs:172-283) without its two MultilinearKZG::verify calls: the output claim,
 * every layer's sumcheck and, for all layers but the input layer, the claim check of zk_gkr_verify_wide.  last_challenges
 * (may be NULL) receives the input layer's 2 * layer_bits[n_layers] challenges: (rb, rc), the points the committed input
 * polynomial must open at (:262-275; zk_kzg_verify).  input_evals == NULL is the reference's behaviour (the opened values
 * are not compared with the last sumcheck claim); with input_evals = {W(rb), W(rc)} as opened, that check is made too. */
int  zk_gkr_verify_wide_succinct(zk_ctx *, const zk_wide_circuit *, const uint64_t *output, const uint64_t *layer_claims,
                                 const uint64_t *coeffs, const uint64_t *wb, const uint64_t *wc, const uint64_t *input_evals,
                                 uint32_t flags, uint64_t *last_challenges, int *ok);

/* ---- one process per GPU: tables sharded on the LOW index bits (rank q holds entries q, q+G, q+2G, ...) ----
 * NCCL over NVLink/NVSwitch carries one all-gather of (D+1) elements per round; folds stay local.
 * Rank 0 calls zk_comm_unique_id and ships the 128 bytes to the other ranks (torch.distributed, MPI, ...). */
int  zk_comm_unique_id(uint8_t out[128]);
int  zk_comm_init(zk_ctx *, int rank, int world, const uint8_t id[128]);   /* world: power of two */
/* optional, same node only: round mailboxes in a POSIX shared-memory segment (Mailbox[2][world]); every rank's round
 * kernel then publishes its partial evaluations straight into host memory all rank processes poll -- no collective
 * on the per-round path.  Rank 0: create = 1 before the others attach; unlink after all have attached. */
int  zk_comm_attach_mailboxes(zk_ctx *, const char *shm_name, int create);
int  zk_comm_unlink_mailboxes(const char *shm_name);
/* 1 if zk_comm_init could map every rank's exchange slots into every other rank (cudaIpc peer access): the sharded provers
 * then run their sharded rounds in ONE persistent launch per rank and the per-round (d+1) x 32 B partial evaluations go from
 * kernel to kernel over NVLink; 0: host-mediated exchange (mailboxes / ncclAllGather).  ZKB200_PEER_EXCHANGE=0 forces 0. */
int  zk_comm_peer_exchange(const zk_ctx *);
int  zk_comm_destroy(zk_ctx *);
int  zk_comm_rank(const zk_ctx *);
int  zk_comm_world(const zk_ctx *);
/* sumcheck_gkr_protocol::prove with `sp` holding this rank's shard; same outputs on every rank.  When the
 * local tables are down to `collapse_len` entries they are gathered and the rest runs unsharded. */
int  zk_prove_product_sharded(zk_ctx *, zk_sumpoly *sp, const uint64_t claimed_sum[4], zk_transcript *,
                              uint64_t *coeffs, uint64_t *challenges, uint64_t *final_values, uint32_t flags,
                              uint64_t collapse_len);
/* basic_sumcheck Prover::prove over a sharded table (any world size).  The whole-table absorb of prover.rs:38-39 is the
 * caller's: every rank passes an identical transcript that has already absorbed it (or deliberately has not). */
int  zk_prove_basic_sharded(zk_ctx *, zk_table *local, zk_transcript *, uint64_t claimed_sum[4], uint64_t *round_polys,
                            uint64_t *challenges, uint64_t final_value[4], uint32_t flags, uint64_t collapse_len);
/* gkr_protocol::prove (gkr_protocol.rs:26-143) with every layer's phase tables and sumchecks spread over the ranks.
 * The circuit object, the input layer (`inputs`: ALL 2^layer_bits[n_layers] values) and the transcript are replicated on every
 * rank; rank q builds and proves the wires whose low index bits are q; every rank returns the same proof, equal to
 * zk_gkr_prove_wide's.  Layers with fewer than two wires per rank run unsharded on every rank.  Outputs as zk_gkr_prove_wide. */
int  zk_gkr_prove_wide_sharded(zk_ctx *, const zk_wide_circuit *, const zk_table *inputs, uint64_t *output,
                               uint64_t *claimed_sum, uint64_t *layer_claims, uint64_t *coeffs, uint64_t *challenges,
                               uint64_t *wb, uint64_t *wc, uint32_t flags, uint64_t collapse_len);
/* MultilinearPolynomial::evaluate over a sharded table (values: all log2(global length) challenges) */
int  zk_mle_evaluate_sharded(zk_ctx *, const zk_table *local, const uint64_t *values, uint32_t n_values, uint64_t out[4]);

/* ---- multilinear KZG over BLS12-381 G1: the input commitment of succinct GKR ----
 * Replaces multilinear_kzg/src/trusted_setup.rs:12-63 (TrustedSetup::initialize_setup, G1 side),
 * multilinear_kzg/src/multilinear_kzg.rs:25-46 (commit_to_polynomial) and :51-127 (open_and_prove), the calls
 * gkr/src/succinct_gkr_protocol.rs:43-44 and :151-154 make.  The context must be a BLS12-381 Fr one (the curve's scalar
 * field).  G1 points cross the boundary in affine form: x then y, 6 + 6 uint64 little-endian limbs, Montgomery form with
 * R = 2^384 -- arkworks' `G1Affine` coordinates byte for byte (`P::G1` values via `into_affine()`); the point at infinity
 * is all zero.  Scalars are Fr elements in Montgomery form like every other table.
 * Every sum over the setup is one bucket-method multi-scalar multiplication on the GPU; the "blown up" quotients of
 * open_and_prove (multilinear_kzg.rs:93-108, :183-214) are summed against the setup folded over its leading variables
 * (built once per setup), so an opening costs 2^n point additions instead of n 2^n scalar multiplications. */
typedef struct zk_kzg_setup zk_kzg_setup;
/* initialize_setup(taus): g1_powers_of_tau[i] = lagrange_basis(taus)[i] * G, computed on the GPU.
 * ZK_ERR_ASSERT "requires at least one variable" (trusted_setup.rs:28). */
int  zk_kzg_setup_create(zk_ctx *, const uint64_t *taus /* 4*n */, uint32_t n, zk_kzg_setup **out);
/* an existing g1_powers_of_tau (2^n points, checked to be on the curve: ZK_ERR_ARG otherwise) */
int  zk_kzg_setup_from_points(zk_ctx *, const uint64_t *g1_points /* 12 * 2^n */, uint32_t n, zk_kzg_setup **out);
void zk_kzg_setup_free(zk_ctx *, zk_kzg_setup *);
uint32_t zk_kzg_setup_num_vars(const zk_kzg_setup *);
/* level 0: g1_powers_of_tau (2^n points); level k: the setup summed over its first k variables (2^(n-k) points) */
int  zk_kzg_setup_points(zk_ctx *, const zk_kzg_setup *, uint32_t level, uint64_t *out /* 12 * 2^(n-level) */);
/* commit_to_polynomial.  ZK_ERR_ASSERT "Polynomial evaluation must match g1 length". */
int  zk_kzg_commit(zk_ctx *, zk_kzg_setup *, const uint64_t *evaluated_values, uint64_t len, uint64_t commitment[12]);
int  zk_kzg_commit_device(zk_ctx *, zk_kzg_setup *, const zk_table *, uint64_t commitment[12]);   /* the table is not modified */
/* open_and_prove -> MultilinearKZGProof{evaluation, proofs[n]}.  ZK_ERR_ASSERT "number of polynomial variables must match
 * length of opening values" / "Opening values must match number of variables from trusted setup". */
int  zk_kzg_open(zk_ctx *, zk_kzg_setup *, const uint64_t *evaluated_values, uint64_t len, const uint64_t *opening_values,
                 uint32_t n_opening, uint64_t evaluation[4], uint64_t *proofs /* 12 * n */);
int  zk_kzg_open_device(zk_ctx *, zk_kzg_setup *, const zk_table *, const uint64_t *opening_values, uint32_t n_opening,
                        uint64_t evaluation[4], uint64_t *proofs);
/* Several GPUs (one process per GPU, after zk_comm_init): the setup and the polynomial are replicated on every rank (like the
 * circuit and the input layer of zk_gkr_prove_wide_sharded); rank q sums the q-th contiguous share of the points of every
 * large sum, the partial results (96 bytes each) are all-gathered once per call and added in rank order: every rank returns
 * the same points, equal to the unsharded calls'. */
int  zk_kzg_commit_sharded(zk_ctx *, zk_kzg_setup *, const zk_table *, uint64_t commitment[12]);
int  zk_kzg_open_sharded(zk_ctx *, zk_kzg_setup *, const zk_table *, const uint64_t *opening_values, uint32_t n_opening,
                         uint64_t evaluation[4], uint64_t *proofs);
/* The verifier's half, host only (no context, no GPU; the prover never calls these).  G2 points: affine x.c0, x.c1, y.c0,
 * y.c1 (4 x 6 limbs, Montgomery; `G2Affine`'s coordinates), infinity all zero.
 * compute_g2_powers_of_tau (trusted_setup.rs:65-78): out[i] = taus[i] * G2.  ZK_ERR_ASSERT for n == 0. */
int  zk_kzg_g2_powers_of_tau(const uint64_t *taus, uint32_t n, uint64_t *out /* 24 * n */);
/* MultilinearKZG::verify (multilinear_kzg.rs:132-159): e(C - v G1, G2) == prod_i e(Q_i, tau_i G2 - r_i G2), one product of
 * Miller loops and one final exponentiation.  ZK_ERR_ASSERT "Number of opening values must match number of proofs". */
int  zk_kzg_verify(const uint64_t *g2_powers_of_tau, uint32_t n_g2, const uint64_t commitment[12],
                   const uint64_t *opening_values, uint32_t n_opening, const uint64_t evaluation[4], const uint64_t *proofs,
                   uint32_t n_proofs, int *ok);
/* prod_i e(g1_points[i], g2_points[i]) == 1 -- the check zk_kzg_verify is built on (one final exponentiation for all pairs) */
int  zk_pairing_product_is_one(const uint64_t *g1_points /* 12*n */, const uint64_t *g2_points /* 24*n */, uint32_t n, int *ok);
int  zk_g1_is_on_curve(const uint64_t point[12]);
void zk_g1_generator(uint64_t out[12]);
void zk_g2_generator(uint64_t out[24]);
/* sum_i scalars[i] * points[i] over caller-supplied host arrays (`.map(mul_bigint).sum()`); any n, points checked */
int  zk_g1_msm(zk_ctx *, const uint64_t *scalars /* 4*n */, const uint64_t *points /* 12*n */, uint64_t n, uint64_t out[12]);

/* ---- measurement: register-resident field arithmetic, no memory traffic (the IMAD-pipe ceiling) ----
 * kind 0: Montgomery product, 1: fold by a per-round scalar, 2: unreduced multiply-accumulate. */
int  zk_arith_probe(zk_ctx *, int kind, uint32_t iters, int blocks_per_sm, double *ops_per_s, double *ms);

/* the same for the curve arithmetic of the multi-scalar multiplication: kind 0 = 381-bit Montgomery products, kind 1 = mixed
 * point additions (XYZZ += affine), both register-resident; operations per second */
int  zk_g1_arith_probe(zk_ctx *, int kind, uint32_t iters, int blocks_per_sm, double *ops_per_s, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* ZK_SUMCHECK_H */
