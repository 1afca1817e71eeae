/*
 * zkoracle.c -- CPU restatement of the reference's sumcheck / GKR prover path.
 * TEST INFRASTRUCTURE ONLY (see zkoracle.h for the rules and the parity status).
 *
 * Each function cites the reference file:line it follows (paths relative to the
 * reference root).  The pass structure of the reference is kept on purpose (fresh
 * vectors per fold, (d+1) full folds per round, ...) because this file is also the
 * timed single-thread CPU baseline of bench.py.
 */
#include "zkoracle.h"
#include "field_consts.h"
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;

/* ------------------------------------------------------------------------------------------
 * Field: ark-ff 0.5.0 `Fp<MontBackend<Config,4>,4>` -- 4x64 limbs, Montgomery, R = 2^256.
 * Call sites in the reference: evaluation_form.rs:1,39,89; prover.rs:1,28,82-83,92;
 * sumcheck_gkr_protocol.rs:1,47,128,148,153; fiat_shamir_transcript.rs:42.
 * ------------------------------------------------------------------------------------------ */
static inline const uint64_t *P_(int fid) { return ZKF_P_64[fid]; }

static inline int ge_p(const uint64_t a[4], const uint64_t p[4]) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > p[i]) return 1;
        if (a[i] < p[i]) return 0;
    }
    return 1;
}
static inline uint64_t sub4(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    u128 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
    return (uint64_t)borrow;
}
static inline uint64_t add4(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    u128 c = 0;
    for (int i = 0; i < 4; ++i) {
        c += (u128)a[i] + b[i];
        r[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
static inline void f_add(int fid, fe *r, const fe *a, const fe *b) {
    uint64_t t[4];
    uint64_t carry = add4(t, a->l, b->l);
    if (carry || ge_p(t, P_(fid))) sub4(t, t, P_(fid));
    memcpy(r->l, t, 32);
}
static inline void f_sub(int fid, fe *r, const fe *a, const fe *b) {
    uint64_t t[4];
    if (sub4(t, a->l, b->l)) add4(t, t, P_(fid));
    memcpy(r->l, t, 32);
}
static inline void f_neg(int fid, fe *r, const fe *a) {
    fe z = {{0, 0, 0, 0}};
    f_sub(fid, r, &z, a);
}
/* CIOS Montgomery multiplication (the algorithm MontBackend::mul_assign implements). */
static inline void f_mul(int fid, fe *r, const fe *a, const fe *b) {
    const uint64_t *p = P_(fid);
    const uint64_t inv = ZKF_INV64[fid];
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)t[j] + (u128)a->l[j] * b->l[i];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * inv;
        c = (u128)t[0] + (u128)m * p[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)t[j] + (u128)m * p[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || ge_p(t, p)) sub4(t, t, p);
    memcpy(r->l, t, 32);
}
static inline int f_eq(const fe *a, const fe *b) { return memcmp(a->l, b->l, 32) == 0; }
static inline fe f_zero(void) { fe z = {{0, 0, 0, 0}}; return z; }
static inline fe f_one(int fid) { fe o; memcpy(o.l, ZKF_R_64[fid], 32); return o; }
static inline void f_from_canonical(int fid, fe *r, const uint64_t in[4]) {
    fe a, r2;
    memcpy(a.l, in, 32);
    memcpy(r2.l, ZKF_R2_64[fid], 32);
    f_mul(fid, r, &a, &r2);
}
static inline void f_to_canonical(int fid, uint64_t out[4], const fe *a) {
    fe one = {{1, 0, 0, 0}}, r;
    f_mul(fid, &r, a, &one);
    memcpy(out, r.l, 32);
}
static inline fe f_from_u64(int fid, uint64_t v) {
    uint64_t c[4] = {v, 0, 0, 0};
    fe r;
    f_from_canonical(fid, &r, c); /* v < 2^64 < p */
    return r;
}
static void f_pow(int fid, fe *r, const fe *a, const uint64_t e[4]) {
    fe acc = f_one(fid);
    for (int i = 255; i >= 0; --i) {
        f_mul(fid, &acc, &acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) f_mul(fid, &acc, &acc, a);
    }
    *r = acc;
}
static void f_inv(int fid, fe *r, const fe *a) { /* a^(p-2); a != 0 */
    uint64_t e[4], two[4] = {2, 0, 0, 0};
    sub4(e, P_(fid), two);
    f_pow(fid, r, a, e);
}

void zko_fe_from_u64(int fid, uint64_t v, uint64_t out[4]) { fe r = f_from_u64(fid, v); memcpy(out, r.l, 32); }
void zko_fe_from_canonical(int fid, const uint64_t in[4], uint64_t out[4]) { fe r; f_from_canonical(fid, &r, in); memcpy(out, r.l, 32); }
void zko_fe_to_canonical(int fid, const uint64_t in[4], uint64_t out[4]) { fe a; memcpy(a.l, in, 32); f_to_canonical(fid, out, &a); }
void zko_fe_add(int fid, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) { fe r; f_add(fid, &r, (const fe *)a, (const fe *)b); memcpy(out, r.l, 32); }
void zko_fe_sub(int fid, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) { fe r; f_sub(fid, &r, (const fe *)a, (const fe *)b); memcpy(out, r.l, 32); }
void zko_fe_mul(int fid, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) { fe r; f_mul(fid, &r, (const fe *)a, (const fe *)b); memcpy(out, r.l, 32); }
void zko_fe_inv(int fid, const uint64_t a[4], uint64_t out[4]) { fe r; f_inv(fid, &r, (const fe *)a); memcpy(out, r.l, 32); }
/* ---- optional all-core mode (NOT the reference's behaviour: the reference is single-threaded, no rayon) ----
 * zko_set_threads(n > 1) lets the data-parallel loops of the prover path (fold, element-wise reduce, sums, byte
 * conversion) run on n OpenMP threads; the results are the same field elements.  bench.py reports it as a separate,
 * labelled "stronger than the reference" baseline.  Default 1: every loop runs exactly as restated from the reference. */
static int g_threads = 1;
void zko_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int zko_get_threads(void) { return g_threads; }
int zko_openmp_enabled(void) {
#ifdef _OPENMP
    return 1;
#else
    return 0;
#endif
}

void zko_fe_sum(int fid, const uint64_t *v, uint64_t n, uint64_t out[4]) {
    fe acc = f_zero();
    if (g_threads > 1 && n >= 4096) {
        enum { MAXT = 256 };
        int nt = g_threads > MAXT ? MAXT : g_threads;
        fe part[MAXT];
        for (int t = 0; t < nt; ++t) part[t] = f_zero();
#pragma omp parallel for num_threads(nt) schedule(static)
        for (int t = 0; t < nt; ++t) {
            uint64_t lo = n * (uint64_t)t / nt, hi = n * (uint64_t)(t + 1) / nt;
            fe a = f_zero();
            for (uint64_t i = lo; i < hi; ++i) f_add(fid, &a, &a, (const fe *)(v + 4 * i));
            part[t] = a;
        }
        for (int t = 0; t < nt; ++t) f_add(fid, &acc, &acc, &part[t]);
    } else {
        for (uint64_t i = 0; i < n; ++i) f_add(fid, &acc, &acc, (const fe *)(v + 4 * i));
    }
    memcpy(out, acc.l, 32);
}

/* ark-ff `PrimeField::from_le_bytes_mod_order`: the little-endian integer reduced mod p
 * (fiat_shamir_transcript.rs:42).  Horner over the bytes from the most significant end. */
void zko_fe_from_le_bytes_mod_order(int fid, const uint8_t *bytes, size_t len, uint64_t out[4]) {
    fe acc = f_zero();
    fe c256 = f_from_u64(fid, 256);
    for (size_t i = len; i-- > 0;) {
        fe b = f_from_u64(fid, bytes[i]);
        f_mul(fid, &acc, &acc, &c256);
        f_add(fid, &acc, &acc, &b);
    }
    memcpy(out, acc.l, 32);
}
/* `into_bigint().to_bytes_be()` (evaluation_form.rs:39, prover.rs:91-93, sumcheck_gkr_protocol.rs:152-154) */
void zko_fe_to_bytes_be(int fid, const uint64_t in[4], uint8_t out[32]) {
    uint64_t c[4];
    zko_fe_to_canonical(fid, in, c);
    for (int i = 0; i < 32; ++i) out[i] = (uint8_t)(c[3 - i / 8] >> (56 - 8 * (i % 8)));
}
/* `into_bigint().to_bytes_le()` (sumcheck_gkr_protocol.rs:145-150) */
void zko_fe_to_bytes_le(int fid, const uint64_t in[4], uint8_t out[32]) {
    uint64_t c[4];
    zko_fe_to_canonical(fid, in, c);
    for (int i = 0; i < 32; ++i) out[i] = (uint8_t)(c[i / 8] >> (8 * (i % 8)));
}

/* ------------------------------------------------------------------------------------------
 * Keccak-256 (sha3 0.10.8 `Keccak256`: Keccak[c=512], rate 136, pad 0x01..0x80).
 * ------------------------------------------------------------------------------------------ */
static const uint64_t KRC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
    0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
    0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
    0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
    0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
static const int KROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
static inline uint64_t rol64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
static void keccak_f(uint64_t s[25]) {
    for (int rnd = 0; rnd < 24; ++rnd) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; ++x) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
        for (int x = 0; x < 5; ++x) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; ++i) s[i] ^= d[i % 5];
        for (int x = 0; x < 5; ++x)
            for (int y = 0; y < 5; ++y) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(s[x + 5 * y], KROT[x + 5 * y]);
        for (int y = 0; y < 5; ++y)
            for (int x = 0; x < 5; ++x) s[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        s[0] ^= KRC[rnd];
    }
}
#define KRATE 136
struct zko_transcript {
    uint64_t s[25];
    uint8_t buf[KRATE];
    size_t pos;
};
static void sponge_init(struct zko_transcript *t) { memset(t, 0, sizeof *t); }
static void sponge_block(struct zko_transcript *t, const uint8_t *blk) {
    for (int i = 0; i < KRATE / 8; ++i) {
        uint64_t w;
        memcpy(&w, blk + 8 * i, 8); /* little-endian host assumed (x86-64) */
        t->s[i] ^= w;
    }
    keccak_f(t->s);
}
static void sponge_update(struct zko_transcript *t, const uint8_t *d, size_t len) {
    if (t->pos) {
        size_t take = KRATE - t->pos;
        if (take > len) take = len;
        memcpy(t->buf + t->pos, d, take);
        t->pos += take; d += take; len -= take;
        if (t->pos == KRATE) { sponge_block(t, t->buf); t->pos = 0; }
    }
    while (len >= KRATE) { sponge_block(t, d); d += KRATE; len -= KRATE; }
    if (len) { memcpy(t->buf, d, len); t->pos = len; }
}
static void sponge_final(const struct zko_transcript *t0, uint8_t out[32]) {
    struct zko_transcript t = *t0; /* finalize a clone: fiat_shamir_transcript.rs:31 */
    memset(t.buf + t.pos, 0, KRATE - t.pos);
    t.buf[t.pos] ^= 0x01;
    t.buf[KRATE - 1] ^= 0x80;
    sponge_block(&t, t.buf);
    memcpy(out, t.s, 32);
}
void zko_keccak256(const uint8_t *data, size_t len, uint8_t out[32]) {
    struct zko_transcript t;
    sponge_init(&t);
    sponge_update(&t, data, len);
    sponge_final(&t, out);
}
/* Transcript::new  -- fiat_shamir_transcript.rs:12-16 */
zko_transcript *zko_transcript_new(void) {
    zko_transcript *t = (zko_transcript *)malloc(sizeof *t);
    sponge_init(t);
    return t;
}
zko_transcript *zko_transcript_clone(const zko_transcript *t) {
    zko_transcript *c = (zko_transcript *)malloc(sizeof *c);
    *c = *t;
    return c;
}
void zko_transcript_free(zko_transcript *t) { free(t); }
/* Transcript::append -- fiat_shamir_transcript.rs:22-24 */
void zko_transcript_append(zko_transcript *t, const uint8_t *data, size_t len) { sponge_update(t, data, len); }
/* Transcript::sample_random_challenge -- fiat_shamir_transcript.rs:29-36:
 * digest = clone().finalize(); the live hasher then absorbs the digest (state is never reset). */
void zko_transcript_sample(zko_transcript *t, uint8_t out[32]) {
    sponge_final(t, out);
    sponge_update(t, out, 32);
}
/* Transcript::random_challenge_as_field_element -- fiat_shamir_transcript.rs:38-43 */
void zko_transcript_challenge(zko_transcript *t, int fid, uint64_t out[4]) {
    uint8_t d[32];
    zko_transcript_sample(t, d);
    zko_fe_from_le_bytes_mod_order(fid, d, 32, out);
}

/* ------------------------------------------------------------------------------------------
 * MultilinearPolynomial -- polynomials/src/multilinear/evaluation_form.rs
 * ------------------------------------------------------------------------------------------ */
static int is_pow2(uint64_t n) { return n && !(n & (n - 1)); }
static uint32_t ilog2_u64(uint64_t n) { uint32_t k = 0; while (n >>= 1) ++k; return k; }

/* partial_evaluate -- evaluation_form.rs:61-106.  Returns 0, or -1 if the result would not be a
 * power of two (the `new` assertion at :13, "Evaluated values must be a power of 2"). */
int zko_mle_partial_evaluate(int fid, const uint64_t *in_, uint64_t len, uint32_t var, const uint64_t r_[4], uint64_t *out_) {
    const fe *in = (const fe *)in_;
    fe *out = (fe *)out_;
    const fe *r = (const fe *)r_;
    uint64_t half = len / 2;
    if (!is_pow2(half)) return -1;
    uint32_t nvars = ilog2_u64(len);
    uint32_t power = nvars - 1 - var;           /* :82 */
    if (g_threads > 1 && half >= 4096) {        /* all-core mode: same pairs, index computed instead of walked */
        const uint64_t low_mask = (1ull << power) - 1;
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (uint64_t i = 0; i < half; ++i) {
            uint64_t j = ((i & ~low_mask) << 1) | (i & low_mask);
            fe d, m;
            f_sub(fid, &d, &in[j | (1ull << power)], &in[j]);
            f_mul(fid, &m, r, &d);
            f_add(fid, &out[i], &in[j], &m);
        }
        return 0;
    }
    uint64_t i = 0, j = 0;
    while (i < half) {                          /* :69 */
        const fe *y1 = &in[j];                  /* :70 */
        const fe *y2 = &in[j | (1ull << power)];/* :84 */
        fe d, m;
        f_sub(fid, &d, y2, y1);
        f_mul(fid, &m, r, &d);
        f_add(fid, &out[i], y1, &m);            /* :90-91  y1 + r*(y2 - y1) */
        ++i;
        j = ((j + 1) % (1ull << power) == 0) ? j + 1 + (1ull << power) : j + 1; /* :98-102 */
    }
    return 0;
}
/* evaluate -- evaluation_form.rs:21-33: clone, then nr folds of variable 0, return entry 0. */
int zko_mle_evaluate(int fid, const uint64_t *in, uint64_t len, const uint64_t *rs, uint32_t nr, uint64_t out[4]) {
    if (!is_pow2(len)) return -1;
    uint64_t *cur = (uint64_t *)malloc(len * 32);
    memcpy(cur, in, len * 32);
    uint64_t n = len;
    for (uint32_t i = 0; i < nr; ++i) {
        uint64_t *nxt = (uint64_t *)malloc((n / 2 ? n / 2 : 1) * 32);
        if (zko_mle_partial_evaluate(fid, cur, n, 0, rs + 4 * i, nxt)) { free(cur); free(nxt); return -1; }
        free(cur);
        cur = nxt;
        n /= 2;
    }
    memcpy(out, cur, 32);
    free(cur);
    return 0;
}
/* convert_to_bytes -- evaluation_form.rs:35-43 */
void zko_mle_to_bytes(int fid, const uint64_t *in, uint64_t len, uint8_t *out) {
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1 && len >= 4096)
    for (uint64_t i = 0; i < len; ++i) zko_fe_to_bytes_be(fid, in + 4 * i, out + 32 * i);
}
/* scalar_mul -- evaluation_form.rs:49-57 */
void zko_mle_scalar_mul(int fid, const uint64_t *in, uint64_t len, const uint64_t s[4], uint64_t *out) {
    for (uint64_t i = 0; i < len; ++i) f_mul(fid, (fe *)(out + 4 * i), (const fe *)(in + 4 * i), (const fe *)s);
}
/* polynomial_tensor_add -- evaluation_form.rs:108-123 (b outer, c inner) */
void zko_mle_tensor_add(int fid, const uint64_t *wb, const uint64_t *wc, uint64_t len, uint64_t *out) {
    for (uint64_t b = 0; b < len; ++b)
        for (uint64_t c = 0; c < len; ++c)
            f_add(fid, (fe *)(out + 4 * (b * len + c)), (const fe *)(wb + 4 * b), (const fe *)(wc + 4 * c));
}
/* polynomial_tensor_mul -- evaluation_form.rs:125-143 */
void zko_mle_tensor_mul(int fid, const uint64_t *wb, const uint64_t *wc, uint64_t len, uint64_t *out) {
    for (uint64_t b = 0; b < len; ++b)
        for (uint64_t c = 0; c < len; ++c)
            f_mul(fid, (fe *)(out + 4 * (b * len + c)), (const fe *)(wb + 4 * b), (const fe *)(wc + 4 * c));
}
/* add_polynomials -- evaluation_form.rs:145-163 */
void zko_mle_add(int fid, const uint64_t *a, const uint64_t *b, uint64_t len, uint64_t *out) {
    for (uint64_t i = 0; i < len; ++i) f_add(fid, (fe *)(out + 4 * i), (const fe *)(a + 4 * i), (const fe *)(b + 4 * i));
}

/* SumPolynomial::add_polynomials_element_wise over ProductPolynomial::multiply_polynomials_element_wise
 * -- sum_polynomial.rs:57-76, product_polynomial.rs:58-73. */
void zko_sumpoly_reduce(int fid, const uint64_t *tables, uint32_t P, uint32_t D, uint64_t len, uint64_t *out_) {
    fe *out = (fe *)out_;
    fe *prod = (fe *)malloc(len * 32);
    for (uint32_t p = 0; p < P; ++p) {
        const fe *t0 = (const fe *)(tables + (uint64_t)(p * D) * len * 4);
        memcpy(prod, t0, len * 32);                                  /* product_polynomial.rs:64 */
        for (uint32_t d = 1; d < D; ++d) {
            const fe *td = (const fe *)(tables + (uint64_t)(p * D + d) * len * 4);
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1 && len >= 4096)
            for (uint64_t i = 0; i < len; ++i) f_mul(fid, &prod[i], &prod[i], &td[i]); /* :66-70 */
        }
        if (p == 0) memcpy(out, prod, len * 32);                     /* sum_polynomial.rs:63-65 */
        else {
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1 && len >= 4096)
            for (uint64_t i = 0; i < len; ++i) f_add(fid, &out[i], &out[i], &prod[i]); /* :67-73 */
        }
    }
    free(prod);
}

/* ------------------------------------------------------------------------------------------
 * DenseUnivariatePolynomial -- polynomials/src/univariate/dense_univariate.rs
 * ------------------------------------------------------------------------------------------ */
/* evaluate -- dense_univariate.rs:57-68 */
void zko_univariate_evaluate(int fid, const uint64_t *coeffs, uint32_t n, const uint64_t x[4], uint64_t out[4]) {
    fe result = f_zero(), power = f_one(fid), t;
    for (uint32_t i = 0; i < n; ++i) {
        f_mul(fid, &t, (const fe *)(coeffs + 4 * i), &power);
        f_add(fid, &result, &result, &t);
        f_mul(fid, &power, &power, (const fe *)x);
    }
    memcpy(out, result.l, 32);
}
/* lagrange_interpolate / lagrange_basis -- dense_univariate.rs:74-98,101-127.
 * numerator = prod_{x != focus} (X - x); scalar = y / numerator(focus); sum the scaled bases. */
void zko_lagrange_interpolate(int fid, const uint64_t *xs_, const uint64_t *ys_, uint32_t n, uint64_t *out_) {
    const fe *xs = (const fe *)xs_, *ys = (const fe *)ys_;
    fe *out = (fe *)out_;
    fe *num = (fe *)malloc((n + 1) * 32), *tmp = (fe *)malloc((n + 1) * 32);
    for (uint32_t i = 0; i < n; ++i) out[i] = f_zero();
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t deg = 0;
        num[0] = f_one(fid);
        for (uint32_t j = 0; j < n; ++j) {
            if (f_eq(&xs[j], &xs[k])) continue;      /* :110 `*x != *focus_x_point` */
            fe negx;
            f_neg(fid, &negx, &xs[j]);
            for (uint32_t i = 0; i <= deg + 1; ++i) tmp[i] = f_zero();
            for (uint32_t i = 0; i <= deg; ++i) {    /* multiply_polynomials(numerator, [-x, 1]) :142-162 */
                fe t;
                f_mul(fid, &t, &num[i], &negx);
                f_add(fid, &tmp[i], &tmp[i], &t);
                f_add(fid, &tmp[i + 1], &tmp[i + 1], &num[i]);
            }
            ++deg;
            memcpy(num, tmp, (deg + 1) * 32);
        }
        fe den, inv, scalar;
        zko_univariate_evaluate(fid, (const uint64_t *)num, deg + 1, xs[k].l, den.l); /* :118-119 */
        f_inv(fid, &inv, &den);
        f_mul(fid, &scalar, &ys[k], &inv);                                           /* :123 y / denominator */
        for (uint32_t i = 0; i <= deg; ++i) {
            fe t;
            f_mul(fid, &t, &scalar, &num[i]);
            f_add(fid, &out[i], &out[i], &t);
        }
    }
    free(num);
    free(tmp);
}

/* ------------------------------------------------------------------------------------------
 * Plain sumcheck -- sumcheck_protocol/src/basic_sumcheck/{prover,verifier}.rs
 * ------------------------------------------------------------------------------------------ */
/* split_polynomial_and_sum_each -- prover.rs:74-89 */
void zko_split_and_sum(int fid, const uint64_t *in, uint64_t len, uint64_t out[8]) {
    uint64_t mid = len / 2;
    zko_fe_sum(fid, in, mid, out);
    zko_fe_sum(fid, in + 4 * mid, len - mid, out + 4);
}
/* Prover::init + Prover::prove -- prover.rs:22-33,35-71 */
int zko_basic_prove(int fid, const uint64_t *table, uint64_t len, uint64_t claimed_sum[4],
                    uint64_t *round_polys, uint64_t *challenges, uint64_t final_eval[4]) {
    if (!is_pow2(len)) return -1;
    uint32_t n = ilog2_u64(len);
    zko_fe_sum(fid, table, len, claimed_sum);                       /* :28 */
    zko_transcript *t = zko_transcript_new();
    uint8_t *bytes = (uint8_t *)malloc(len * 32);
    zko_mle_to_bytes(fid, table, len, bytes);                       /* :38-39 */
    zko_transcript_append(t, bytes, len * 32);
    free(bytes);
    uint8_t b32[64];
    zko_fe_to_bytes_be(fid, claimed_sum, b32);                      /* :40-41 */
    zko_transcript_append(t, b32, 32);
    uint64_t *cur = (uint64_t *)malloc(len * 32);
    memcpy(cur, table, len * 32);                                   /* :44 clone */
    uint64_t m = len;
    for (uint32_t k = 0; k < n; ++k) {                              /* :46 */
        uint64_t *rp = round_polys + 8 * k;
        zko_split_and_sum(fid, cur, m, rp);                         /* :50 */
        zko_mle_to_bytes(fid, rp, 2, b32);                          /* :51-52 */
        zko_transcript_append(t, b32, 64);                          /* :55 */
        uint64_t r[4];
        zko_transcript_challenge(t, fid, r);                        /* :58 */
        if (challenges) memcpy(challenges + 4 * k, r, 32);
        uint64_t *nxt = (uint64_t *)malloc((m / 2) * 32);
        zko_mle_partial_evaluate(fid, cur, m, 0, r, nxt);           /* :61-63 */
        free(cur);
        cur = nxt;
        m /= 2;
    }
    if (final_eval) memcpy(final_eval, cur, 32);
    free(cur);
    zko_transcript_free(t);
    return 0;
}
/* Verifier::verify -- verifier.rs:23-71.  Round polynomials are 2-entry MLEs; p(x) = evaluate([x]). */
int zko_basic_verify(int fid, const uint64_t *table, uint64_t len, const uint64_t claimed_sum[4],
                     const uint64_t *round_polys, uint32_t n_rounds) {
    if (!is_pow2(len) || n_rounds != ilog2_u64(len)) return 0;      /* :26-30 */
    fe claim;
    memcpy(claim.l, claimed_sum, 32);
    zko_transcript *t = zko_transcript_new();
    uint8_t *bytes = (uint8_t *)malloc(len * 32);
    zko_mle_to_bytes(fid, table, len, bytes);                       /* :34-35 */
    zko_transcript_append(t, bytes, len * 32);
    free(bytes);
    uint8_t b32[64];
    zko_fe_to_bytes_be(fid, claimed_sum, b32);
    zko_transcript_append(t, b32, 32);                              /* :36-37 */
    uint64_t *chal = (uint64_t *)malloc((n_rounds ? n_rounds : 1) * 32);
    int ok = 1;
    fe zero = f_zero(), one = f_one(fid);
    for (uint32_t i = 0; i < n_rounds && ok; ++i) {
        const uint64_t *rp = round_polys + 8 * i;
        fe e0, e1, s;
        zko_mle_evaluate(fid, rp, 2, zero.l, 1, e0.l);              /* :48-52 */
        zko_mle_evaluate(fid, rp, 2, one.l, 1, e1.l);
        f_add(fid, &s, &e0, &e1);
        if (!f_eq(&s, &claim)) { ok = 0; break; }
        zko_mle_to_bytes(fid, rp, 2, b32);
        zko_transcript_append(t, b32, 64);                          /* :58-59 */
        zko_transcript_challenge(t, fid, chal + 4 * i);             /* :61 */
        zko_mle_evaluate(fid, rp, 2, chal + 4 * i, 1, claim.l);     /* :64 */
    }
    if (ok) {
        fe fin;
        zko_mle_evaluate(fid, table, len, chal, n_rounds, fin.l);   /* :67 */
        ok = f_eq(&fin, &claim);                                    /* :70 */
    }
    free(chal);
    zko_transcript_free(t);
    return ok;
}

/* ------------------------------------------------------------------------------------------
 * Product sumcheck -- sumcheck_protocol/src/gkr_sumcheck/sumcheck_gkr_protocol.rs
 * ------------------------------------------------------------------------------------------ */
/* SumPolynomial::partial_evaluate -- sum_polynomial.rs:40-53 -> product_polynomial.rs:36-54 */
static uint64_t *sumpoly_fold(int fid, const uint64_t *tables, uint32_t T, uint64_t len, const uint64_t r[4]) {
    uint64_t half = len / 2;
    uint64_t *out = (uint64_t *)malloc((uint64_t)T * (half ? half : 1) * 32);
    for (uint32_t t = 0; t < T; ++t)
        zko_mle_partial_evaluate(fid, tables + (uint64_t)t * len * 4, len, 0, r, out + (uint64_t)t * half * 4);
    return out;
}
/* generate_round_univariate -- sumcheck_gkr_protocol.rs:113-143 */
void zko_generate_round_univariate(int fid, const uint64_t *tables, uint32_t P, uint32_t D, uint64_t len, uint64_t *out_evals) {
    uint64_t half = len / 2;
    for (uint32_t i = 0; i <= D; ++i) {                                        /* :127 */
        fe x = f_from_u64(fid, i);                                             /* :128 */
        uint64_t *folded = sumpoly_fold(fid, tables, P * D, len, x.l);         /* :129 */
        uint64_t *red = (uint64_t *)malloc(half * 32);
        zko_sumpoly_reduce(fid, folded, P, D, half, red);                      /* :134 */
        zko_fe_sum(fid, red, half, out_evals + 4 * i);                         /* :135-137 */
        free(folded);
        free(red);
    }
}
/* prove -- sumcheck_gkr_protocol.rs:24-67 */
int zko_product_prove(int fid, const uint64_t *tables, uint32_t P, uint32_t D, uint64_t len,
                      const uint64_t claimed_sum[4], zko_transcript *t,
                      uint64_t *coeffs, uint64_t *challenges, uint64_t *final_tables) {
    if (!is_pow2(len) || P < 2 || D < 2) return -1;  /* asserts at sum_polynomial.rs:58-61, product_polynomial.rs:59-62 */
    uint32_t n = ilog2_u64(len), T = P * D;
    uint8_t b[32 * 16];
    zko_fe_to_bytes_be(fid, claimed_sum, b);
    zko_transcript_append(t, b, 32);                                           /* :35 */
    uint64_t *cur = (uint64_t *)malloc((uint64_t)T * len * 32);
    memcpy(cur, tables, (uint64_t)T * len * 32);                               /* :33 clone */
    uint64_t m = len;
    uint64_t *xs = (uint64_t *)malloc((D + 1) * 32), *ev = (uint64_t *)malloc((D + 1) * 32);
    for (uint32_t i = 0; i <= D; ++i) zko_fe_from_u64(fid, i, xs + 4 * i);     /* :46-48 */
    for (uint32_t k = 0; k < n; ++k) {                                         /* :37 */
        zko_generate_round_univariate(fid, cur, P, D, m, ev);                  /* :41 */
        uint64_t *c = coeffs + (uint64_t)k * (D + 1) * 4;
        zko_lagrange_interpolate(fid, xs, ev, D + 1, c);                       /* :49-50 */
        for (uint32_t i = 0; i <= D; ++i) zko_fe_to_bytes_le(fid, c + 4 * i, b + 32 * i); /* :145-150 */
        zko_transcript_append(t, b, 32 * (D + 1));                             /* :52 */
        uint64_t r[4];
        zko_transcript_challenge(t, fid, r);                                   /* :55 */
        uint64_t *nxt = sumpoly_fold(fid, cur, T, m, r);                       /* :57 */
        free(cur);
        cur = nxt;
        m /= 2;
        memcpy(challenges + 4 * k, r, 32);                                     /* :59 */
    }
    if (final_tables) memcpy(final_tables, cur, (uint64_t)T * 32);
    free(cur); free(xs); free(ev);
    return 0;
}
/* verify -- sumcheck_gkr_protocol.rs:69-106 */
int zko_product_verify(int fid, const uint64_t claimed_sum[4], const uint64_t *coeffs, uint32_t n_rounds, uint32_t D,
                       zko_transcript *t, uint64_t *challenges, uint64_t last_claim[4]) {
    uint8_t b[32 * 16];
    zko_fe_to_bytes_be(fid, claimed_sum, b);
    zko_transcript_append(t, b, 32);                                           /* :73 */
    fe cur, zero = f_zero(), one = f_one(fid);
    memcpy(cur.l, claimed_sum, 32);
    for (uint32_t k = 0; k < n_rounds; ++k) {
        const uint64_t *c = coeffs + (uint64_t)k * (D + 1) * 4;
        fe e0, e1, s;
        zko_univariate_evaluate(fid, c, D + 1, zero.l, e0.l);                  /* :81-82 */
        zko_univariate_evaluate(fid, c, D + 1, one.l, e1.l);
        f_add(fid, &s, &e0, &e1);
        if (!f_eq(&s, &cur)) { memcpy(last_claim, cur.l, 32); return 0; }      /* :84-90 */
        for (uint32_t i = 0; i <= D; ++i) zko_fe_to_bytes_le(fid, c + 4 * i, b + 32 * i);
        zko_transcript_append(t, b, 32 * (D + 1));                             /* :92 */
        uint64_t r[4];
        zko_transcript_challenge(t, fid, r);                                   /* :94 */
        zko_univariate_evaluate(fid, c, D + 1, r, cur.l);                      /* :96 */
        memcpy(challenges + 4 * k, r, 32);                                     /* :98 */
    }
    memcpy(last_claim, cur.l, 32);
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * Circuit -- circuit/src/arithmetic_circuit.rs
 * ------------------------------------------------------------------------------------------ */
/* evaluate -- arithmetic_circuit.rs:65-109.  Layers are walked input-side first (`.rev()`, :72);
 * gate results ACCUMULATE into output_index (:96). */
int zko_circuit_evaluate(int fid, const zko_circuit *c, const uint64_t *inputs, uint64_t n_inputs,
                         uint64_t *sizes, uint64_t *values, uint64_t values_cap) {
    uint32_t L = c->n_layers;
    /* first pass: sizes */
    sizes[L] = n_inputs;
    uint64_t total = n_inputs;
    for (uint32_t li = L; li-- > 0;) {
        uint64_t mx = 0;
        for (uint64_t g = c->layer_off[li]; g < c->layer_off[li + 1]; ++g)
            if (c->out[g] > mx) mx = c->out[g];
        sizes[li] = mx + 1;                                                    /* :73-80 */
        total += sizes[li];
    }
    if (total > values_cap) return -2;
    uint64_t *off = (uint64_t *)malloc((L + 1) * sizeof(uint64_t));
    off[0] = 0;
    for (uint32_t i = 0; i < L; ++i) off[i + 1] = off[i] + sizes[i];
    memcpy(values + 4 * off[L], inputs, n_inputs * 32);
    for (uint32_t li = L; li-- > 0;) {
        const fe *in = (const fe *)(values + 4 * off[li + 1]);
        fe *out = (fe *)(values + 4 * off[li]);
        for (uint64_t i = 0; i < sizes[li]; ++i) out[i] = f_zero();
        for (uint64_t g = c->layer_off[li]; g < c->layer_off[li + 1]; ++g) {
            if (c->left[g] >= sizes[li + 1] || c->right[g] >= sizes[li + 1]) { free(off); return -1; }
            fe v;
            if (c->op[g] == 0) f_add(fid, &v, &in[c->left[g]], &in[c->right[g]]);   /* :91 */
            else               f_mul(fid, &v, &in[c->left[g]], &in[c->right[g]]);   /* :92 */
            f_add(fid, &out[c->out[g]], &out[c->out[g]], &v);                       /* :96 */
        }
    }
    free(off);
    return 0;
}
/* num_of_layer_variables -- arithmetic_circuit.rs:166-178 */
uint32_t zko_num_of_layer_variables(uint32_t layer_index) {
    if (layer_index == 0) return 3;
    return layer_index + 2 * (layer_index + 1);
}
/* convert_to_binary_and_to_decimal -- arithmetic_circuit.rs:180-200.  `format!("{:0>width$b}")` pads
 * to at least `width` digits and never truncates; 0 prints as "0" even for width 0 (layer 0: one a-bit). */
static uint32_t printed_width(uint64_t v, uint32_t width) {
    uint32_t w = 1;
    while (v >> w) ++w;      /* number of binary digits of v (1 for v == 0) */
    return w > width ? w : width;
}
uint64_t zko_gate_position(uint32_t layer_index, uint64_t a, uint64_t b, uint64_t c) {
    uint32_t wb = printed_width(b, layer_index + 1), wc = printed_width(c, layer_index + 1);
    (void)printed_width(a, layer_index);
    return (a << (wb + wc)) | (b << wc) | c;
}
/* add_i_and_mul_i_mle -- arithmetic_circuit.rs:126-163 (`= one`, not `+=`) */
void zko_add_i_mul_i(int fid, const zko_circuit *c, uint32_t layer, uint64_t *add_i, uint64_t *mul_i) {
    uint64_t n = 1ull << zko_num_of_layer_variables(layer);
    memset(add_i, 0, n * 32);
    memset(mul_i, 0, n * 32);
    fe one = f_one(fid);
    for (uint64_t g = c->layer_off[layer]; g < c->layer_off[layer + 1]; ++g) {
        uint64_t pos = zko_gate_position(layer, c->out[g], c->left[g], c->right[g]);
        memcpy((c->op[g] == 0 ? add_i : mul_i) + 4 * pos, one.l, 32);
    }
}

/* ------------------------------------------------------------------------------------------
 * GKR -- gkr/src/gkr_protocol.rs, gkr/src/utils.rs
 * ------------------------------------------------------------------------------------------ */
static uint32_t gkr_rounds(uint32_t layer) { return zko_num_of_layer_variables(layer) - (layer == 0 ? 1 : layer); }
uint64_t zko_gkr_total_rounds(uint32_t n_layers) {
    uint64_t s = 0;
    for (uint32_t i = 0; i < n_layers; ++i) s += gkr_rounds(i);
    return s;
}
/* fold the leading `k` variables of `tbl` (len entries) at rs[0..k) -- the loops of utils.rs:38-58 */
static uint64_t *fold_leading(int fid, const uint64_t *tbl, uint64_t len, const uint64_t *rs, uint32_t k) {
    uint64_t *cur = (uint64_t *)malloc(len * 32);
    memcpy(cur, tbl, len * 32);
    for (uint32_t i = 0; i < k; ++i) {
        uint64_t *nxt = (uint64_t *)malloc((len / 2) * 32);
        zko_mle_partial_evaluate(fid, cur, len, 0, rs + 4 * i, nxt);
        free(cur);
        cur = nxt;
        len /= 2;
    }
    return cur;
}
/* compute_new_add_i_mul_i -- utils.rs:23-68 (one table at a time) */
static uint64_t *alpha_beta_fold(int fid, const uint64_t *abc, uint64_t len, const uint64_t *rb, const uint64_t *rc,
                                 uint32_t k, const uint64_t alpha[4], const uint64_t beta[4]) {
    uint64_t *frb = fold_leading(fid, abc, len, rb, k);
    uint64_t *frc = fold_leading(fid, abc, len, rc, k);
    uint64_t out_len = len >> k;
    uint64_t *out = (uint64_t *)malloc(out_len * 32);
    zko_mle_scalar_mul(fid, frb, out_len, alpha, frb);   /* :59-66 scalar_mul(alpha) + scalar_mul(beta) */
    zko_mle_scalar_mul(fid, frc, out_len, beta, frc);
    zko_mle_add(fid, frb, frc, out_len, out);
    free(frb);
    free(frc);
    return out;
}
/* compute_fbc_polynomial -- utils.rs:8-21: tables in SumPolynomial order
 *   product 0 = [add_i_bc, W_b (+) W_c], product 1 = [mul_i_bc, W_b (x) W_c]. */
static uint64_t *build_fbc(int fid, const uint64_t *add_bc, const uint64_t *mul_bc, const uint64_t *w, uint64_t wlen) {
    uint64_t n = wlen * wlen;
    uint64_t *t = (uint64_t *)malloc(4 * n * 32);
    memcpy(t, add_bc, n * 32);
    zko_mle_tensor_add(fid, w, w, wlen, t + 4 * n);
    memcpy(t + 8 * n, mul_bc, n * 32);
    zko_mle_tensor_mul(fid, w, w, wlen, t + 12 * n);
    return t;
}
/* prove -- gkr_protocol.rs:26-143 */
int zko_gkr_prove(int fid, const zko_circuit *c, const uint64_t *inputs, uint64_t n_inputs, zko_gkr_proof *pf) {
    uint32_t L = c->n_layers;
    uint64_t *sizes = (uint64_t *)malloc((L + 1) * sizeof(uint64_t));
    uint64_t cap = n_inputs;
    for (uint32_t i = 0; i < L; ++i) {
        uint64_t mx = 0;
        for (uint64_t g = c->layer_off[i]; g < c->layer_off[i + 1]; ++g) if (c->out[g] > mx) mx = c->out[g];
        cap += mx + 1;
    }
    uint64_t *values = (uint64_t *)malloc(cap * 32);
    int rc = zko_circuit_evaluate(fid, c, inputs, n_inputs, sizes, values, cap);          /* :27 */
    if (rc) { free(sizes); free(values); return rc; }
    uint64_t *off = (uint64_t *)malloc((L + 2) * sizeof(uint64_t));
    off[0] = 0;
    for (uint32_t i = 0; i <= L; ++i) off[i + 1] = off[i] + sizes[i];
    for (uint32_t i = 0; i <= L; ++i)
        if (!is_pow2(sizes[i])) { free(sizes); free(values); free(off); return -1; }       /* MLE::new assert */

    zko_transcript *t = zko_transcript_new();
    fe alpha = f_zero(), beta = f_zero();
    uint64_t *rb = NULL, *rcv = NULL;
    uint32_t nrb = 0;

    /* layer 0 -- :39-51 */
    uint64_t w0len = sizes[0];
    uint64_t *w0 = (uint64_t *)malloc((w0len < 2 ? 2 : w0len) * 32);
    memcpy(w0, values, w0len * 32);
    if (w0len == 1) { memset(w0 + 4, 0, 32); w0len = 2; }                                  /* :43-47 */
    uint8_t *bytes = (uint8_t *)malloc(w0len * 32);
    zko_mle_to_bytes(fid, w0, w0len, bytes);
    zko_transcript_append(t, bytes, w0len * 32);                                           /* :49 */
    free(bytes);
    fe ra, claimed;
    zko_transcript_challenge(t, fid, ra.l);                                                /* :50 */
    if (zko_mle_evaluate(fid, w0, w0len, ra.l, 1, claimed.l)) return -1;                   /* :51 */
    free(w0);

    memcpy(pf->output, values, sizes[0] * 32);
    pf->n_output = sizes[0];
    uint64_t round_off = 0;
    for (uint32_t li = 0; li < L; ++li) {                                                  /* :57 */
        uint64_t nabc = 1ull << zko_num_of_layer_variables(li);
        uint64_t *add_abc = (uint64_t *)malloc(nabc * 32), *mul_abc = (uint64_t *)malloc(nabc * 32);
        zko_add_i_mul_i(fid, c, li, add_abc, mul_abc);                                     /* :58 */
        uint64_t *add_bc, *mul_bc;
        if (li == 0) {                                                                     /* :60-72 */
            add_bc = fold_leading(fid, add_abc, nabc, ra.l, 1);
            mul_bc = fold_leading(fid, mul_abc, nabc, ra.l, 1);
        } else {                                                                           /* :73-82 */
            if (nrb != li) return -3;
            add_bc = alpha_beta_fold(fid, add_abc, nabc, rb, rcv, li, alpha.l, beta.l);
            mul_bc = alpha_beta_fold(fid, mul_abc, nabc, rb, rcv, li, alpha.l, beta.l);
        }
        free(add_abc); free(mul_abc);
        const uint64_t *w = values + 4 * off[li + 1];                                      /* :88-89 */
        uint64_t wlen = sizes[li + 1];
        uint32_t rounds = gkr_rounds(li);
        if (wlen * wlen != (1ull << rounds)) return -4; /* reference would trip "different number of variables" */
        uint64_t *fbc = build_fbc(fid, add_bc, mul_bc, w, wlen);                           /* :95 */
        free(add_bc); free(mul_bc);
        memcpy(pf->layer_claims + 4 * li, claimed.l, 32);
        uint64_t *chal = pf->challenges + 4 * round_off;
        zko_product_prove(fid, fbc, 2, 2, wlen * wlen, claimed.l, t, pf->coeffs + 12 * round_off, chal, NULL); /* :99 */
        free(fbc);
        if (li < L - 1) {                                                                  /* :109 */
            uint32_t mid = rounds / 2;
            fe wbv, wcv;
            zko_mle_evaluate(fid, w, wlen, chal, mid, wbv.l);                              /* utils.rs:70-82 */
            zko_mle_evaluate(fid, w, wlen, chal + 4 * mid, rounds - mid, wcv.l);
            memcpy(pf->wb + 4 * li, wbv.l, 32);
            memcpy(pf->wc + 4 * li, wcv.l, 32);
            rb = chal; rcv = chal + 4 * mid; nrb = mid;                                    /* :120-123 */
            uint8_t b[32];
            zko_fe_to_bytes_be(fid, wbv.l, b);
            zko_transcript_append(t, b, 32);
            zko_transcript_challenge(t, fid, alpha.l);                                     /* :125-126 */
            zko_fe_to_bytes_be(fid, wcv.l, b);
            zko_transcript_append(t, b, 32);
            zko_transcript_challenge(t, fid, beta.l);                                      /* :128-129 */
            fe x, y;
            f_mul(fid, &x, &alpha, &wbv);
            f_mul(fid, &y, &beta, &wcv);
            f_add(fid, &claimed, &x, &y);                                                  /* :132 */
        }
        round_off += rounds;
    }
    memcpy(pf->claimed_sum, claimed.l, 32);
    zko_transcript_free(t);
    free(sizes); free(values); free(off);
    return 0;
}
/* verify -- gkr_protocol.rs:146-236, claim helpers utils.rs:84-135 */
int zko_gkr_verify(int fid, const zko_circuit *c, const zko_gkr_proof *pf, const uint64_t *inputs, uint64_t n_inputs) {
    uint32_t L = c->n_layers;
    zko_transcript *t = zko_transcript_new();
    fe alpha = f_zero(), beta = f_zero();
    uint64_t w0len = pf->n_output;
    uint64_t *w0 = (uint64_t *)malloc((w0len < 2 ? 2 : w0len) * 32);
    memcpy(w0, pf->output, w0len * 32);
    if (w0len == 1) { memset(w0 + 4, 0, 32); w0len = 2; }                                  /* :153-159 */
    uint8_t *bytes = (uint8_t *)malloc(w0len * 32);
    zko_mle_to_bytes(fid, w0, w0len, bytes);
    zko_transcript_append(t, bytes, w0len * 32);                                           /* :161 */
    free(bytes);
    fe ra, claimed;
    zko_transcript_challenge(t, fid, ra.l);
    if (zko_mle_evaluate(fid, w0, w0len, ra.l, 1, claimed.l)) { free(w0); return 0; }      /* :164 */
    free(w0);
    uint64_t round_off = 0;
    uint64_t *prev = NULL;
    uint32_t nprev = 0;
    int ok = 1;
    for (uint32_t li = 0; li < L && ok; ++li) {
        uint32_t rounds = gkr_rounds(li);
        if (memcmp(claimed.l, pf->layer_claims + 4 * li, 32)) { ok = 0; break; }           /* :167-169 */
        uint64_t *chal = (uint64_t *)malloc(rounds * 32);
        fe last;
        if (!zko_product_verify(fid, pf->layer_claims + 4 * li, pf->coeffs + 12 * round_off, rounds, 2, t, chal, last.l)) {
            free(chal); ok = 0; break;                                                     /* :172-176 */
        }
        uint32_t mid = rounds / 2;
        fe wbv, wcv;
        if (li < L - 1) {                                                                  /* :183-187 */
            memcpy(wbv.l, pf->wb + 4 * li, 32);
            memcpy(wcv.l, pf->wc + 4 * li, 32);
        } else {                                                                           /* :188-194 */
            if (!is_pow2(n_inputs) || n_inputs != (1ull << mid)) { free(chal); ok = 0; break; }
            zko_mle_evaluate(fid, inputs, n_inputs, chal, mid, wbv.l);
            zko_mle_evaluate(fid, inputs, n_inputs, chal + 4 * mid, rounds - mid, wcv.l);
        }
        uint64_t nabc = 1ull << zko_num_of_layer_variables(li);
        uint64_t *add_abc = (uint64_t *)malloc(nabc * 32), *mul_abc = (uint64_t *)malloc(nabc * 32);
        zko_add_i_mul_i(fid, c, li, add_abc, mul_abc);
        uint64_t *add_bc, *mul_bc;
        if (li == 0) {                                                                     /* utils.rs:84-111 */
            add_bc = fold_leading(fid, add_abc, nabc, ra.l, 1);
            mul_bc = fold_leading(fid, mul_abc, nabc, ra.l, 1);
        } else {                                                                           /* utils.rs:113-135 */
            uint32_t pm = nprev / 2;
            add_bc = alpha_beta_fold(fid, add_abc, nabc, prev, prev + 4 * pm, li, alpha.l, beta.l);
            mul_bc = alpha_beta_fold(fid, mul_abc, nabc, prev, prev + 4 * pm, li, alpha.l, beta.l);
        }
        free(add_abc); free(mul_abc);
        fe addr, mulr, s, pr, x, y, expect;
        zko_mle_evaluate(fid, add_bc, 1ull << rounds, chal, rounds, addr.l);
        zko_mle_evaluate(fid, mul_bc, 1ull << rounds, chal, rounds, mulr.l);
        free(add_bc); free(mul_bc);
        f_add(fid, &s, &wbv, &wcv);
        f_mul(fid, &pr, &wbv, &wcv);
        f_mul(fid, &x, &addr, &s);
        f_mul(fid, &y, &mulr, &pr);
        f_add(fid, &expect, &x, &y);                                                       /* utils.rs:110,134 */
        if (!f_eq(&expect, &last)) { free(chal); ok = 0; break; }                          /* :220-222 */
        free(prev);
        prev = chal; nprev = rounds;                                                       /* :224 */
        uint8_t b[32];
        zko_fe_to_bytes_be(fid, wbv.l, b);
        zko_transcript_append(t, b, 32);
        zko_transcript_challenge(t, fid, alpha.l);                                         /* :226-227 */
        zko_fe_to_bytes_be(fid, wcv.l, b);
        zko_transcript_append(t, b, 32);
        zko_transcript_challenge(t, fid, beta.l);                                          /* :229-230 */
        f_mul(fid, &x, &alpha, &wbv);
        f_mul(fid, &y, &beta, &wcv);
        f_add(fid, &claimed, &x, &y);                                                      /* :232 */
        round_off += rounds;
    }
    free(prev);
    zko_transcript_free(t);
    return ok;
}

/* ------------------------------------------------------------------------------------------
 * GKR over layers of EXPLICIT width, wiring predicates evaluated from the gate list
 * (gkr/src/gkr_protocol.rs:26-143 prove, :146-236 verify; gkr/src/utils.rs:8-135).
 *
 * The reference stores add_i / mul_i as dense 2^(3i+2) tables (arithmetic_circuit.rs:126-163) and W(b)+W(c),
 * W(b)W(c) as 4^(i+1) tensors (utils.rs:8-21), which caps it at widths of ~2^7.  The functions of the protocol are
 * the same whatever their storage: with f(b,c) = add(b,c)(W(b)+W(c)) + mul(b,c)W(b)W(c) the round polynomial of
 * round k is  s_k(X) = sum over boolean x of f(r_0..r_{k-1}, X, x)  (sumcheck_gkr_protocol.rs:113-143), and since
 * add/mul are sums of indicators -- one per gate -- every term of that sum belongs to exactly one gate:
 *
 *   s_k(X) = sum_g w(out_g) . eq(prefix of (l_g, r_g), r_0..r_{k-1}) . eq1(bit k of (l_g, r_g), X)
 *                  . op_g( W~(b-part at (r.., X, rest of l_g)), W~(c-part ...) )
 *
 * where w(a) binds the `a` variables: eq(r_a, a) at the output layer (gkr_protocol.rs:60-72), and
 * alpha eq(r_b, a) + beta eq(r_c, a) below (utils.rs:23-68).  That is O(gates) per round with no dense table, and it
 * is a DIFFERENT algorithm from the product's two-phase bucketed tables (csrc/gkr_wide.cu): per-gate running prefix
 * weights here, per-wire h tables there.  Pinned to zko_gkr_prove (the dense restatement) on reference-shaped
 * circuits by tests/test_oracle.py; generalisation of the reference's single output challenge r_a to layer_bits[0]
 * successive challenges as in the product.  Gate lists must be duplicate-free (the dense indicator stores `= one`).
 * ------------------------------------------------------------------------------------------ */
static void eq_at_index(int fid, fe *out, const uint64_t *rs, uint32_t k, uint64_t idx) {
    /* eq(r, idx) = prod_v (bit_v ? r_v : 1 - r_v), variable 0 = most significant bit of idx */
    fe acc = f_one(fid), one = f_one(fid);
    for (uint32_t v = 0; v < k; ++v) {
        fe r, t;
        memcpy(r.l, rs + 4 * v, 32);
        if ((idx >> (k - 1 - v)) & 1) t = r;
        else f_sub(fid, &t, &one, &r);
        f_mul(fid, &acc, &acc, &t);
    }
    *out = acc;
}
/* value at X of the line through (0 -> a0) and (1 -> a1), X a small integer given as a field element */
static inline void line_at(int fid, fe *out, const fe *a0, const fe *a1, const fe *x) {
    fe d, t;
    f_sub(fid, &d, a1, a0);
    f_mul(fid, &t, x, &d);
    f_add(fid, out, a0, &t);
}
/* eq1(bit, X) = bit ? X : 1 - X */
static inline void eq1_at(int fid, fe *out, int bit, const fe *x) {
    if (bit) { *out = *x; return; }
    fe one = f_one(fid);
    f_sub(fid, out, &one, x);
}
static int sparse_layer_values(int fid, uint32_t L, const uint32_t *bits, const uint64_t *layer_off, const uint32_t *left,
                               const uint32_t *right, const uint32_t *out, const uint8_t *op, const uint64_t *inputs,
                               uint64_t n_inputs, fe **W) {
    /* Circuit::evaluate, arithmetic_circuit.rs:65-109: input side first, gate results accumulate (`+=`, :96) */
    if (n_inputs != (1ull << bits[L])) return -1;
    W[L] = (fe *)malloc(n_inputs * sizeof(fe));
    memcpy(W[L], inputs, n_inputs * 32);
    for (uint32_t li = L; li-- > 0;) {
        uint64_t n_out = 1ull << bits[li], n_in = 1ull << bits[li + 1];
        W[li] = (fe *)calloc(n_out, sizeof(fe));
        for (uint64_t g = layer_off[li]; g < layer_off[li + 1]; ++g) {
            if (out[g] >= n_out || left[g] >= n_in || right[g] >= n_in) return -2;
            fe v;
            if (op[g] == 0) f_add(fid, &v, &W[li + 1][left[g]], &W[li + 1][right[g]]);
            else f_mul(fid, &v, &W[li + 1][left[g]], &W[li + 1][right[g]]);
            f_add(fid, &W[li][out[g]], &W[li][out[g]], &v);
        }
    }
    return 0;
}
/* proof layout as zko_gkr_proof: output = 2^layer_bits[0] elements, rounds of layer li = 2 layer_bits[li+1].
 * layer_bits[0] == 0 (a single output) is the reference's padded case: [out, 0], one challenge (gkr_protocol.rs:43-47). */
int zko_gkr_prove_sparse(int fid, uint32_t L, const uint32_t *layer_bits, const uint64_t *layer_off, const uint32_t *left,
                         const uint32_t *right, const uint32_t *out, const uint8_t *op, const uint64_t *inputs,
                         uint64_t n_inputs, zko_gkr_proof *pf) {
    if (L == 0) return -1;
    uint32_t *bits = (uint32_t *)malloc((L + 1) * sizeof(uint32_t));
    memcpy(bits, layer_bits, (L + 1) * sizeof(uint32_t));
    const int padded = bits[0] == 0;
    if (padded) bits[0] = 1;                                        /* [out, 0]: index 1 holds the zero no gate drives */
    for (uint32_t li = 1; li <= L; ++li) if (bits[li] == 0) { free(bits); return -1; }
    fe **W = (fe **)calloc(L + 1, sizeof(fe *));
    int rc = sparse_layer_values(fid, L, bits, layer_off, left, right, out, op, inputs, n_inputs, W);
    if (rc) return rc;
    const fe one = f_one(fid);
    fe xs[3];
    for (int i = 0; i < 3; ++i) xs[i] = f_from_u64(fid, (uint64_t)i);
    zko_transcript *t = zko_transcript_new();
    /* output layer -- gkr_protocol.rs:39-51 */
    const uint64_t n0 = 1ull << bits[0];
    memcpy(pf->output, W[0], (padded ? 1 : n0) * 32);
    pf->n_output = padded ? 1 : n0;
    {
        uint8_t *bytes = (uint8_t *)malloc(n0 * 32);
        zko_mle_to_bytes(fid, (const uint64_t *)W[0], n0, bytes);
        zko_transcript_append(t, bytes, n0 * 32);                                          /* :49 */
        free(bytes);
    }
    uint64_t *ra = (uint64_t *)malloc(bits[0] * 32);
    for (uint32_t v = 0; v < bits[0]; ++v) zko_transcript_challenge(t, fid, ra + 4 * v);   /* :50 */
    fe claimed;
    if (zko_mle_evaluate(fid, (const uint64_t *)W[0], n0, ra, bits[0], claimed.l)) return -1;   /* :51 */
    fe alpha = f_zero(), beta = f_zero();
    const uint64_t *rb = NULL, *rcv = NULL;
    uint64_t round_off = 0;
    for (uint32_t li = 0; li < L; ++li) {                                                  /* :57 */
        const uint32_t m = bits[li + 1], ka = bits[li];
        const uint64_t g0 = layer_off[li], ng = layer_off[li + 1] - g0, nm = 1ull << m;
        /* per-gate weight w(out_g): the `a` variables bound -- :60-82, utils.rs:23-68 */
        fe *pw = (fe *)malloc((ng ? ng : 1) * sizeof(fe));
        for (uint64_t g = 0; g < ng; ++g) {
            if (li == 0) eq_at_index(fid, &pw[g], ra, ka, out[g0 + g]);
            else {
                fe eb, ec, x, y;
                eq_at_index(fid, &eb, rb, ka, out[g0 + g]);
                eq_at_index(fid, &ec, rcv, ka, out[g0 + g]);
                f_mul(fid, &x, &alpha, &eb);
                f_mul(fid, &y, &beta, &ec);
                f_add(fid, &pw[g], &x, &y);
            }
        }
        memcpy(pf->layer_claims + 4 * li, claimed.l, 32);
        uint64_t *chal = pf->challenges + 4 * round_off;
        uint64_t *coeffs = pf->coeffs + 12 * round_off;
        /* sumcheck over (b, c): prove -- sumcheck_gkr_protocol.rs:24-67 */
        uint8_t by[96];
        zko_fe_to_bytes_be(fid, claimed.l, by);
        zko_transcript_append(t, by, 32);                                                  /* :35 */
        fe *Wb = (fe *)malloc(nm * sizeof(fe));     /* W over the not-yet-bound b variables */
        fe *Wc = (fe *)malloc(nm * sizeof(fe));     /* W over the not-yet-bound c variables */
        memcpy(Wb, W[li + 1], nm * sizeof(fe));
        memcpy(Wc, W[li + 1], nm * sizeof(fe));
        const fe *Wfull = W[li + 1];
        fe Wu = f_zero();
        for (uint32_t k = 0; k < 2 * m; ++k) {                                             /* :37 */
            const int phase_c = k >= m;
            const uint32_t kk = phase_c ? k - m : k;           /* variable inside its half */
            const uint32_t rest = m - kk - 1;                  /* unbound variables after it */
            const fe *tab = phase_c ? Wc : Wb;
            fe ev[3] = {f_zero(), f_zero(), f_zero()};
            for (uint64_t g = 0; g < ng; ++g) {
                const uint32_t idx = phase_c ? right[g0 + g] : left[g0 + g];
                const int bit = (idx >> rest) & 1;
                const uint64_t low = idx & ((1ull << rest) - 1);
                for (int X = 0; X < 3; ++X) {                                              /* :127-137, X = 0..d */
                    fe e1, wv, other, val, term;
                    eq1_at(fid, &e1, bit, &xs[X]);
                    line_at(fid, &wv, &tab[low], &tab[(1ull << rest) + low], &xs[X]);
                    other = phase_c ? Wu : Wfull[right[g0 + g]];   /* the factor the sum over boolean c pins to W(r_g) */
                    if (op[g0 + g] == 0) f_add(fid, &val, &wv, &other);
                    else f_mul(fid, &val, &wv, &other);
                    f_mul(fid, &term, &pw[g], &e1);
                    f_mul(fid, &term, &term, &val);
                    f_add(fid, &ev[X], &ev[X], &term);
                }
            }
            uint64_t *c = coeffs + 12ull * k;
            zko_lagrange_interpolate(fid, (const uint64_t *)xs, (const uint64_t *)ev, 3, c);   /* :46-50 */
            for (int i = 0; i < 3; ++i) zko_fe_to_bytes_le(fid, c + 4 * i, by + 32 * i);   /* :145-150 */
            zko_transcript_append(t, by, 96);                                              /* :52 */
            fe r;
            zko_transcript_challenge(t, fid, r.l);                                         /* :55 */
            memcpy(chal + 4 * k, r.l, 32);                                                 /* :59 */
            /* bind the variable -- :57: prefix weights, then the W table of this half */
            for (uint64_t g = 0; g < ng; ++g) {
                const uint32_t idx = phase_c ? right[g0 + g] : left[g0 + g];
                fe e1;
                eq1_at(fid, &e1, (idx >> rest) & 1, &r);
                f_mul(fid, &pw[g], &pw[g], &e1);
            }
            fe *tw = phase_c ? Wc : Wb;
            for (uint64_t j = 0; j < (1ull << rest); ++j) line_at(fid, &tw[j], &tw[j], &tw[(1ull << rest) + j], &r);
            if (!phase_c && rest == 0) Wu = Wb[0];                                         /* W(r_b) */
        }
        const fe Wv = Wc[0];                                                               /* W(r_c) */
        (void)one;
        free(Wb); free(Wc); free(pw);
        if (li + 1 < L) {                                                                  /* :109-132 */
            memcpy(pf->wb + 4 * li, Wu.l, 32);                                             /* utils.rs:70-82 */
            memcpy(pf->wc + 4 * li, Wv.l, 32);
            rb = chal; rcv = chal + 4 * m;                                                 /* :120-123 */
            zko_fe_to_bytes_be(fid, Wu.l, by);
            zko_transcript_append(t, by, 32);
            zko_transcript_challenge(t, fid, alpha.l);                                     /* :125-126 */
            zko_fe_to_bytes_be(fid, Wv.l, by);
            zko_transcript_append(t, by, 32);
            zko_transcript_challenge(t, fid, beta.l);                                      /* :128-129 */
            fe x, y;
            f_mul(fid, &x, &alpha, &Wu);
            f_mul(fid, &y, &beta, &Wv);
            f_add(fid, &claimed, &x, &y);                                                  /* :132 */
        }
        round_off += 2ull * m;
    }
    memcpy(pf->claimed_sum, claimed.l, 32);
    zko_transcript_free(t);
    for (uint32_t li = 0; li <= L; ++li) free(W[li]);
    free(W); free(ra); free(bits);
    return 0;
}
/* verify -- gkr_protocol.rs:146-236 with the claim helpers of utils.rs:84-135 evaluated from the gate list:
 * add_i(r_a.., r_b, r_c) = sum over add gates of w(out_g) eq(r_b, l_g) eq(r_c, r_g).  Returns 1 iff accepted. */
int zko_gkr_verify_sparse(int fid, uint32_t L, const uint32_t *layer_bits, const uint64_t *layer_off, const uint32_t *left,
                          const uint32_t *right, const uint32_t *out, const uint8_t *op, const zko_gkr_proof *pf,
                          const uint64_t *inputs, uint64_t n_inputs) {
    if (L == 0) return 0;
    uint32_t *bits = (uint32_t *)malloc((L + 1) * sizeof(uint32_t));
    memcpy(bits, layer_bits, (L + 1) * sizeof(uint32_t));
    const int padded = bits[0] == 0;
    if (padded) bits[0] = 1;
    zko_transcript *t = zko_transcript_new();
    const uint64_t n0 = 1ull << bits[0];
    if (pf->n_output != (padded ? 1 : n0)) { free(bits); zko_transcript_free(t); return 0; }
    uint64_t *w0 = (uint64_t *)calloc(n0, 32);
    memcpy(w0, pf->output, pf->n_output * 32);                                             /* :153-159 */
    uint8_t *bytes = (uint8_t *)malloc(n0 * 32);
    zko_mle_to_bytes(fid, w0, n0, bytes);
    zko_transcript_append(t, bytes, n0 * 32);                                              /* :161 */
    free(bytes);
    uint64_t *ra = (uint64_t *)malloc(bits[0] * 32);
    for (uint32_t v = 0; v < bits[0]; ++v) zko_transcript_challenge(t, fid, ra + 4 * v);
    fe claimed, alpha = f_zero(), beta = f_zero();
    zko_mle_evaluate(fid, w0, n0, ra, bits[0], claimed.l);                                 /* :164 */
    free(w0);
    uint64_t round_off = 0;
    uint64_t *prev = NULL;
    int ok = 1;
    for (uint32_t li = 0; li < L && ok; ++li) {
        const uint32_t m = bits[li + 1], ka = bits[li], rounds = 2 * m;
        const uint64_t g0 = layer_off[li], ng = layer_off[li + 1] - g0;
        if (memcmp(claimed.l, pf->layer_claims + 4 * li, 32)) { ok = 0; break; }           /* :167-169 */
        uint64_t *chal = (uint64_t *)malloc(rounds * 32);
        fe last;
        if (!zko_product_verify(fid, pf->layer_claims + 4 * li, pf->coeffs + 12 * round_off, rounds, 2, t, chal, last.l)) {
            free(chal); ok = 0; break;                                                     /* :172-176 */
        }
        fe wbv, wcv;
        if (li + 1 < L) {                                                                  /* :183-187 */
            memcpy(wbv.l, pf->wb + 4 * li, 32);
            memcpy(wcv.l, pf->wc + 4 * li, 32);
        } else {                                                                           /* :188-194 */
            if (n_inputs != (1ull << m)) { free(chal); ok = 0; break; }
            zko_mle_evaluate(fid, inputs, n_inputs, chal, m, wbv.l);
            zko_mle_evaluate(fid, inputs, n_inputs, chal + 4 * m, m, wcv.l);
        }
        fe addr = f_zero(), mulr = f_zero();
        for (uint64_t g = 0; g < ng; ++g) {                                                /* utils.rs:84-135 */
            fe w, eb, ec, term;
            if (li == 0) eq_at_index(fid, &w, ra, ka, out[g0 + g]);
            else {
                fe x, y, e1, e2;
                eq_at_index(fid, &e1, prev, ka, out[g0 + g]);
                eq_at_index(fid, &e2, prev + 4 * ka, ka, out[g0 + g]);
                f_mul(fid, &x, &alpha, &e1);
                f_mul(fid, &y, &beta, &e2);
                f_add(fid, &w, &x, &y);
            }
            eq_at_index(fid, &eb, chal, m, left[g0 + g]);
            eq_at_index(fid, &ec, chal + 4 * m, m, right[g0 + g]);
            f_mul(fid, &term, &eb, &ec);
            f_mul(fid, &term, &term, &w);
            if (op[g0 + g] == 0) f_add(fid, &addr, &addr, &term);
            else f_add(fid, &mulr, &mulr, &term);
        }
        fe s, pr, x, y, expect;
        f_add(fid, &s, &wbv, &wcv);
        f_mul(fid, &pr, &wbv, &wcv);
        f_mul(fid, &x, &addr, &s);
        f_mul(fid, &y, &mulr, &pr);
        f_add(fid, &expect, &x, &y);                                                       /* utils.rs:110,134 */
        if (!f_eq(&expect, &last)) { free(chal); ok = 0; break; }                          /* :220-222 */
        free(prev);
        prev = chal;                                                                       /* :224 */
        uint8_t b[32];
        zko_fe_to_bytes_be(fid, wbv.l, b);
        zko_transcript_append(t, b, 32);
        zko_transcript_challenge(t, fid, alpha.l);                                         /* :226-227 */
        zko_fe_to_bytes_be(fid, wcv.l, b);
        zko_transcript_append(t, b, 32);
        zko_transcript_challenge(t, fid, beta.l);                                          /* :229-230 */
        f_mul(fid, &x, &alpha, &wbv);
        f_mul(fid, &y, &beta, &wcv);
        f_add(fid, &claimed, &x, &y);                                                      /* :232 */
        round_off += rounds;
    }
    free(prev); free(ra); free(bits);
    zko_transcript_free(t);
    return ok;
}

/* ------------------------------------------------------------------------------------------
 * Synthetic tables -- the bench workload's inputs, regenerated on the host exactly as the
 * CUDA generator makes them: entry i of table `table_id` under `seed` = from_le_bytes_mod_order of the 32 bytes
 * le64(w0) | le64(w1) | le64(w2) | le64(w3), w_l = splitmix64 stream of (seed ^ table_id * GOLDEN) at counter 4 i + l.
 * Local entry j is global entry first + j * step (a shard).  Not part of the reference: it turns bytes into elements
 * the way the reference does (fiat_shamir_transcript.rs:42 `from_le_bytes_mod_order`).
 * ------------------------------------------------------------------------------------------ */
static inline uint64_t splitmix64_at(uint64_t base, uint64_t ctr) {
    uint64_t z = base + (ctr + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void zko_table_generate(int fid, uint64_t seed, uint64_t table_id, uint64_t n, uint64_t first, uint64_t step, uint64_t *out) {
    const uint64_t base = seed ^ (table_id * 0x9E3779B97F4A7C15ull);
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1 && n >= 4096)
    for (uint64_t j = 0; j < n; ++j) {
        const uint64_t g = first + j * step;
        uint64_t w[4];
        for (int l = 0; l < 4; ++l) w[l] = splitmix64_at(base, 4 * g + (uint64_t)l);
        zko_fe_from_le_bytes_mod_order(fid, (const uint8_t *)w, 32, out + 4 * j);   /* little-endian host */
    }
}
