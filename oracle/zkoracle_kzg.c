/*
 * zkoracle_kzg.c -- CPU restatement of the reference's multilinear KZG (the input commitment of succinct GKR).
 * TEST INFRASTRUCTURE ONLY (see zkoracle.h for the rules).
 *
 * Reference functions restated (paths relative to the reference root):
 *   multilinear_kzg/src/trusted_setup.rs:26-52   compute_lagrange_basis
 *   multilinear_kzg/src/trusted_setup.rs:54-63   compute_g1_powers_of_tau
 *   multilinear_kzg/src/multilinear_kzg.rs:25-46   commit_to_polynomial
 *   multilinear_kzg/src/multilinear_kzg.rs:51-127  open_and_prove (+ :166-214 quotient / blow_up / expand_vec)
 *   multilinear_kzg/src/multilinear_kzg.rs:132-159 verify -- restated WITHOUT the pairing: the same equation is checked in
 *       G1 with the toxic waste known (zko_kzg_verify_trapdoor); the pairing form lives in oracle/pykzg.py.
 *
 * The curve arithmetic itself is in third-party crates that are not under /root/reference (ark-ec 0.5.0,
 * ark-bls12-381 0.5.0, multilinear_kzg/Cargo.toml): BLS12-381 G1, y^2 = x^3 + 4 over the 381-bit Fq, restated from the
 * published parameters: 6x64-limb Montgomery field (R = 2^384) like ark-ff's MontBackend<_,6>, Jacobian coordinates,
 * `mul_bigint` as MSB-first double-and-add.  Group elements are compared in affine form, which is canonical, so the
 * coordinate system and the order of the additions cannot show in a result.
 *
 * PARITY STATUS: the reference pins the Lagrange basis by known answers (trusted_setup.rs:101-126, reproduced in
 * tests/test_oracle_kzg.py) and everything else only through verify() == true (multilinear_kzg.rs:218-303); those three
 * tests are reproduced with the pairing of oracle/pykzg.py, and this file is pinned to pykzg point for point.
 *
 * Point layout: affine (x, y), 6 + 6 uint64 little-endian limbs, Montgomery form, canonical; the point at infinity is
 * x = y = 0 (not on the curve).  Scalars: BLS12-381 Fr elements in Montgomery form as everywhere else in the oracle.
 */
#include "zkoracle.h"
#include "curve_consts.h"
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[6]; } fq;
typedef struct { fq x, y, z; } jac;    /* z == 0: infinity */

static inline int q_ge(const uint64_t a[6]) {
    for (int i = 5; i >= 0; --i) {
        if (a[i] > ZKC_Q_64[i]) return 1;
        if (a[i] < ZKC_Q_64[i]) return 0;
    }
    return 1;
}
static inline uint64_t q_sub6(uint64_t r[6], const uint64_t a[6], const uint64_t b[6]) {
    u128 borrow = 0;
    for (int i = 0; i < 6; ++i) {
        u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
    return (uint64_t)borrow;
}
static inline void q_add6(uint64_t r[6], const uint64_t a[6], const uint64_t b[6]) {
    u128 c = 0;
    for (int i = 0; i < 6; ++i) {
        c += (u128)a[i] + b[i];
        r[i] = (uint64_t)c;
        c >>= 64;
    }
}
static inline void q_add(fq *r, const fq *a, const fq *b) {
    uint64_t t[6];
    q_add6(t, a->l, b->l);                 /* 2q < 2^384: no carry out */
    if (q_ge(t)) q_sub6(t, t, ZKC_Q_64);
    memcpy(r->l, t, 48);
}
static inline void q_sub(fq *r, const fq *a, const fq *b) {
    uint64_t t[6];
    if (q_sub6(t, a->l, b->l)) q_add6(t, t, ZKC_Q_64);
    memcpy(r->l, t, 48);
}
static inline void q_mul(fq *r, const fq *a, const fq *b) {   /* CIOS, the algorithm MontBackend::mul_assign implements */
    uint64_t t[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma GCC unroll 6
    for (int i = 0; i < 6; ++i) {
        const uint64_t bi = b->l[i];
        uint64_t carry = 0, t7;
#pragma GCC unroll 6
        for (int j = 0; j < 6; ++j) {
            const u128 p = (u128)a->l[j] * bi + t[j] + carry;   /* < 2^128: no overflow */
            t[j] = (uint64_t)p;
            carry = (uint64_t)(p >> 64);
        }
        u128 s = (u128)t[6] + carry;
        t[6] = (uint64_t)s;
        t7 = (uint64_t)(s >> 64);
        const uint64_t m = t[0] * ZKC_INV64;
        u128 p = (u128)m * ZKC_Q_64[0] + t[0];
        carry = (uint64_t)(p >> 64);
#pragma GCC unroll 5
        for (int j = 1; j < 6; ++j) {
            p = (u128)m * ZKC_Q_64[j] + t[j] + carry;
            t[j - 1] = (uint64_t)p;
            carry = (uint64_t)(p >> 64);
        }
        s = (u128)t[6] + carry;
        t[5] = (uint64_t)s;
        t[6] = t7 + (uint64_t)(s >> 64);
    }
    if (t[6] || q_ge(t)) q_sub6(t, t, ZKC_Q_64);
    memcpy(r->l, t, 48);
}
static inline int q_is_zero(const fq *a) {
    uint64_t o = 0;
    for (int i = 0; i < 6; ++i) o |= a->l[i];
    return o == 0;
}
static inline int q_eq(const fq *a, const fq *b) { return memcmp(a->l, b->l, 48) == 0; }
static inline void q_dbl(fq *r, const fq *a) { q_add(r, a, a); }
static void q_inv(fq *r, const fq *a) {   /* a^(q-2) */
    uint64_t e[6], two[6] = {2, 0, 0, 0, 0, 0};
    q_sub6(e, ZKC_Q_64, two);
    fq acc;
    memcpy(acc.l, ZKC_R_64, 48);
    for (int i = 383; i >= 0; --i) {
        q_mul(&acc, &acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) q_mul(&acc, &acc, a);
    }
    *r = acc;
}

/* ---- G1, Jacobian (x = X/Z^2, y = Y/Z^3) ---- */
static inline void j_set_inf(jac *p) { memset(p, 0, sizeof *p); }
static inline int  j_is_inf(const jac *p) { return q_is_zero(&p->z); }
static void j_double(jac *r, const jac *p) {
    if (j_is_inf(p)) { *r = *p; return; }
    fq A, B, C, D, E, F, t, x3, y3, z3;
    q_mul(&A, &p->x, &p->x);
    q_mul(&B, &p->y, &p->y);
    q_mul(&C, &B, &B);
    q_add(&t, &p->x, &B);
    q_mul(&t, &t, &t);
    q_sub(&t, &t, &A);
    q_sub(&t, &t, &C);
    q_dbl(&D, &t);
    q_dbl(&E, &A);
    q_add(&E, &E, &A);
    q_mul(&F, &E, &E);
    q_sub(&x3, &F, &D);
    q_sub(&x3, &x3, &D);
    q_sub(&t, &D, &x3);
    q_mul(&y3, &E, &t);
    q_dbl(&t, &C); q_dbl(&t, &t); q_dbl(&t, &t);
    q_sub(&y3, &y3, &t);
    q_mul(&z3, &p->y, &p->z);
    q_dbl(&z3, &z3);
    r->x = x3; r->y = y3; r->z = z3;
}
static void j_add(jac *r, const jac *a, const jac *b) {
    if (j_is_inf(a)) { *r = *b; return; }
    if (j_is_inf(b)) { *r = *a; return; }
    fq z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t, x3, y3, z3;
    q_mul(&z1z1, &a->z, &a->z);
    q_mul(&z2z2, &b->z, &b->z);
    q_mul(&u1, &a->x, &z2z2);
    q_mul(&u2, &b->x, &z1z1);
    q_mul(&s1, &a->y, &b->z); q_mul(&s1, &s1, &z2z2);
    q_mul(&s2, &b->y, &a->z); q_mul(&s2, &s2, &z1z1);
    if (q_eq(&u1, &u2)) {
        if (q_eq(&s1, &s2)) { j_double(r, a); return; }
        j_set_inf(r);
        return;
    }
    q_sub(&h, &u2, &u1);
    q_dbl(&i, &h); q_mul(&i, &i, &i);
    q_mul(&j, &h, &i);
    q_sub(&rr, &s2, &s1); q_dbl(&rr, &rr);
    q_mul(&v, &u1, &i);
    q_mul(&x3, &rr, &rr);
    q_sub(&x3, &x3, &j);
    q_sub(&x3, &x3, &v);
    q_sub(&x3, &x3, &v);
    q_sub(&t, &v, &x3);
    q_mul(&y3, &rr, &t);
    q_mul(&t, &s1, &j); q_dbl(&t, &t);
    q_sub(&y3, &y3, &t);
    q_add(&z3, &a->z, &b->z);
    q_mul(&z3, &z3, &z3);
    q_sub(&z3, &z3, &z1z1);
    q_sub(&z3, &z3, &z2z2);
    q_mul(&z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
static void j_from_affine(jac *r, const uint64_t p[12]) {
    memcpy(r->x.l, p, 48);
    memcpy(r->y.l, p + 6, 48);
    if (q_is_zero(&r->x) && q_is_zero(&r->y)) { j_set_inf(r); return; }
    memcpy(r->z.l, ZKC_R_64, 48);
}
static void j_to_affine(uint64_t out[12], const jac *p) {
    if (j_is_inf(p)) { memset(out, 0, 96); return; }
    fq zi, zi2, zi3, x, y;
    q_inv(&zi, &p->z);
    q_mul(&zi2, &zi, &zi);
    q_mul(&zi3, &zi2, &zi);
    q_mul(&x, &p->x, &zi2);
    q_mul(&y, &p->y, &zi3);
    memcpy(out, x.l, 48);
    memcpy(out + 6, y.l, 48);
}
/* `mul_bigint`: MSB-first double-and-add over the canonical integer k (4 limbs) */
static void j_mul(jac *r, const jac *p, const uint64_t k[4]) {
    jac acc;
    j_set_inf(&acc);
    for (int i = 255; i >= 0; --i) {
        j_double(&acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) j_add(&acc, &acc, p);
    }
    *r = acc;
}

/* ---- exported group helpers ---- */
void zko_g1_generator(uint64_t out[12]) {
    memcpy(out, ZKC_GX_MONT_64, 48);
    memcpy(out + 6, ZKC_GY_MONT_64, 48);
}
void zko_fq_from_canonical(const uint64_t in[6], uint64_t out[6]) {
    fq a, r2, r;
    memcpy(a.l, in, 48);
    memcpy(r2.l, ZKC_R2_64, 48);
    q_mul(&r, &a, &r2);
    memcpy(out, r.l, 48);
}
void zko_fq_to_canonical(const uint64_t in[6], uint64_t out[6]) {
    fq a, one = {{1, 0, 0, 0, 0, 0}}, r;
    memcpy(a.l, in, 48);
    q_mul(&r, &a, &one);
    memcpy(out, r.l, 48);
}
void zko_fq_op(int op, const uint64_t a[6], const uint64_t b[6], uint64_t out[6]) {   /* 0 add, 1 sub, 2 mul (Montgomery) */
    fq x, y, r;
    memcpy(x.l, a, 48);
    memcpy(y.l, b, 48);
    if (op == 0) q_add(&r, &x, &y);
    else if (op == 1) q_sub(&r, &x, &y);
    else q_mul(&r, &x, &y);
    memcpy(out, r.l, 48);
}
int zko_g1_is_on_curve(const uint64_t p[12]) {
    fq x, y, l, r, b;
    memcpy(x.l, p, 48);
    memcpy(y.l, p + 6, 48);
    if (q_is_zero(&x) && q_is_zero(&y)) return 1;
    if (q_ge(x.l) || q_ge(y.l)) return 0;
    memcpy(b.l, ZKC_B_MONT_64, 48);
    q_mul(&l, &y, &y);
    q_mul(&r, &x, &x);
    q_mul(&r, &r, &x);
    q_add(&r, &r, &b);
    return q_eq(&l, &r);
}
void zko_g1_add(const uint64_t a[12], const uint64_t b[12], uint64_t out[12]) {
    jac x, y, r;
    j_from_affine(&x, a);
    j_from_affine(&y, b);
    j_add(&r, &x, &y);
    j_to_affine(out, &r);
}
void zko_g1_neg(const uint64_t a[12], uint64_t out[12]) {
    fq y, z = {{0}};
    memcpy(out, a, 48);
    memcpy(y.l, a + 6, 48);
    q_sub(&y, &z, &y);
    memcpy(out + 6, y.l, 48);
}
void zko_g1_mul(const uint64_t p[12], const uint64_t k_canonical[4], uint64_t out[12]) {
    jac x, r;
    j_from_affine(&x, p);
    j_mul(&r, &x, k_canonical);
    j_to_affine(out, &r);
}

/* sum_i vals[i] * points[i]  (the `.map(power.mul_bigint(value.into_bigint())).sum()` of multilinear_kzg.rs:39-43 and :101-108);
 * `mask` selects vals[i & mask] (the blown-up quotient of :183-214 without materialising the copies) */
static void dot_g1(jac *out, const uint64_t *vals, uint64_t mask, const uint64_t *points, uint64_t n) {
    const int fid = ZKO_BLS12_381_FR;
    int nt = zko_get_threads();
    if (nt < 1) nt = 1;
    jac *part = (jac *)calloc((size_t)nt, sizeof(jac));
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int t = 0; t < nt; ++t) {
        jac acc;
        j_set_inf(&acc);
        const uint64_t lo = n * (uint64_t)t / (uint64_t)nt, hi = n * (uint64_t)(t + 1) / (uint64_t)nt;
        for (uint64_t i = lo; i < hi; ++i) {
            uint64_t k[4];
            zko_fe_to_canonical(fid, vals + 4 * (i & mask), k);
            if (!(k[0] | k[1] | k[2] | k[3])) continue;    /* 0 * P = infinity */
            jac p, m;
            j_from_affine(&p, points + 12 * i);
            j_mul(&m, &p, k);
            j_add(&acc, &acc, &m);
        }
        part[t] = acc;
    }
    j_set_inf(out);
    for (int t = 0; t < nt; ++t) j_add(out, out, &part[t]);
    free(part);
}

/* trusted_setup.rs:26-63: lagrange basis over the hypercube (variable 0 = most significant index bit), then basis[i] * G */
int zko_kzg_setup_g1(const uint64_t *taus, uint32_t n, uint64_t *g1_out) {
    const int fid = ZKO_BLS12_381_FR;
    if (n == 0 || n > 30) return -1;                      /* "requires at least one variable" */
    const uint64_t len = 1ull << n;
    uint64_t one[4], *om = (uint64_t *)malloc(32 * (size_t)n);
    zko_fe_from_u64(fid, 1, one);
    for (uint32_t i = 0; i < n; ++i) zko_fe_sub(fid, one, taus + 4 * i, om + 4 * i);
    int nt = zko_get_threads();
    if (nt < 1) nt = 1;
    uint64_t gen[12];
    zko_g1_generator(gen);
#pragma omp parallel for num_threads(nt) schedule(static)
    for (uint64_t index = 0; index < len; ++index) {
        uint64_t e[4], k[4];
        memcpy(e, one, 32);
        for (uint32_t i = 0; i < n; ++i) {
            const int bit = (int)((index >> (n - 1 - i)) & 1);
            zko_fe_mul(fid, e, bit ? taus + 4 * i : om + 4 * i, e);
        }
        zko_fe_to_canonical(fid, e, k);
        zko_g1_mul(gen, k, g1_out + 12 * index);
    }
    free(om);
    return 0;
}

/* multilinear_kzg.rs:25-46 */
int zko_kzg_commit(const uint64_t *vals, uint64_t len, const uint64_t *g1, uint64_t g1_len, uint64_t out[12]) {
    if (len != g1_len) return -1;                          /* "Polynomial evaluation must match g1 length" */
    jac r;
    dot_g1(&r, vals, ~0ull, g1, len);
    j_to_affine(out, &r);
    return 0;
}

/* multilinear_kzg.rs:51-127; eval: 1 element, proofs: nvars points */
int zko_kzg_open(const uint64_t *vals, uint32_t nvars, const uint64_t *g1, uint64_t g1_len, const uint64_t *opening,
                 uint32_t n_opening, uint64_t eval[4], uint64_t *proofs) {
    const int fid = ZKO_BLS12_381_FR;
    if (n_opening != nvars) return -1;     /* "number of polynomial variables must match length of opening values" */
    const uint64_t len = 1ull << nvars;
    if (g1_len != len) return -2;          /* "Opening values must match number of variables from trusted setup" */
    zko_mle_evaluate(fid, vals, len, opening, nvars, eval);
    uint64_t *sub = (uint64_t *)malloc(32 * (size_t)len), *next = (uint64_t *)malloc(32 * (size_t)len);
    uint64_t *quot = (uint64_t *)malloc(32 * (size_t)len);
    for (uint64_t i = 0; i < len; ++i) zko_fe_sub(fid, vals + 4 * i, eval, sub + 4 * i);
    uint64_t cur = len;
    for (uint32_t i = 0; i < nvars; ++i) {
        const uint64_t half = cur / 2;
        for (uint64_t j = 0; j < half; ++j) zko_fe_sub(fid, sub + 4 * (j + half), sub + 4 * j, quot + 4 * j);   /* :166-181 */
        jac pr;
        dot_g1(&pr, quot, half - 1, g1, len);              /* blown up i+1 times == index & (half-1); :93-108 */
        j_to_affine(proofs + 12 * i, &pr);
        zko_mle_partial_evaluate(fid, sub, cur, 0, opening + 4 * i, next);   /* :113-119 */
        uint64_t *t = sub; sub = next; next = t;
        cur = half;
    }
    free(sub); free(next); free(quot);
    return 0;
}

/* multilinear_kzg.rs:132-159 with the pairings replaced by the known trapdoor: C - v G == sum_i (tau_i - r_i) Q_i */
int zko_kzg_verify_trapdoor(const uint64_t *taus, uint32_t n, const uint64_t commitment[12], const uint64_t *opening,
                            const uint64_t eval[4], const uint64_t *proofs) {
    const int fid = ZKO_BLS12_381_FR;
    uint64_t gen[12], k[4], vg[12], lhs[12], rhs[12], t[12], d[4];
    zko_g1_generator(gen);
    zko_fe_to_canonical(fid, eval, k);
    zko_g1_mul(gen, k, vg);
    zko_g1_neg(vg, vg);
    zko_g1_add(commitment, vg, lhs);
    memset(rhs, 0, sizeof rhs);
    for (uint32_t i = 0; i < n; ++i) {
        zko_fe_sub(fid, taus + 4 * i, opening + 4 * i, d);
        zko_fe_to_canonical(fid, d, k);
        zko_g1_mul(proofs + 12 * i, k, t);
        zko_g1_add(rhs, t, rhs);
    }
    return memcmp(lhs, rhs, 96) == 0;
}
