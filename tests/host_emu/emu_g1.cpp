// Host emulation of zk::Fq381 / zk::G1 (fq381.cuh, g1.cuh built with -DZK_HOST_EMU) checked against the C oracle.
// Test-only program (tests/test_host_emu.py builds and runs it); exits non-zero on mismatch.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "g1.cuh"
#include "../../oracle/zkoracle.h"

static uint64_t rng_state = 0x7654321ull;
static uint64_t rnd() { uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

static void rand_fq(uint64_t out[6], int kind) {
    uint64_t c[6];
    for (int i = 0; i < 6; ++i) c[i] = rnd();
    c[5] &= (1ull << 60) - 1;             // < 2^380 < q
    if (kind == 1) memset(c, 0, 48);
    if (kind == 2) { memcpy(c, ZKC_Q_64, 48); c[0] -= 1; memcpy(out, c, 48); return; }   // raw limbs q-1 (largest canonical)
    if (kind == 3) { memset(c, 0, 48); c[0] = 1; }
    zko_fq_from_canonical(c, out);
}
static zk::Fq to_fq(const uint64_t a[6]) { zk::Fq r; memcpy(r.v, a, 48); return r; }
static bool eq(const zk::Fq& a, const uint64_t b[6]) { return memcmp(a.v, b, 48) == 0; }
static zk::G1Affine to_aff(const uint64_t p[12]) { zk::G1Affine r; memcpy(r.x.v, p, 48); memcpy(r.y.v, p + 6, 48); return r; }
static void from_xyzz(uint64_t out[12], const zk::G1Xyzz& p) {
    if (zk::G1::is_inf(p)) { memset(out, 0, 96); return; }
    zk::Fq i3;
    zk::G1::inv(i3, p.zzz);
    zk::G1Affine a;
    zk::G1::to_affine_with_inverse(a, p, i3);
    memcpy(out, a.x.v, 48);
    memcpy(out + 6, a.y.v, 48);
}

int main() {
    typedef zk::Fq381 F;
    int bad = 0;
    for (int it = 0; it < 20000; ++it) {
        int ka = it < 64 ? it % 4 : 0, kb = it < 64 ? (it / 4) % 4 : 0;
        uint64_t a[6], b[6], ref[6];
        rand_fq(a, ka); rand_fq(b, kb);
        zk::Fq A = to_fq(a), B = to_fq(b), R;
        F::add(R, A, B); zko_fq_op(0, a, b, ref); if (!eq(R, ref)) { ++bad; printf("add mismatch\n"); }
        F::sub(R, A, B); zko_fq_op(1, a, b, ref); if (!eq(R, ref)) { ++bad; printf("sub mismatch\n"); }
        F::mul(R, A, B); zko_fq_op(2, a, b, ref); if (!eq(R, ref)) { ++bad; printf("mul mismatch\n"); }
    }
    printf("fq381 %s\n", bad ? "FAILED" : "ok");
    {   // inversion: a * a^-1 == 1
        for (int it = 0; it < 20; ++it) {
            uint64_t a[6];
            rand_fq(a, it == 0 ? 3 : 0);
            zk::Fq A = to_fq(a), I, P;
            zk::G1::inv(I, A);
            F::mul(P, A, I);
            zk::Fq one = F::mont_one();
            if (!F::eq(P, one)) { ++bad; printf("inv mismatch\n"); }
        }
        printf("inv %s\n", bad ? "FAILED" : "ok");
    }
    // group law: sums of multiples of the generator in XYZZ coordinates against the oracle's Jacobian arithmetic
    uint64_t gen[12], pts[8][12];
    zko_g1_generator(gen);
    for (int i = 0; i < 8; ++i) {
        uint64_t k[4] = {rnd(), rnd(), rnd(), rnd() >> 2};
        if (i == 0) { k[0] = 1; k[1] = k[2] = k[3] = 0; }
        zko_g1_mul(gen, k, pts[i]);
    }
    for (int it = 0; it < 300; ++it) {
        zk::G1Xyzz acc = zk::G1::infinity(), acc2 = zk::G1::infinity();
        uint64_t ref[12], ref2[12], got[12];
        memset(ref, 0, 96);
        memset(ref2, 0, 96);
        const int n = 1 + (int)(rnd() % 12);
        for (int j = 0; j < n; ++j) {
            int pi = (int)(rnd() % 8);
            const bool negate = (rnd() & 1) != 0;
            if (it % 7 == 0 && j > 0) pi = (int)(it / 7 % 8);        // force repeats: doubling and P + (-P)
            zk::G1::add_affine(acc, to_aff(pts[pi]), negate);
            uint64_t t[12];
            memcpy(t, pts[pi], 96);
            if (negate) zko_g1_neg(t, t);
            zko_g1_add(ref, t, ref);
            if (j & 1) { zk::G1::add_affine(acc2, to_aff(pts[pi]), negate); zko_g1_add(ref2, t, ref2); }
        }
        from_xyzz(got, acc);
        if (memcmp(got, ref, 96)) { ++bad; printf("add_affine chain mismatch at %d\n", it); }
        zk::G1Xyzz s = acc;
        zk::G1::add(s, acc2);
        uint64_t r3[12];
        zko_g1_add(ref, ref2, r3);
        from_xyzz(got, s);
        if (memcmp(got, r3, 96)) { ++bad; printf("add mismatch at %d\n", it); }
        s = acc;
        zk::G1::add(s, acc);                                          // P + P through the general addition
        zko_g1_add(ref, ref, r3);
        from_xyzz(got, s);
        if (memcmp(got, r3, 96)) { ++bad; printf("add(P,P) mismatch at %d\n", it); }
        zk::G1::dbl(s, acc);
        from_xyzz(got, s);
        if (memcmp(got, r3, 96)) { ++bad; printf("dbl mismatch at %d\n", it); }
    }
    {   // affine infinity inputs are skipped
        zk::G1Xyzz acc = zk::G1::infinity();
        uint64_t zero[12] = {0}, got[12];
        zk::G1::add_affine(acc, to_aff(zero));
        zk::G1::add_affine(acc, to_aff(pts[1]));
        zk::G1::add_affine(acc, to_aff(zero));
        from_xyzz(got, acc);
        if (memcmp(got, pts[1], 96)) { ++bad; printf("infinity handling mismatch\n"); }
    }
    printf("g1 %s\n", bad ? "FAILED" : "ok");
    return bad ? 1 : 0;
}
