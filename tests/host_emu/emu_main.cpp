// Host emulation of zk::Fp (fp.cuh built with -DZK_HOST_EMU) checked against the C oracle.
// Test-only program (tests/test_host_emu.py builds and runs it); exits non-zero on mismatch.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "fp.cuh"
#include "../../oracle/zkoracle.h"

static uint64_t rng_state = 0x1234567ull;
static uint64_t rnd() { uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

static void rand_fe(int fid, uint64_t out[4], int kind) {
    uint8_t b[32];
    for (int i = 0; i < 32; ++i) b[i] = (uint8_t)rnd();
    if (kind == 1) memset(b, 0, 32);
    if (kind == 2) memset(b, 0xff, 32);
    zko_fe_from_le_bytes_mod_order(fid, b, 32, out);  // a Montgomery-form canonical element
    if (kind == 3) { uint64_t one[4]; zko_fe_from_u64(fid, 1, one); uint64_t z[4] = {0,0,0,0}; zko_fe_sub(fid, z, one, out); } // p-1
    if (kind == 4) zko_fe_from_u64(fid, 1, out);
    if (kind == 5) { uint64_t c[4]; memcpy(c, ZKF_P_64[fid], 32); c[0] -= 1; memcpy(out, c, 32); } // raw limbs p-1 (max canonical repr)
}
static zk::Fe to_fe(const uint64_t a[4]) { zk::Fe r; memcpy(r.v, a, 32); return r; }
static bool eq(const zk::Fe& a, const uint64_t b[4]) { return memcmp(a.v, b, 32) == 0; }

template <int FID> static int run() {
    typedef zk::Fp<FID> P;
    int bad = 0;
    for (int it = 0; it < 20000; ++it) {
        int ka = it < 400 ? (it % 6) : 0, kb = it < 400 ? ((it / 6) % 6) : 0, kc = it < 400 ? ((it / 36) % 6) : 0;
        uint64_t a[4], b[4], c[4], ref[4];
        rand_fe(FID, a, ka); rand_fe(FID, b, kb); rand_fe(FID, c, kc);
        zk::Fe A = to_fe(a), B = to_fe(b), R;
        P::add(R, A, B); zko_fe_add(FID, a, b, ref); if (!eq(R, ref)) { ++bad; printf("add mismatch\n"); }
        P::sub(R, A, B); zko_fe_sub(FID, a, b, ref); if (!eq(R, ref)) { ++bad; printf("sub mismatch\n"); }
        P::mont_mul(R, A, B); zko_fe_mul(FID, a, b, ref); if (!eq(R, ref)) { ++bad; printf("mont_mul mismatch\n"); }
        // lazy accumulation: acc = a*b + c*a + b*c (+ many copies), then redc_wide
        {
            uint32_t acc[17] = {0};
            P::mul_acc(acc, A, B); P::mul_acc(acc, to_fe(c), A); P::mul_acc(acc, B, to_fe(c));
            uint64_t t1[4], t2[4], t3[4], s[4];
            zko_fe_mul(FID, a, b, t1); zko_fe_mul(FID, c, a, t2); zko_fe_mul(FID, b, c, t3);
            zko_fe_add(FID, t1, t2, s); zko_fe_add(FID, s, t3, s);
            if (it % 50 == 0) {  // stress the 17th limb: add a*b 3000 more times
                for (int k = 0; k < 3000; ++k) { P::mul_acc(acc, A, B); zko_fe_add(FID, s, t1, s); }
            }
            P::redc_wide(R, acc);
            if (!eq(R, s)) { ++bad; printf("redc_wide mismatch\n"); }
        }
        // fold by scalar table: out = lo + r*(hi-lo); r = b, lo = a, hi = c
        {
            zk::FoldTable tab;
            uint64_t rplain[4], cur[4];
            zko_fe_to_canonical(FID, b, rplain);          // plain r
            memcpy(cur, rplain, 32);                      // tab[i] = r * 2^(32 i) mod p, as plain integers:
            // multiply by 2^32 mod p == Montgomery-multiply "cur" by the Montgomery form of 2^32
            uint64_t m232[4]; zko_fe_from_u64(FID, 1ull << 32, m232);
            for (int i = 0; i < 8; ++i) { memcpy(tab.w[i], cur, 32); zko_fe_mul(FID, cur, m232, cur); }
            zk::FoldScalar<FID>::fold(R, A, to_fe(c), tab);
            uint64_t d[4], m[4];
            zko_fe_sub(FID, c, a, d); zko_fe_mul(FID, b, d, m); zko_fe_add(FID, a, m, ref);
            if (!eq(R, ref)) { ++bad; printf("fold mismatch it=%d\n", it); }
        }
        // 9-limb plain sums
        {
            uint32_t acc[9] = {0};
            uint64_t s[4] = {0, 0, 0, 0};
            int reps = (it % 100 == 0) ? 5000 : 3;
            for (int k = 0; k < reps; ++k) { P::acc9_add(acc, A); zko_fe_add(FID, s, a, s); P::acc9_add(acc, B); zko_fe_add(FID, s, b, s); }
            P::reduce9(R, acc);
            if (!eq(R, s)) { ++bad; printf("reduce9 mismatch\n"); }
        }
        if (bad > 10) break;
    }
    // barrett on extreme inputs: s = 2^291 - 1 and neighbours of multiples of p
    {
        uint32_t s[10]; for (int k = 0; k < 9; ++k) s[k] = 0xffffffffu; s[9] = 7;
        uint32_t r[8]; P::barrett(r, s);
        // reference via python-free check: (2^291-1) mod p computed with the oracle: from_le_bytes of 37 bytes
        uint8_t bytes[37]; memset(bytes, 0xff, 36); bytes[36] = 7;
        uint64_t m[4], cplain[4]; zko_fe_from_le_bytes_mod_order(FID, bytes, 37, m); zko_fe_to_canonical(FID, m, cplain);
        if (memcmp(r, cplain, 32)) { ++bad; printf("barrett max mismatch\n"); }
    }
    printf("field %d: %s\n", FID, bad ? "FAIL" : "ok");
    return bad;
}

int main() {
    int bad = run<0>() + run<1>() + run<2>();
    return bad ? 1 : 0;
}
