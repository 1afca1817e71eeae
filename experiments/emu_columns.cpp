// experiments/emu_columns.cpp -- host emulation check of experiments/fp_columns.cuh against the C oracle
// (moved out of tests/host_emu/emu_main.cpp with the experiment code).  Build and run (not part of the test suite):
//   g++ -O1 -std=c++17 -DZK_HOST_EMU -x c++ -I zk_cryptography_research_implementations_b200/csrc \
//       experiments/emu_columns.cpp oracle/zkoracle.c -o /tmp/emu_columns && /tmp/emu_columns
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "fp_columns.cuh"
#include "fp_karatsuba.cuh"
#include "../oracle/zkoracle.h"

static uint64_t rng_state = 0x1234567ull;
static uint64_t rnd() { uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static void rand_fe(int fid, uint64_t out[4]) {
    uint8_t b[32];
    for (int i = 0; i < 32; ++i) b[i] = (uint8_t)rnd();
    zko_fe_from_le_bytes_mod_order(fid, b, 32, out);
}
static zk::Fe to_fe(const uint64_t a[4]) { zk::Fe r; memcpy(r.v, a, 32); return r; }
static bool eq(const zk::Fe& a, const uint64_t b[4]) { return memcmp(a.v, b, 32) == 0; }

template <int FID> static int run() {
    typedef zk::Fp<FID> P;
    typedef zk::FpColumns<FID> C;
    int bad = 0;
    for (int it = 0; it < 5000; ++it) {
        uint64_t a[4], b[4], c[4], ref[4];
        rand_fe(FID, a); rand_fe(FID, b); rand_fe(FID, c);
        zk::Fe A = to_fe(a), B = to_fe(b), R;
        // the same sums through the carry-chain-free column accumulator
        {
            typename C::ColAcc ca; C::cols_init(ca);
            C::mul_acc_cols(ca, A, B); C::mul_acc_cols(ca, to_fe(c), A); C::mul_acc_cols(ca, B, to_fe(c));
            uint64_t t1[4], t2[4], t3[4], s[4];
            zko_fe_mul(FID, a, b, t1); zko_fe_mul(FID, c, a, t2); zko_fe_mul(FID, b, c, t3);
            zko_fe_add(FID, t1, t2, s); zko_fe_add(FID, s, t3, s);
            if (it % 50 == 0) {
                for (int k = 0; k < 3000; ++k) { C::mul_acc_cols(ca, A, B); zko_fe_add(FID, s, t1, s); }
            }
            uint32_t limbs[17]; C::cols_to_limbs(limbs, ca);
            P::redc_wide(R, limbs);
            if (!eq(R, s)) { ++bad; printf("mul_acc_cols mismatch\n"); }
            // and limb for limb against the chained accumulator
            uint32_t acc[17] = {0};
            P::mul_acc(acc, A, B); P::mul_acc(acc, to_fe(c), A); P::mul_acc(acc, B, to_fe(c));
            if (it % 50 == 0) for (int k = 0; k < 3000; ++k) P::mul_acc(acc, A, B);
            if (memcmp(acc, limbs, sizeof acc)) { ++bad; printf("column accumulator differs from the chained accumulator\n"); }
        }
        // the same sums through flag-free radix-2^29 columns (six multiplications per flush)
        {
            typename C::Cols29 cc; C::cols29_init(cc);
            uint32_t acc29[17] = {0}, accw[17] = {0};
            int pending = 0;
            int reps = (it % 50 == 0) ? 500 : 1;
            for (int rep = 0; rep < reps; ++rep) {
                const zk::Fe* xs[3] = {&A, &B, &A};
                zk::Fe Cc = to_fe(c);
                const zk::Fe* ys[3] = {&B, &Cc, &Cc};
                for (int q = 0; q < 3; ++q) {
                    typename C::Digits29 da, db;
                    C::to_digits29(da, *xs[q]); C::to_digits29(db, *ys[q]);
                    C::mul_cols29(cc, da, db);
                    if (++pending == C::kCols29Budget) { C::cols29_flush(acc29, cc); pending = 0; }
                    P::mul_acc(accw, *xs[q], *ys[q]);
                }
            }
            C::cols29_flush(acc29, cc);
            if (memcmp(acc29, accw, sizeof accw)) { ++bad; printf("radix-2^29 accumulator differs from the chained accumulator (it=%d)\n", it); }
        }
        // one-level Karatsuba product against the schoolbook one, limb for limb (edge operands included above: 0, 1, p-1, all ones)
        {
            uint32_t t1[16], t2[16];
            P::mul_wide(t1, A, B);
            zk::FpKaratsuba<FID>::mul_wide_karatsuba(t2, A, B);
            if (memcmp(t1, t2, sizeof t1)) { ++bad; printf("karatsuba product differs from the schoolbook product \n"); }
            zk::Fe X = A, Y = to_fe(c);          // halves that make a1 - a0 / b1 - b0 negative, zero, extreme
            for (int k = 0; k < 4; ++k) { X.v[k + 4] = A.v[k]; Y.v[k] = 0xffffffffu; }
            if (it & 1) for (int k = 0; k < 4; ++k) X.v[k] = 0xffffffffu;
            P::mul_wide(t1, X, Y);
            zk::FpKaratsuba<FID>::mul_wide_karatsuba(t2, X, Y);
            if (memcmp(t1, t2, sizeof t1)) { ++bad; printf("karatsuba product differs on the edge halves \n"); }
        }
        {   // column form of the fold: out = a + b (c - a)
            zk::FoldTable tab;
            uint64_t cur[4], m232[4];
            zko_fe_to_canonical(FID, b, cur);
            zko_fe_from_u64(FID, 1ull << 32, m232);
            for (int i = 0; i < 8; ++i) { memcpy(tab.w[i], cur, 32); zko_fe_mul(FID, cur, m232, cur); }
            C::fold_cols(R, A, to_fe(c), tab);
            uint64_t d[4], m[4];
            zko_fe_sub(FID, c, a, d); zko_fe_mul(FID, b, d, m); zko_fe_add(FID, a, m, ref);
            if (!eq(R, ref)) { ++bad; printf("fold_cols mismatch\n"); }
        }
        if (bad > 10) break;
    }
    printf("columns field %d: %s\n", FID, bad ? "FAIL" : "ok");
    return bad;
}
int main() { return (run<0>() + run<1>() + run<2>()) ? 1 : 0; }
