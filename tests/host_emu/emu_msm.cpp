// The reduction of the bucket-method multi-scalar multiplication (csrc/kzg.cu) replayed on the CPU with the product's own
// headers: the signed-digit decomposition and the plan (msm_plan.cuh, compiled with -DZK_HOST_EMU), buckets, running sums per
// chunk, bit planes of the chunk index, and the host combine (msm_host.h) -- against sums of the oracle's scalar
// multiplications.  Every window width, the grouped (halving) layout, edge scalars.  Test-only program.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "msm_plan.cuh"
#include "msm_host.h"
#include "../../oracle/zkoracle.h"

using namespace zk;

static uint64_t rng_state = 0xB200B200ull;
static uint64_t rnd() { uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

struct Input {
    std::vector<uint64_t> scalars;   // 4 per entry, Montgomery
    std::vector<uint64_t> points;    // 12 per entry, affine Montgomery
};

static Input make_input(size_t n) {
    Input in;
    in.scalars.resize(4 * n);
    in.points.resize(12 * n);
    uint64_t gen[12];
    zko_g1_generator(gen);
    for (size_t i = 0; i < n; ++i) {
        uint8_t b[32];
        for (int k = 0; k < 32; ++k) b[k] = (uint8_t)rnd();
        if (i == 0) memset(b, 0, 32);                         // scalar 0
        if (i == 1) { memset(b, 0, 32); b[0] = 1; }           // 1
        if (i == 2) memset(b, 0xff, 32);                      // 2^256 - 1 mod r
        zko_fe_from_le_bytes_mod_order(ZKO_BLS12_381_FR, b, 32, &in.scalars[4 * i]);
        if (i == 3) {                                         // r - 1
            uint64_t one[4], z[4] = {0, 0, 0, 0};
            zko_fe_from_u64(ZKO_BLS12_381_FR, 1, one);
            zko_fe_sub(ZKO_BLS12_381_FR, z, one, &in.scalars[4 * i]);
        }
        if (i == 4) zko_fe_from_u64(ZKO_BLS12_381_FR, 1ull << 15, &in.scalars[4 * i]);     // digit boundaries of c = 16
        if (i == 5) zko_fe_from_u64(ZKO_BLS12_381_FR, (1ull << 15) + 1, &in.scalars[4 * i]);
        if (i == 6) zko_fe_from_u64(ZKO_BLS12_381_FR, (1ull << 16) - 1, &in.scalars[4 * i]);
        uint64_t k[4] = {rnd(), rnd(), 0, 0};
        if (i % 7 == 3) memcpy(&in.points[12 * i], &in.points[12 * (i - 1)], 96);          // a repeated point
        else if (i % 11 == 5) memset(&in.points[12 * i], 0, 96);                           // infinity among the points
        else zko_g1_mul(gen, k, &in.points[12 * i]);
    }
    return in;
}

static HG1Affine affine_at(const uint64_t* p) { HG1Affine r; memcpy(&r, p, sizeof r); return r; }

// what kzg.cu's kernels compute, one plane array per group
static std::vector<HG1Affine> emulate(const Input& in, size_t n, const MsmPlan& pl) {
    const size_t WG = (size_t)pl.W * pl.groups, keys = WG * pl.B, nT = pl.B / pl.S;
    uint32_t nb = 0, log_s = 0;
    while ((1u << nb) < nT) ++nb;
    while ((1u << log_s) < pl.S) ++log_s;
    std::vector<HG1Xyzz> buckets(keys, HostG1::infinity());
    for (size_t i = 0; i < n; ++i) {                                        // msm_count / msm_scatter / msm_bucket ...
        Fe s;
        memcpy(s.v, &in.scalars[4 * i], 32);
        uint32_t k[8];
        canonical_scalar(k, s);
        const HG1Xyzz p = HostG1::from_affine(affine_at(&in.points[12 * i]));
        for_each_digit(k, pl, group_of(i, pl), [&](int w, uint32_t b, bool neg) {
            HG1Xyzz& acc = buckets[(size_t)w * pl.B + b];
            acc = HostG1::add(acc, neg ? HostG1::neg(p) : p);
        });
    }
    std::vector<HG1Xyzz> chunk_acc(WG * nT), chunk_run(WG * nT);            // msm_chunk_kernel
    for (size_t t = 0; t < WG * nT; ++t) {
        HG1Xyzz run = HostG1::infinity(), acc = HostG1::infinity();
        for (int i = (int)pl.S - 1; i >= 0; --i) {
            run = HostG1::add(run, buckets[t * pl.S + i]);
            acc = HostG1::add(acc, run);
        }
        chunk_acc[t] = acc;
        chunk_run[t] = run;
    }
    std::vector<HG1Xyzz> planes(WG * (nb + 1), HostG1::infinity());        // msm_plane_kernel
    for (size_t w = 0; w < WG; ++w)
        for (uint32_t p = 0; p <= nb; ++p) {
            HG1Xyzz acc = HostG1::infinity();
            for (size_t t = 0; t < nT; ++t)
                if (p == nb || ((t >> p) & 1)) acc = HostG1::add(acc, (p == nb ? chunk_acc : chunk_run)[w * nT + t]);
            planes[w * (nb + 1) + p] = acc;
        }
    std::vector<HG1Affine> out(pl.groups);
    for (uint32_t g = 0; g < pl.groups; ++g) out[g] = msm_combine_planes(planes.data() + (size_t)g * pl.W * (nb + 1), pl, nb, log_s);
    return out;
}

static void expected(const Input& in, size_t lo, size_t hi, uint64_t out[12]) {
    memset(out, 0, 96);
    for (size_t i = lo; i < hi; ++i) {
        uint64_t k[4], t[12];
        zko_fe_to_canonical(ZKO_BLS12_381_FR, &in.scalars[4 * i], k);
        zko_g1_mul(&in.points[12 * i], k, t);
        zko_g1_add(out, t, out);
    }
}

int main() {
    int bad = 0;
    {   // one plain sum, every window width the library can be asked for
        const size_t n = 45;
        Input in = make_input(n);
        uint64_t want[12];
        expected(in, 0, n, want);
        const int widths[] = {2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
        for (int c : widths) {
            const MsmPlan pl = plan_with_window(n, c);
            const std::vector<HG1Affine> got = emulate(in, n, pl);
            if (memcmp(&got[0], want, 96)) { ++bad; printf("plain sum mismatch at c = %d\n", c); }
        }
        printf("window widths %s\n", bad ? "FAILED" : "ok");
    }
    {   // the library's own plans at the sizes it switches at
        const size_t sizes[] = {1, 2, 3, 255, 256, 300};
        for (size_t n : sizes) {
            Input in = make_input(n);
            uint64_t want[12];
            expected(in, 0, n, want);
            const std::vector<HG1Affine> got = emulate(in, n, plan_for(n));
            if (memcmp(&got[0], want, 96)) { ++bad; printf("plan_for mismatch at n = %zu\n", n); }
        }
        printf("library plans %s\n", bad ? "FAILED" : "ok");
    }
    {   // several sums in one pass: consecutive ranges of halving size, each with its own windows
        const int cs[] = {4, 7, 10};
        for (uint32_t groups = 2; groups <= 6; ++groups) {
            const size_t n = ((size_t)1 << groups) - 1;
            Input in = make_input(n);
            for (int c : cs) {
                const MsmPlan pl = plan_with_window(n, c, groups, groups - 1);
                const std::vector<HG1Affine> got = emulate(in, n, pl);
                size_t lo = 0;
                for (uint32_t g = 0; g < groups; ++g) {
                    const size_t len = (size_t)1 << (groups - 1 - g);
                    uint64_t want[12];
                    expected(in, lo, lo + len, want);
                    if (memcmp(&got[g], want, 96)) { ++bad; printf("grouped sum mismatch: groups = %u, c = %d, group %u\n", groups, c, g); }
                    lo += len;
                }
            }
        }
        printf("grouped sums %s\n", bad ? "FAILED" : "ok");
    }
    return bad ? 1 : 0;
}
