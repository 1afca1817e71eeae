// msm_plan.cuh -- the shape of one bucket-method multi-scalar multiplication (window width, windows, buckets, chunking) and the
// scalar -> signed-digit decomposition, shared by the kernels of kzg.cu and, compiled for the host (-DZK_HOST_EMU), by
// tests/host_emu/emu_msm.cpp, which replays the whole reduction (buckets, chunk running sums, bit planes, host combine) on the CPU
// against the oracle.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include "fp.cuh"

namespace zk {

struct MsmPlan {
    int c;            // window width in bits (signed digits in [-2^(c-1), 2^(c-1)])
    int W;            // windows = ceil(256 / c): the top window also takes the last carry (scalars are < 2^255)
    uint32_t B;       // buckets per window = 2^(c-1); bucket b holds the points whose digit is +-(b + 1)
    uint32_t S;       // buckets per running-sum chunk
    uint32_t cap0;    // entries per level-0 segment of the bucket sums
    // several sums in one pass (the small levels of an opening): `groups` consecutive index ranges of halving size, the first
    // 2^n0_log long (2^n0_log, 2^(n0_log-1), .., 1); every group has its own windows.  groups == 1: one plain sum.
    uint32_t groups, n0_log;
};
inline MsmPlan plan_for(uint64_t n, uint32_t groups = 1, uint32_t n0_log = 0) {
    int c = n >= (1u << 20) ? 16 : n >= (1u << 16) ? 13 : n >= (1u << 12) ? 10 : n >= (1u << 8) ? 7 : 4;
    if (const char* e = getenv("ZKB200_MSM_WINDOW")) {
        const int v = atoi(e);
        if (v >= 2 && v <= 16) c = v;
    }
    MsmPlan p;
    p.c = c;
    p.W = (256 + c - 1) / c;
    p.B = 1u << (c - 1);
    // serial depth of the window sums: 2 S additions per chunk, then B / (128 S) + 8 in the block-per-plane reduction
    p.S = p.B >= 16384 ? 8 : p.B >= 2048 ? 4 : p.B >= 256 ? 2 : 1;
    // level-0 segments: long enough to amortise a thread, short enough that a small problem still fills the machine
    const uint64_t entries = n * (uint64_t)p.W;
    p.cap0 = 8;
    while (p.cap0 < 128 && entries / p.cap0 > 65536) p.cap0 <<= 1;
    p.groups = groups;
    p.n0_log = n0_log;
    return p;
}

// the plan for an explicitly chosen window width (tests)
inline MsmPlan plan_with_window(uint64_t n, int c, uint32_t groups = 1, uint32_t n0_log = 0) {
    MsmPlan p = plan_for(n, groups, n0_log);
    p.c = c;
    p.W = (256 + c - 1) / c;
    p.B = 1u << (c - 1);
    p.S = p.B >= 16384 ? 8 : p.B >= 2048 ? 4 : p.B >= 256 ? 2 : 1;
    return p;
}

// ---------------------------------------------------------------- scalars -> signed digits
// `into_bigint()`: the canonical integer of a Montgomery-form scalar
ZK_DEV void canonical_scalar(uint32_t k[8], const Fe& s) {
    Fe one, r;
#pragma unroll
    for (int i = 0; i < 8; ++i) one.v[i] = i == 0 ? 1u : 0u;
    Fp<BLS12_381_FR>::mont_mul(r, s, one);
#pragma unroll
    for (int i = 0; i < 8; ++i) k[i] = r.v[i];
}
ZK_DEV uint32_t window_bits(const uint32_t k[8], int lo, int c) {
    const int word = lo >> 5, sh = lo & 31;
    uint64_t v = k[word];
    if (word + 1 < 8) v |= (uint64_t)k[word + 1] << 32;
    return (uint32_t)(v >> sh) & ((1u << c) - 1u);
}
// the group of index i under the halving layout: group g covers 2^(n0_log - g) indices
ZK_DEV uint32_t group_of(uint64_t i, const MsmPlan& pl) {
    if (pl.groups == 1) return 0;
    const uint64_t r = ((2ull << pl.n0_log) - 1) - i;   // counts down from 2^(n0_log+1) - 1
#if defined(ZK_HOST_EMU)
    return pl.n0_log - (63 - (uint32_t)__builtin_clzll(r));
#else
    return pl.n0_log - (63 - __clzll((long long)r));
#endif
}
// calls f(window, bucket, negative) for every non-zero digit of k; `window` already carries the group's offset
template <typename Fn> ZK_DEV void for_each_digit(const uint32_t k[8], const MsmPlan& pl, uint32_t group, Fn f) {
    uint32_t carry = 0;
    const int w0 = (int)group * pl.W;
    for (int w = 0; w < pl.W; ++w) {
        uint32_t d = window_bits(k, w * pl.c, pl.c) + carry;
        carry = 0;
        bool neg = false;
        if (d > pl.B) {
            d = (1u << pl.c) - d;
            neg = true;
            carry = 1;
        }
        if (d) f(w0 + w, d - 1u, neg);
    }
}


}  // namespace zk
