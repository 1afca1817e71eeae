// dev_transcript.cuh -- the reference's Fiat-Shamir transcript on the device.
//
// transcripts/src/fiat_shamir/fiat_shamir_transcript.rs:12-43: a Keccak-256 sponge that is never reset;
// `append` absorbs bytes, `sample_random_challenge` finalises a CLONE of the hasher and then absorbs the
// digest into the live one, `random_challenge_as_field_element` reads the digest little-endian mod p.
// The sponge state (25 lanes + the byte position inside the 136-byte rate block) is handed over from the
// host transcript (host_field.h Keccak256::export_state) when the round loop moves onto the GPU for the
// latency-bound tail of a sumcheck (tail.cuh), and handed back afterwards.
//
// Two forms of the permutation: one thread (keccak_f1600: the plain restatement, the cross-check) and one warp
// (warp_keccak_f1600: what the tail kernel runs; a single thread needs ~6000 dependent-ish instructions per
// permutation and was 3/4 of a tail round).  Everything here is plain integer code and also compiles for the host
// with -DZK_HOST_EMU, where tests/host_emu checks both forms against the oracle's transcript byte for byte.
#pragma once
#include <stdint.h>
#include "ptx_carry.cuh"   // ZK_DEV

namespace zk {

struct KeccakState {
    uint64_t s[25];
    uint32_t pos;      // bytes absorbed into the current rate block, 0..135
    uint32_t pad_;
};
constexpr uint32_t kKeccakRate = 136;

#define ZK_KECCAK_RC_INIT                                                                                   \
    {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,            \
     0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,            \
     0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,            \
     0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,            \
     0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,            \
     0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull}
#if defined(ZK_HOST_EMU)
static const uint64_t kKeccakRC[24] = ZK_KECCAK_RC_INIT;
#else
static __device__ __constant__ uint64_t kKeccakRC[24] = ZK_KECCAK_RC_INIT;
#endif
#undef ZK_KECCAK_RC_INIT

ZK_DEV uint64_t rol64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

// Keccak-f[1600] on 25 lanes held in registers.
ZK_DEV void keccak_f1600(uint64_t (&a)[25]) {
#pragma unroll 1
    for (int r = 0; r < 24; ++r) {
        uint64_t c0 = a[0] ^ a[5] ^ a[10] ^ a[15] ^ a[20], c1 = a[1] ^ a[6] ^ a[11] ^ a[16] ^ a[21],
                 c2 = a[2] ^ a[7] ^ a[12] ^ a[17] ^ a[22], c3 = a[3] ^ a[8] ^ a[13] ^ a[18] ^ a[23],
                 c4 = a[4] ^ a[9] ^ a[14] ^ a[19] ^ a[24];
        uint64_t d0 = c4 ^ rol64(c1, 1), d1 = c0 ^ rol64(c2, 1), d2 = c1 ^ rol64(c3, 1), d3 = c2 ^ rol64(c4, 1),
                 d4 = c3 ^ rol64(c0, 1);
        uint64_t b[25];   // theta, rho and pi: lane (x, y) moves to (y, 2x + 3y)
        b[0] = a[0] ^ d0;
        b[10] = rol64(a[1] ^ d1, 1);   b[20] = rol64(a[2] ^ d2, 62);  b[5] = rol64(a[3] ^ d3, 28);   b[15] = rol64(a[4] ^ d4, 27);
        b[16] = rol64(a[5] ^ d0, 36);  b[1] = rol64(a[6] ^ d1, 44);   b[11] = rol64(a[7] ^ d2, 6);   b[21] = rol64(a[8] ^ d3, 55);
        b[6] = rol64(a[9] ^ d4, 20);   b[7] = rol64(a[10] ^ d0, 3);   b[17] = rol64(a[11] ^ d1, 10); b[2] = rol64(a[12] ^ d2, 43);
        b[12] = rol64(a[13] ^ d3, 25); b[22] = rol64(a[14] ^ d4, 39); b[23] = rol64(a[15] ^ d0, 41); b[8] = rol64(a[16] ^ d1, 45);
        b[18] = rol64(a[17] ^ d2, 15); b[3] = rol64(a[18] ^ d3, 21);  b[13] = rol64(a[19] ^ d4, 8);  b[14] = rol64(a[20] ^ d0, 18);
        b[24] = rol64(a[21] ^ d1, 2);  b[9] = rol64(a[22] ^ d2, 61);  b[19] = rol64(a[23] ^ d3, 56); b[4] = rol64(a[24] ^ d4, 14);
#pragma unroll
        for (int y = 0; y < 25; y += 5) {   // chi
            a[y + 0] = b[y + 0] ^ (~b[y + 1] & b[y + 2]);
            a[y + 1] = b[y + 1] ^ (~b[y + 2] & b[y + 3]);
            a[y + 2] = b[y + 2] ^ (~b[y + 3] & b[y + 4]);
            a[y + 3] = b[y + 3] ^ (~b[y + 4] & b[y + 0]);
            a[y + 4] = b[y + 4] ^ (~b[y + 0] & b[y + 1]);
        }
        a[0] ^= kKeccakRC[r];   // iota
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Keccak-f[1600] across one warp: lane l = x + 5 y (l < 25) holds state word A[x][y] as two 32-bit halves; a round is
// three shuffle stages -- column parities, the theta correction D[x] = C[x-1] ^ rol(C[x+1], 1), and rho + pi + chi
// fetched straight from the source lanes -- about 45 warp instructions instead of ~250 for one thread.  On the
// device a permutation takes ~1.5 us against ~5.5 us for the single-thread version, which is what bounds a tail
// round (profiles/).  The lane routing lives in two packed words per lane (tools/gen_keccak_lanes.py checks them
// against a textbook Keccak-f and SHA3-256("")):
//   A: the four other lanes of the column (4 x 5 bits), the lanes holding columns x-1 and x+1 (2 x 5 bits)
//   B: rho offset (6 bits), then the SOURCE lanes (before pi) of this lane's b[x], b[x+1], b[x+2] (3 x 5 bits)
// `L` is the lane-value type: uint32_t on the device (one value per thread, wk_shfl = SHFL), a 32-entry array on the
// host (tests/host_emu), so the same source is checked against the scalar permutation without a GPU.
#define ZK_WK_A_INIT                                                                                                   \
    {0x024a3d45u, 0x040ac166u, 0x061b4587u, 0x082bc9a8u, 0x003c4dc9u, 0x0c9051eau, 0x0e50d60bu, 0x10615a2cu, 0x1271de4du, \
     0x0a82626eu, 0x16e2828fu, 0x18a306b0u, 0x1ab38ad1u, 0x1cc40ef2u, 0x14d49313u, 0x21351414u, 0x22f59835u, 0x25061c56u, \
     0x2716a077u, 0x1f272498u, 0x2b87a8a0u, 0x2d482cc1u, 0x2f58b0e2u, 0x31693503u, 0x2979b924u, 0x339ce739u, 0x35ad6b5au, \
     0x37bdef7bu, 0x39ce739cu, 0x3bdef7bdu, 0x3def7bdeu, 0x3fffffffu}
#define ZK_WK_B_INIT                                                                                                   \
    {0x000c3000u, 0x00126181u, 0x0018933eu, 0x0000c49cu, 0x0006061bu, 0x000a48e4u, 0x0010526cu, 0x00168286u, 0x0003b437u, \
     0x00091d94u, 0x000d3843u, 0x001369cau, 0x00149b6bu, 0x0001a4d9u, 0x00070d27u, 0x000b2929u, 0x0011596du, 0x00178acfu, \
     0x0004bc55u, 0x000525c8u, 0x000e4092u, 0x000f7202u, 0x00157bbdu, 0x0002abf8u, 0x0008154eu, 0x0019ce40u, 0x001ad680u, \
     0x001bdec0u, 0x001ce700u, 0x001def40u, 0x001ef780u, 0x001fffc0u}
#if defined(ZK_HOST_EMU)
static const uint32_t kWkA[32] = ZK_WK_A_INIT;
static const uint32_t kWkB[32] = ZK_WK_B_INIT;
#else
static __device__ __constant__ uint32_t kWkA[32] = ZK_WK_A_INIT;
static __device__ __constant__ uint32_t kWkB[32] = ZK_WK_B_INIT;
// device lane ops (the host versions on 32-entry arrays live in tests/host_emu/emu_main.cpp)
ZK_DEV uint32_t wk_shfl(uint32_t v, uint32_t src) { return __shfl_sync(0xffffffffu, v, (int)src); }
ZK_DEV uint32_t wk_funnel_l(uint32_t lo, uint32_t hi, uint32_t n) { return __funnelshift_l(lo, hi, n); }   // (hi:lo << n) >> 32, n in [0, 31]
ZK_DEV uint32_t wk_lane0_mask(uint32_t) { return (threadIdx.x & 31u) == 0u ? 0xffffffffu : 0u; }
#endif
#undef ZK_WK_A_INIT
#undef ZK_WK_B_INIT

template <class L> ZK_DEV void warp_keccak_f1600(L& lo, L& hi, const L A, const L B) {
    const L s1 = A & 31u, s2 = (A >> 5) & 31u, s3 = (A >> 10) & 31u, s4 = (A >> 15) & 31u, dm1 = (A >> 20) & 31u, dp1 = (A >> 25) & 31u;
    const L p0 = (B >> 6) & 31u, p1 = (B >> 11) & 31u, p2 = (B >> 16) & 31u;
    const L rot = B & 31u;                        // rho mod 32
    const L swap = ~(((B >> 5) & 1u) + 0xffffffffu);   // all ones where rho >= 32: the halves trade places first
    const L lane0 = wk_lane0_mask(lo);
#if defined(ZK_HOST_EMU)
#pragma unroll 1
#else
#pragma unroll   // 24 rounds inline: the round constants become immediates and the scheduler overlaps a round's tail with the next one's head
#endif
    for (int r = 0; r < 24; ++r) {
        // theta: column parity, then D
        L cl = lo ^ wk_shfl(lo, s1) ^ wk_shfl(lo, s2) ^ wk_shfl(lo, s3) ^ wk_shfl(lo, s4);
        L ch = hi ^ wk_shfl(hi, s1) ^ wk_shfl(hi, s2) ^ wk_shfl(hi, s3) ^ wk_shfl(hi, s4);
        L nl = wk_shfl(cl, dp1), nh = wk_shfl(ch, dp1);          // C[x+1]
        L tl = lo ^ wk_shfl(cl, dm1) ^ wk_funnel_l(nh, nl, 1u);  // low half of rol(C[x+1], 1) = (nl << 1) | (nh >> 31)
        L th = hi ^ wk_shfl(ch, dm1) ^ wk_funnel_l(nl, nh, 1u);
        // rho at the source lane
        L al = (tl & ~swap) | (th & swap), ah = (th & ~swap) | (tl & swap);
        L rl = wk_funnel_l(ah, al, rot), rh = wk_funnel_l(al, ah, rot);
        // pi + chi: fetch b[x], b[x+1], b[x+2] of this lane's row from where they were before pi
        L b0l = wk_shfl(rl, p0), b0h = wk_shfl(rh, p0);
        L b1l = wk_shfl(rl, p1), b1h = wk_shfl(rh, p1);
        L b2l = wk_shfl(rl, p2), b2h = wk_shfl(rh, p2);
        lo = b0l ^ (~b1l & b2l);
        hi = b0h ^ (~b1h & b2h);
        // iota
        const uint64_t rc = kKeccakRC[r];
        lo = lo ^ (lane0 & (uint32_t)rc);
        hi = hi ^ (lane0 & (uint32_t)(rc >> 32));
    }
}

// permute a state that lives in (shared / global / host) memory
ZK_DEV void keccak_permute_mem(uint64_t* s) {
    uint64_t a[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) a[i] = s[i];
    keccak_f1600(a);
#pragma unroll
    for (int i = 0; i < 25; ++i) s[i] = a[i];
}

ZK_DEV void sponge_absorb_byte(KeccakState* st, uint32_t byte) {
    st->s[st->pos >> 3] ^= (uint64_t)(byte & 0xffu) << (8 * (st->pos & 7));
    if (++st->pos == kKeccakRate) {
        keccak_permute_mem(st->s);
        st->pos = 0;
    }
}
// absorb the 8 bytes of w, least significant byte first (`hasher.update`, fiat_shamir_transcript.rs:22-24)
ZK_DEV void sponge_absorb_word(KeccakState* st, uint64_t w) {
    if ((st->pos & 7) == 0) {   // the rate is a multiple of 8: an aligned word never straddles a block
        st->s[st->pos >> 3] ^= w;
        st->pos += 8;
        if (st->pos == kKeccakRate) {
            keccak_permute_mem(st->s);
            st->pos = 0;
        }
    } else {
        for (int b = 0; b < 8; ++b) sponge_absorb_byte(st, (uint32_t)(w >> (8 * b)));
    }
}
// sample_random_challenge (fiat_shamir_transcript.rs:29-36): digest = finalize(clone); live.update(digest).
// The digest comes back as 4 little-endian words -- already the integer `from_le_bytes_mod_order` reads (:42).
ZK_DEV void sponge_sample(KeccakState* st, uint64_t digest[4]) {
    uint64_t a[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) a[i] = st->s[i];
    const uint32_t pos = st->pos;
    const uint64_t pad = 0x01ull << (8 * (pos & 7));   // original Keccak padding 0x01 .. 0x80
    const uint32_t w = pos >> 3;
#pragma unroll
    for (int i = 0; i < 17; ++i) a[i] ^= (i == (int)w) ? pad : 0ull;
    a[16] ^= 0x8000000000000000ull;
    keccak_f1600(a);
#pragma unroll
    for (int i = 0; i < 4; ++i) digest[i] = a[i];
    for (int i = 0; i < 4; ++i) sponge_absorb_word(st, digest[i]);
}

// ---------------------------------------------------------------------------------------------------------------
// The transcript as ONE WARP runs it (tail.cuh).  `Exec` supplies lane(), permute(s) and finalize(s, pos, digest):
// on the device the permutation is warp_keccak_f1600 over shuffles, lane 0 does the byte bookkeeping, and `pos` is a
// warp-uniform register.  `s` (25 words) and `digest` (4 words) live in memory the whole warp sees.
template <class Exec> ZK_DEV void coop_absorb_words(Exec& ex, uint64_t* s, uint32_t& pos, const uint64_t* words, int n) {
    for (int i = 0; i < n; ++i) {
        const uint64_t w = words[i];
        if ((pos & 7) == 0) {   // the rate is a multiple of 8: an aligned word never straddles a block
            if (ex.lane() == 0) s[pos >> 3] ^= w;
            pos += 8;
            if (pos == kKeccakRate) {
                ex.permute(s);
                pos = 0;
            }
        } else {
            for (int b = 0; b < 8; ++b) {
                if (ex.lane() == 0) s[pos >> 3] ^= ((w >> (8 * b)) & 0xffull) << (8 * (pos & 7));
                if (++pos == kKeccakRate) {
                    ex.permute(s);
                    pos = 0;
                }
            }
        }
    }
}
// sample_random_challenge (fiat_shamir_transcript.rs:29-36): digest = finalize(clone); live.update(digest)
template <class Exec> ZK_DEV void coop_sample(Exec& ex, uint64_t* s, uint32_t& pos, uint64_t* digest) {
    ex.finalize(s, pos, digest);
    coop_absorb_words(ex, s, pos, digest, 4);
}

ZK_DEV uint64_t bswap64(uint64_t x) {
    x = ((x & 0x00ff00ff00ff00ffull) << 8) | ((x >> 8) & 0x00ff00ff00ff00ffull);
    x = ((x & 0x0000ffff0000ffffull) << 16) | ((x >> 16) & 0x0000ffff0000ffffull);
    return (x << 32) | (x >> 32);
}

}  // namespace zk
