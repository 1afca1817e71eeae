// dev_transcript.cuh -- the reference's Fiat-Shamir transcript on the device.
//
// transcripts/src/fiat_shamir/fiat_shamir_transcript.rs:12-43: a Keccak-256 sponge that is never reset;
// `append` absorbs bytes, `sample_random_challenge` finalises a CLONE of the hasher and then absorbs the
// digest into the live one, `random_challenge_as_field_element` reads the digest little-endian mod p.
// The sponge state (25 lanes + the byte position inside the 136-byte rate block) is handed over from the
// host transcript (host_field.h Keccak256::export_state) when the round loop moves onto the GPU for the
// latency-bound tail of a sumcheck (tail.cuh), and handed back afterwards.
//
// One thread runs the permutation (24 rounds x ~190 32-bit instructions, rolled so the loop stays in the
// instruction cache); everything here is plain integer code and also compiles for the host with
// -DZK_HOST_EMU, where tests/host_emu checks it against the oracle's transcript byte for byte.
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

ZK_DEV uint64_t bswap64(uint64_t x) {
    x = ((x & 0x00ff00ff00ff00ffull) << 8) | ((x >> 8) & 0x00ff00ff00ff00ffull);
    x = ((x & 0x0000ffff0000ffffull) << 16) | ((x >> 16) & 0x0000ffff0000ffffull);
    return (x << 32) | (x >> 32);
}

}  // namespace zk
