// host_field.h -- host-side (CPU) prime-field arithmetic and Keccak-256 transcript of the library.
//
// This is the part of the reference's flow that stays on the host: per round the prover absorbs
// d+1 field elements and squeezes one challenge (transcripts/src/fiat_shamir/fiat_shamir_transcript.rs:12-43,
// sumcheck_gkr_protocol.rs:46-55, prover.rs:51-58).  It is a few hundred nanoseconds of work per
// round; the tables never come back to the host.
//
// Independent of oracle/ (the product must not use the checker).
#pragma once
#include <stdint.h>
#include <string.h>
#include <vector>
#include "field_consts.h"

namespace zk {

typedef unsigned __int128 u128;

struct HFe {
    uint64_t l[4];
    bool operator==(const HFe& o) const { return memcmp(l, o.l, 32) == 0; }
    bool operator!=(const HFe& o) const { return !(*this == o); }
};

// 4 x 64-bit-limb Montgomery field, R = 2^256 (arkworks' MontBackend layout).
class HostField {
  public:
    explicit HostField(int fid) : fid_(fid) {
        memcpy(p_, ZKF_P_64[fid], 32);
        inv_ = ZKF_INV64[fid];
        memcpy(one_.l, ZKF_R_64[fid], 32);
        memcpy(r2_.l, ZKF_R2_64[fid], 32);
        memcpy(two_inv_.l, ZKF_TWO_INV_64[fid], 32);
    }
    int fid() const { return fid_; }
    HFe zero() const { return HFe{{0, 0, 0, 0}}; }
    HFe one() const { return one_; }
    HFe two_inv() const { return two_inv_; }
    const uint64_t* modulus() const { return p_; }

    HFe add(const HFe& a, const HFe& b) const {
        HFe r;
        u128 c = 0;
        for (int i = 0; i < 4; ++i) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
        if (c || geq_p(r.l)) sub_p(r.l);
        return r;
    }
    HFe sub(const HFe& a, const HFe& b) const {
        HFe r;
        uint64_t borrow = 0;
        for (int i = 0; i < 4; ++i) {
            u128 d = (u128)a.l[i] - b.l[i] - borrow;
            r.l[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
        if (borrow) {
            u128 c = 0;
            for (int i = 0; i < 4; ++i) { c += (u128)r.l[i] + p_[i]; r.l[i] = (uint64_t)c; c >>= 64; }
        }
        return r;
    }
    HFe neg(const HFe& a) const { return sub(zero(), a); }
    // Montgomery product a*b/R, separated operand scanning (multiply, then reduce word by word)
    HFe mul(const HFe& a, const HFe& b) const {
        uint64_t t[9] = {0};
        for (int i = 0; i < 4; ++i) {
            u128 c = 0;
            for (int j = 0; j < 4; ++j) {
                c += (u128)a.l[i] * b.l[j] + t[i + j];
                t[i + j] = (uint64_t)c;
                c >>= 64;
            }
            t[i + 4] = (uint64_t)c;
        }
        for (int i = 0; i < 4; ++i) {
            uint64_t m = t[i] * inv_;
            u128 c = 0;
            for (int j = 0; j < 4; ++j) {
                c += (u128)m * p_[j] + t[i + j];
                t[i + j] = (uint64_t)c;
                c >>= 64;
            }
            for (int k = i + 4; k < 9 && c; ++k) { c += t[k]; t[k] = (uint64_t)c; c >>= 64; }
        }
        HFe r;
        memcpy(r.l, t + 4, 32);
        if (t[8] || geq_p(r.l)) sub_p(r.l);
        return r;
    }
    HFe from_u64(uint64_t v) const { return to_mont(HFe{{v, 0, 0, 0}}); }
    HFe to_mont(const HFe& plain) const { return mul(plain, r2_); }
    HFe from_mont(const HFe& m) const { return mul(m, HFe{{1, 0, 0, 0}}); }
    HFe pow(const HFe& a, const uint64_t e[4]) const {
        HFe acc = one_;
        for (int i = 255; i >= 0; --i) {
            acc = mul(acc, acc);
            if ((e[i / 64] >> (i % 64)) & 1) acc = mul(acc, a);
        }
        return acc;
    }
    HFe inv(const HFe& a) const {
        uint64_t e[4];
        memcpy(e, p_, 32);
        e[0] -= 2;  // p is odd and > 2: no borrow
        return pow(a, e);
    }
    // `from_le_bytes_mod_order` of a 32-byte digest (fiat_shamir_transcript.rs:42)
    HFe from_le_bytes32_mod_order(const uint8_t d[32]) const {
        HFe v;
        memcpy(v.l, d, 32);  // little-endian host
        // v < 2^256 < 6p (BN254) / 3p (BLS12-381 Fr): subtract p until canonical
        while (geq_p(v.l)) sub_p(v.l);
        return to_mont(v);
    }
    // `into_bigint().to_bytes_be()` / `to_bytes_le()`
    void to_bytes_be(const HFe& m, uint8_t out[32]) const {
        HFe c = from_mont(m);
        for (int i = 0; i < 32; ++i) out[i] = (uint8_t)(c.l[3 - i / 8] >> (56 - 8 * (i % 8)));
    }
    void to_bytes_le(const HFe& m, uint8_t out[32]) const {
        HFe c = from_mont(m);
        memcpy(out, c.l, 32);
    }
    // Horner evaluation of a coefficient-form polynomial (what the verifier does with a round polynomial)
    HFe horner(const HFe* coeffs, int n, const HFe& x) const {
        HFe acc = zero();
        for (int i = n - 1; i >= 0; --i) acc = add(mul(acc, x), coeffs[i]);
        return acc;
    }

  private:
    bool geq_p(const uint64_t a[4]) const {
        for (int i = 3; i >= 0; --i) {
            if (a[i] > p_[i]) return true;
            if (a[i] < p_[i]) return false;
        }
        return true;
    }
    void sub_p(uint64_t a[4]) const {
        uint64_t borrow = 0;
        for (int i = 0; i < 4; ++i) {
            u128 d = (u128)a[i] - p_[i] - borrow;
            a[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
    }
    int fid_;
    uint64_t p_[4];
    uint64_t inv_;
    HFe one_, r2_, two_inv_;
};

// Coefficients of the degree-d polynomial through (0,y0)..(d,yd).  The reference runs a generic
// Lagrange interpolation every round (dense_univariate.rs:74-127, d+1 field inversions); the
// coefficients are a fixed linear map of the evaluations, so the inverse Vandermonde matrix on the
// nodes 0..d is computed once per degree and applied per round.  Same field elements, same limbs.
class Interpolator {
  public:
    Interpolator(const HostField& f, int degree) : f_(f), n_(degree + 1), m_((size_t)n_ * n_) {
        // column k of the matrix = coefficients of the k-th Lagrange basis polynomial on nodes 0..d
        for (int k = 0; k < n_; ++k) {
            std::vector<HFe> num(1, f.one());
            HFe den = f.one();
            for (int j = 0; j < n_; ++j) {
                if (j == k) continue;
                HFe xj = f.from_u64((uint64_t)j), negx = f.neg(xj);
                std::vector<HFe> nxt(num.size() + 1, f.zero());
                for (size_t i = 0; i < num.size(); ++i) {
                    nxt[i] = f.add(nxt[i], f.mul(num[i], negx));
                    nxt[i + 1] = f.add(nxt[i + 1], num[i]);
                }
                num.swap(nxt);
                den = f.mul(den, f.sub(f.from_u64((uint64_t)k), xj));
            }
            HFe dinv = f.inv(den);
            for (int i = 0; i < n_; ++i) m_[(size_t)i * n_ + k] = f.mul(num[i], dinv);
        }
    }
    int degree() const { return n_ - 1; }
    const HFe* matrix() const { return m_.data(); }   // row-major (degree+1)^2, Montgomery form
    void coefficients(const HFe* evals, HFe* coeffs) const {
        for (int i = 0; i < n_; ++i) {
            HFe acc = f_.zero();
            for (int k = 0; k < n_; ++k) acc = f_.add(acc, f_.mul(m_[(size_t)i * n_ + k], evals[k]));
            coeffs[i] = acc;
        }
    }

  private:
    const HostField& f_;
    int n_;
    std::vector<HFe> m_;
};

// ---------------------------------------------------------------- Keccak-256 sponge
class Keccak256 {
  public:
    Keccak256() { reset(); }
    void reset() { memset(s_, 0, sizeof s_); pos_ = 0; }
    void update(const uint8_t* d, size_t len) {
        uint8_t* sb = reinterpret_cast<uint8_t*>(s_);
        while (len) {
            size_t take = kRate - pos_;
            if (take > len) take = len;
            if (pos_ == 0 && take == kRate) {
                const uint64_t* w = reinterpret_cast<const uint64_t*>(d);
                if ((reinterpret_cast<uintptr_t>(d) & 7) == 0) {
                    for (size_t i = 0; i < kRate / 8; ++i) s_[i] ^= w[i];
                } else {
                    for (size_t i = 0; i < kRate; ++i) sb[i] ^= d[i];
                }
                permute();
            } else {
                for (size_t i = 0; i < take; ++i) sb[pos_ + i] ^= d[i];
                pos_ += take;
                if (pos_ == kRate) { permute(); pos_ = 0; }
            }
            d += take;
            len -= take;
        }
    }
    // digest of the data absorbed so far; the live state is left untouched (clone-and-finalize)
    void peek_digest(uint8_t out[32]) const {
        Keccak256 c = *this;
        uint8_t* sb = reinterpret_cast<uint8_t*>(c.s_);
        sb[c.pos_] ^= 0x01;
        sb[kRate - 1] ^= 0x80;
        c.permute();
        memcpy(out, c.s_, 32);
    }
    // hand the sponge over to / take it back from the device transcript (dev_transcript.cuh KeccakState)
    void export_state(uint64_t s[25], uint32_t* pos) const { memcpy(s, s_, sizeof s_); *pos = (uint32_t)pos_; }
    void import_state(const uint64_t s[25], uint32_t pos) { memcpy(s_, s, sizeof s_); pos_ = pos; }

  private:
    static constexpr size_t kRate = 136;
    static inline uint64_t rol(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }
    void permute() {
        static const uint64_t RC[24] = {
            0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
            0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
            0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
            0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
            0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
            0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
        uint64_t* a = s_;
        for (int r = 0; r < 24; ++r) {
            uint64_t c0 = a[0] ^ a[5] ^ a[10] ^ a[15] ^ a[20], c1 = a[1] ^ a[6] ^ a[11] ^ a[16] ^ a[21],
                     c2 = a[2] ^ a[7] ^ a[12] ^ a[17] ^ a[22], c3 = a[3] ^ a[8] ^ a[13] ^ a[18] ^ a[23],
                     c4 = a[4] ^ a[9] ^ a[14] ^ a[19] ^ a[24];
            uint64_t d0 = c4 ^ rol(c1, 1), d1 = c0 ^ rol(c2, 1), d2 = c1 ^ rol(c3, 1), d3 = c2 ^ rol(c4, 1),
                     d4 = c3 ^ rol(c0, 1);
            // theta + rho + pi into b (lane (x,y) -> (y, 2x+3y))
            uint64_t b[25];
            b[0] = a[0] ^ d0;
            b[10] = rol(a[1] ^ d1, 1);   b[20] = rol(a[2] ^ d2, 62);  b[5] = rol(a[3] ^ d3, 28);   b[15] = rol(a[4] ^ d4, 27);
            b[16] = rol(a[5] ^ d0, 36);  b[1] = rol(a[6] ^ d1, 44);   b[11] = rol(a[7] ^ d2, 6);   b[21] = rol(a[8] ^ d3, 55);
            b[6] = rol(a[9] ^ d4, 20);   b[7] = rol(a[10] ^ d0, 3);   b[17] = rol(a[11] ^ d1, 10); b[2] = rol(a[12] ^ d2, 43);
            b[12] = rol(a[13] ^ d3, 25); b[22] = rol(a[14] ^ d4, 39); b[23] = rol(a[15] ^ d0, 41); b[8] = rol(a[16] ^ d1, 45);
            b[18] = rol(a[17] ^ d2, 15); b[3] = rol(a[18] ^ d3, 21);  b[13] = rol(a[19] ^ d4, 8);  b[14] = rol(a[20] ^ d0, 18);
            b[24] = rol(a[21] ^ d1, 2);  b[9] = rol(a[22] ^ d2, 61);  b[19] = rol(a[23] ^ d3, 56); b[4] = rol(a[24] ^ d4, 14);
            for (int y = 0; y < 25; y += 5) {
                a[y + 0] = b[y + 0] ^ (~b[y + 1] & b[y + 2]);
                a[y + 1] = b[y + 1] ^ (~b[y + 2] & b[y + 3]);
                a[y + 2] = b[y + 2] ^ (~b[y + 3] & b[y + 4]);
                a[y + 3] = b[y + 3] ^ (~b[y + 4] & b[y + 0]);
                a[y + 4] = b[y + 4] ^ (~b[y + 0] & b[y + 1]);
            }
            a[0] ^= RC[r];
        }
    }
    uint64_t s_[25];
    size_t pos_;
};

// The reference's `Transcript` (fiat_shamir_transcript.rs:5-43): append / sample / challenge.
class HostTranscript {
  public:
    void append(const uint8_t* d, size_t len) { h_.update(d, len); }                 // :22-24
    void sample(uint8_t out[32]) {                                                   // :29-36
        h_.peek_digest(out);
        h_.update(out, 32);
    }
    HFe challenge(const HostField& f) {                                              // :38-43
        uint8_t d[32];
        sample(d);
        return f.from_le_bytes32_mod_order(d);
    }
    void export_state(uint64_t s[25], uint32_t* pos) const { h_.export_state(s, pos); }
    void import_state(const uint64_t s[25], uint32_t pos) { h_.import_state(s, pos); }
    void append_be(const HostField& f, const HFe& x) { uint8_t b[32]; f.to_bytes_be(x, b); append(b, 32); }
    void append_le(const HostField& f, const HFe& x) { uint8_t b[32]; f.to_bytes_le(x, b); append(b, 32); }

  private:
    Keccak256 h_;
};

}  // namespace zk
