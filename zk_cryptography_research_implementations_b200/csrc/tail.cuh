// tail.cuh -- the latency-bound tail of a sumcheck, run entirely on the GPU: one launch, no host round trips.
//
// Once the tables have shrunk to a few thousand entries a round is a microsecond of arithmetic wrapped in a
// kernel launch, a PCIe mailbox write, a host Keccak and the next launch.  From that point on ONE block runs
// all remaining rounds: fold by the previous challenge + round sums (the same RoundAcc / FoldScalar code as the
// big kernels), exact column sums over the block, the d+1 evaluations, Lagrange coefficients
// (dense_univariate.rs:74-127 as a fixed matrix), the transcript absorb and challenge
// (sumcheck_gkr_protocol.rs:46-55 / prover.rs:51-58, fiat_shamir_transcript.rs:29-43 -- dev_transcript.cuh), the
// next fold table, and finally the last partial_evaluate (sumcheck_gkr_protocol.rs:57 / prover.rs:61-63).  The
// sponge state enters with the launch and leaves with the results, so the host transcript carries on seamlessly
// (the GKR prover keeps absorbing after every layer sumcheck).
//
// The body is written against an `Exec` policy (thread id, barrier, block-wide column sums, 256-bit memory
// access, the warp-wide Keccak permutation).  The kernel instantiates it with the CUDA policy; tests/host_emu instantiates it with a one-thread
// host policy and checks whole tails against the oracle's provers without a GPU.
#pragma once
#include "dev_transcript.cuh"
#include "round_acc.cuh"

namespace zk {

constexpr int kTailMaxLog = 16;   // a tail starts on tables of at most 2^16 entries: at most 16 rounds
enum TailMode { kTailProduct = 0, kTailPlain = 1 };

struct TailOut {                               // mapped pinned host memory
    Fe round_vals[kTailMaxLog][kMaxEvals];     // product: the d+1 coefficients per round; plain: [sum left, sum right]
    Fe challenges[kTailMaxLog];                // Montgomery form, as the host provers report them
    Fe finals[kMaxTables];                     // the tables' single entries after the last fold
    KeccakState sponge;                        // transcript state after the last round
    uint32_t rounds;
    uint32_t seq;                              // written last, after a system-scope fence
};

struct TailArgs {
    TablePtrs tp;
    uint32_t log_len;      // every table holds 2^log_len entries on entry
    uint32_t pending;      // 1: the tables still have to be folded by `ft` (entered from the host round loop)
    uint32_t mode;         // TailMode
    uint32_t seq;          // value to publish in out->seq
    FoldTable ft;          // fold table of the pending challenge
    Fe interp[kMaxEvals * kMaxEvals];   // inverse Vandermonde on the nodes 0..D, row-major, Montgomery form
    Fe pow32[8];           // Montgomery forms of 2^(32 i): (plain r) x pow32[i] = r 2^(32 i) mod p, the fold table rows
    KeccakState sponge;    // transcript state on entry
    TailOut* out;
};

struct TailShared {
    FoldTable ft;
    KeccakState sponge;
    Fe evals[kMaxEvals];
    uint64_t words[kMaxEvals][4];   // what the transcript absorbs this round, as little-endian words of the byte stream
    uint64_t digest[4];
    Fe r_plain;
    unsigned long long tot[kMaxCols];
};

template <int FID> struct TailField {
    typedef Fp<FID> P;
    // canonical plain integer of a Montgomery element (`into_bigint`)
    ZK_DEV static void from_mont(uint32_t out[8], const Fe& x) {
        P::redc256(out, x.v);
        P::cond_sub_p(out);
    }
    // `to_bytes_le` as 4 words of the byte stream (sumcheck_gkr_protocol.rs:145-150)
    ZK_DEV static void le_words(uint64_t w[4], const Fe& x) {
        uint32_t c[8];
        from_mont(c, x);
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = (uint64_t)c[2 * i] | ((uint64_t)c[2 * i + 1] << 32);
    }
    // `to_bytes_be` (prover.rs:91-93): the most significant byte first
    ZK_DEV static void be_words(uint64_t w[4], const Fe& x) {
        uint64_t le[4];
        le_words(le, x);
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = bswap64(le[3 - i]);
    }
    // `from_le_bytes_mod_order` of a 32-byte digest (fiat_shamir_transcript.rs:42), as a PLAIN canonical integer
    ZK_DEV static void challenge_plain(Fe& out, const uint64_t digest[4]) {
        uint32_t s[10];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            s[2 * i] = (uint32_t)digest[i];
            s[2 * i + 1] = (uint32_t)(digest[i] >> 32);
        }
        s[8] = s[9] = 0;
        P::barrett(out.v, s);
    }
};

template <int FID, int P, int D, int NLIN, class Exec>
ZK_DEV void sumcheck_tail_body(const TailArgs& a, TailShared& sh, Exec& ex) {
    constexpr int T = P * D + NLIN, NE = D + 1;
    typedef RoundAcc<FID, P, D, false, NLIN> RA;   // s(1) is summed directly: no running claim to carry
    typedef TailField<FID> TF;
    const int tid = ex.tid(), nt = ex.nthreads();
    for (int i = tid; i < 64; i += nt) sh.ft.w[i >> 3][i & 7] = a.ft.w[i >> 3][i & 7];
    for (int i = tid; i < 25; i += nt) sh.sponge.s[i] = a.sponge.s[i];
    if (tid == 0) sh.sponge.pos = a.sponge.pos;
    ex.sync();
    uint64_t len = 1ull << a.log_len;
    bool pending = a.pending != 0;
    uint32_t round = 0;
    for (;;) {
        if (pending ? len == 2 : len < 2) break;
        RA ra;
        ra.init();
        if (pending) {   // fold table_{k-1} by r_{k-1} in place and evaluate round k (fold_evals_kernel's body)
            const uint64_t q = len / 4;
            for (uint64_t j = tid; j < q; j += nt) {
                Fe lo[T], hi[T];
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    Fe a0 = ex.load(a.tp.t[t] + j), a1 = ex.load(a.tp.t[t] + j + q);
                    Fe a2 = ex.load(a.tp.t[t] + j + 2 * q), a3 = ex.load(a.tp.t[t] + j + 3 * q);
                    FoldScalar<FID>::fold(lo[t], a0, a2, sh.ft);
                    FoldScalar<FID>::fold(hi[t], a1, a3, sh.ft);
                    ex.store(a.tp.t[t] + j, lo[t]);
                    ex.store(a.tp.t[t] + j + q, hi[t]);
                }
                ra.add_pair(lo, hi);
            }
            len /= 2;
        } else {         // first round of a small sumcheck: evaluations only (round_evals_kernel's body)
            const uint64_t half = len / 2;
            for (uint64_t j = tid; j < half; j += nt) {
                Fe lo[T], hi[T];
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    lo[t] = ex.load(a.tp.t[t] + j);
                    hi[t] = ex.load(a.tp.t[t] + j + half);
                }
                ra.add_pair(lo, hi);
            }
        }
        {
            uint32_t col[RA::NC];
            ra.columns(col);
            ex.template column_sums<RA::NC>(col, sh.tot);   // ends with a barrier: the folded tables are visible too
        }
        for (int e = tid; e < NE; e += nt) RA::finalize(sh.evals[e], e, sh.tot);
        ex.sync();
        if (a.mode == kTailProduct) {   // lagrange_interpolate on 0..D, then the coefficients little-endian
            for (int i = tid; i < NE; i += nt) {
                Fe c;
                Fp<FID>::mont_mul(c, a.interp[i * NE], sh.evals[0]);
#pragma unroll
                for (int k = 1; k < NE; ++k) {
                    Fe t;
                    Fp<FID>::mont_mul(t, a.interp[i * NE + k], sh.evals[k]);
                    Fp<FID>::add(c, c, t);
                }
                a.out->round_vals[round][i] = c;
                TF::le_words(sh.words[i], c);
            }
        } else {                        // plain sumcheck: the two half sums big-endian
            for (int i = tid; i < NE; i += nt) {
                a.out->round_vals[round][i] = sh.evals[i];
                TF::be_words(sh.words[i], sh.evals[i]);
            }
        }
        ex.sync();
        if (ex.warp() == 0) {   // the transcript step, one warp: absorb, sample, challenge (dev_transcript.cuh)
            uint32_t pos = sh.sponge.pos;
            coop_absorb_words(ex, sh.sponge.s, pos, &sh.words[0][0], NE * 4);
            coop_sample(ex, sh.sponge.s, pos, sh.digest);
            if (ex.lane() == 0) {
                sh.sponge.pos = pos;
                TF::challenge_plain(sh.r_plain, sh.digest);
            }
        }
        ex.sync();
        for (int i = tid; i < 9; i += nt) {   // rows of the next fold table, and the challenge as the proof reports it
            if (i < 8) {
                Fe row;
                Fp<FID>::mont_mul(row, sh.r_plain, a.pow32[i]);
#pragma unroll
                for (int k = 0; k < 8; ++k) sh.ft.w[i][k] = row.v[k];
            } else {
                Fe r2, rm;
#pragma unroll
                for (int k = 0; k < 8; ++k) r2.v[k] = FieldParams<FID>::r2(k);
                Fp<FID>::mont_mul(rm, sh.r_plain, r2);
                a.out->challenges[round] = rm;
            }
        }
        ex.sync();
        pending = true;
        ++round;
    }
    for (int t = tid; t < T; t += nt) {
        Fe o = ex.load(a.tp.t[t]);
        if (pending) {   // len == 2: the last partial_evaluate
            Fe hi = ex.load(a.tp.t[t] + 1);
            FoldScalar<FID>::fold(o, o, hi, sh.ft);
            ex.store(a.tp.t[t], o);
        }
        a.out->finals[t] = o;
    }
    for (int i = tid; i < 25; i += nt) a.out->sponge.s[i] = sh.sponge.s[i];
    if (tid == 0) {
        a.out->sponge.pos = sh.sponge.pos;
        a.out->rounds = round;
    }
    ex.sync();
    if (tid == 0) ex.publish(&a.out->seq, a.seq);
}

}  // namespace zk
