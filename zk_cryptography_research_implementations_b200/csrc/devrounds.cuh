// devrounds.cuh -- the rounds of a sumcheck as ONE device-resident launch: every round's sums, the fold, the
// Fiat-Shamir transcript and the challenge stay on the GPU; the host launches once and reads the proof.
//
// What one round is (reference: sumcheck_gkr_protocol.rs:37-60 / prover.rs:46-64): the round polynomial's evaluations
// summed over the table (generate_round_univariate, :113-143), Lagrange coefficients (dense_univariate.rs:74-127), the
// transcript absorb and challenge (fiat_shamir_transcript.rs:22-43), the fold of every table by the challenge
// (partial_evaluate, evaluation_form.rs:61-106).  Host-driven, that is one kernel + a PCIe mailbox + a host Keccak + the
// next launch per round: ~16 us of turnaround around a kernel that, below ~2^20 entries, runs for a few microseconds.
// Here a persistent grid (all blocks co-resident: cooperative launch) loops over the rounds:
//
//   every participating block   fold table_{k-1} by r_{k-1} in place + accumulate round k (RoundAcc / FoldScalar: the same
//                               code as the per-round kernels), exact column sums -> one RED per column into the grid
//                               accumulator, fence, arrive on a counter
//   block 0 (the leader)        waits for the arrivals, takes the column totals, d+1 evaluations (ONE Montgomery reduction
//                               each), [sharded: exchanges the partial evaluations with the peer GPUs by writing them
//                               straight into every peer's slot over NVLink and summing the G slots it received -- every
//                               rank then runs the same transcript, so the challenge needs no broadcast], coefficients,
//                               bytes, warp-wide Keccak-256, challenge, next fold table -> global memory, release
//   the other blocks            spin on the release word, fetch the fold table, next round
//
// A block whose slice of the (halving) tables has become empty leaves the loop -- the barrier only ever counts the blocks
// that still have work -- so the launch degenerates into the single-block tail by itself.  Tables are read with L2-only
// loads (another block wrote them a round ago).  Every spin is bounded (wall clock): a lost peer or a bug ends in an error
// status, never in a hung GPU.
//
// The body is written against an execution policy and is split into the three sub-steps of a round (compute / post /
// finish) so that tests/host_emu can run whole multi-block, multi-rank proves on the CPU against the oracle.
#pragma once
#include "dev_transcript.cuh"
#include "round_acc.cuh"

namespace zk {

constexpr int kDevMaxRounds = 32;    // rounds one launch can run (tables of at most 2^32 entries)
constexpr int kDevMaxLog = 32;
constexpr int kMaxRanks = 16;
enum DevMode { kDevProduct = 0, kDevPlain = 1 };
enum DevStatus { kDevOk = 0, kDevTimeoutArrive = 1, kDevTimeoutRelease = 2, kDevTimeoutPeer = 3 };

struct DevOut {                                 // mapped pinned host memory
    Fe round_vals[kDevMaxRounds][kMaxEvals];    // product: the d+1 coefficients per round; plain: [sum left, sum right]
    Fe challenges[kDevMaxRounds];               // Montgomery form, as the host provers report them
    Fe finals[kMaxTables];                      // the tables' single entries after the last fold
    KeccakState sponge;                         // transcript state after the last round
    uint32_t rounds;
    uint32_t status;                            // DevStatus
    uint32_t seq;                               // written last, after a system-scope fence
#ifdef ZK_DEV_TIMING
    // measurement builds only (build.py ZKB200_DEFINES=-DZK_DEV_TIMING): the leader's clock (ns) per round at
    // [0] round start, [1] own slice computed, [2] all blocks arrived, [3] evaluations (+ peer exchange) done, [4] challenge out
    unsigned long long t_ns[kDevMaxRounds][5];
#endif
};

struct PeerSlot {                               // one rank's partial evaluations of one round, in the RECEIVER's memory
    Fe vals[kMaxEvals];
    uint32_t seq;
    uint32_t pad[7];
};

struct DevGlobal {                              // device memory, one per context; all zero between launches
    unsigned long long gacc[kMaxCols];          // grid-wide column totals of the current round
    // the two barrier words live in cache lines of their own: the waiting blocks poll `release` while the arriving blocks'
    // atomics go to `arrive` -- in one line the polling loads would queue up in front of the atomics
    alignas(128) uint32_t arrive;               // blocks that have contributed (cumulative over the rounds of a launch)
    alignas(128) uint32_t release;              // rounds completed by the leader in this launch
    uint32_t abort_;
    alignas(128) FoldTable ft;                  // fold table of the latest challenge
};

struct DevArgs {
    TablePtrs tp;
    uint32_t log_len;      // every table holds 2^log_len entries on entry
    uint32_t pending;      // 1: the tables still have to be folded by `ft`
    uint32_t mode;         // DevMode
    uint32_t seq;          // value to publish in out->seq
    uint32_t max_rounds;   // stop after this many rounds even if the tables are not exhausted (the last challenge is then
                           // left pending: sharded provers stop at the collapse point)
    uint32_t final_fold;   // 1: when the rounds end with two entries per table and a pending challenge, do that last
                           // partial_evaluate and report the single entries (an unsharded prove); 0: leave it pending (a rank's
                           // shard: more rounds follow after the collapse even if the LOCAL tables are exhausted)
    uint32_t world, rank;  // world > 1: the tables are this rank's shard, partial evaluations are exchanged per round
    uint32_t xseq;         // sequence number of the round before this launch's first one (agreed by all ranks)
    FoldTable ft;          // fold table of the pending challenge
    Fe interp[kMaxEvals * kMaxEvals];   // inverse Vandermonde on the nodes 0..D, row-major, Montgomery form
    Fe interp_plain[kMaxEvals * kMaxEvals];   // the same matrix as plain canonical integers (gives the coefficients' bytes directly)
    Fe pow32[8];           // Montgomery forms of 2^(32 i)
    KeccakState sponge;    // transcript state on entry
    DevOut* out;
    DevGlobal* g;
    PeerSlot* peers[kMaxRanks];   // peers[q]: rank q's slot array [2][world] as mapped in this process (peers[rank]: our own)
};

struct DevShared {
    FoldTable ft;
    unsigned long long tot[kMaxCols];
    // leader only
    KeccakState sponge;
    Fe evals[kMaxEvals];
    uint64_t words[kMaxEvals][4];   // what the transcript absorbs this round, as little-endian words of the byte stream
    uint64_t digest[4];
    Fe r_plain;
};

template <int FID> struct DevField {
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

// Work distribution: the `work` items of a round go to WARPS, and consecutive warps sit in different blocks (global warp
// w * nblocks + b is warp w of block b), so that a small round spreads over many SMs -- one warp each -- instead of
// filling a few of them: below ~2^15 items the round is then bound by the latency of ONE iteration, not by one SM's
// multiplier throughput.  Block b has work iff its warp 0 has: b * lanes < work.
ZK_DEV uint32_t dev_participants(uint64_t work, uint32_t nblocks, uint32_t lanes) {
    const uint64_t b = (work + lanes - 1) / lanes;
    return b >= nblocks ? nblocks : (b < 1 ? 1u : (uint32_t)b);
}

// One block's view of the launch.  init(), then step() until it returns false.
template <int FID, int P, int D, int NLIN, class Exec> struct DevRounds {
    static constexpr int T = P * D + NLIN, NE = D + 1;
    typedef RoundAcc<FID, P, D, false, NLIN> RA;   // s(1) is summed directly
    typedef DevField<FID> TF;
    const DevArgs& a;
    DevShared& sh;
    Exec& ex;
    uint64_t len;
    bool pending;
    uint32_t round;        // rounds of this launch completed so far
    uint32_t arrived;      // leader: arrivals expected so far (cumulative)
    uint32_t np;           // participants of the round being computed
    bool failed;

    ZK_DEV DevRounds(const DevArgs& a_, DevShared& sh_, Exec& ex_) : a(a_), sh(sh_), ex(ex_) {}

    ZK_DEV void init() {
        const int tid = ex.tid(), nt = ex.nthreads();
        for (int i = tid; i < 64; i += nt) sh.ft.w[i >> 3][i & 7] = a.ft.w[i >> 3][i & 7];
        if (ex.bid() == 0) {
            for (int i = tid; i < 25; i += nt) sh.sponge.s[i] = a.sponge.s[i];
            if (tid == 0) sh.sponge.pos = a.sponge.pos;
        }
        ex.sync();
        len = 1ull << a.log_len;
        pending = a.pending != 0;
        round = 0;
        arrived = 0;
        failed = false;
    }
    ZK_DEV bool rounds_left() const { return !(pending ? len == 2 : len < 2) && round < a.max_rounds; }
    ZK_DEV uint64_t work_now() const { return pending ? len / 4 : len / 2; }
    ZK_DEV uint64_t work_after() const { return (pending ? len / 2 : len) / 4; }   // next round's quads (it folds)

    // ---- sub-step 1, every participating block: (fetch the challenge,) fold + accumulate, contribute, arrive
    ZK_DEV void stamp(int slot) {
#ifdef ZK_DEV_TIMING
        if (ex.bid() == 0 && ex.tid() == 0 && round < (uint32_t)kDevMaxRounds) a.out->t_ns[round][slot] = ex.now_ns();
#else
        (void)slot;
#endif
    }
    ZK_DEV bool step_compute() {
        const int tid = ex.tid(), nt = ex.nthreads();
        stamp(0);
        np = dev_participants(work_now(), ex.nblocks(), (uint32_t)ex.lanes());
        if (ex.bid() != 0 && round > 0) {   // the leader released round - 1: its fold table is in global memory
            if (!ex.wait_release(a.g, round)) { failed = true; return false; }
            for (int i = tid; i < 64; i += nt) sh.ft.w[i >> 3][i & 7] = ex.load_word(&a.g->ft.w[i >> 3][i & 7]);
            ex.sync();
        }
        RA ra;
        ra.init();
        const uint64_t stride = (uint64_t)ex.nblocks() * nt;
        const uint64_t first = ((uint64_t)ex.warp() * ex.nblocks() + ex.bid()) * ex.lanes() + ex.lane();
        if (pending) {   // fold table_{k-1} by r_{k-1} in place and evaluate round k (fold_evals_kernel's body)
            const uint64_t q = len / 4;
            for (uint64_t j = first; j < q; j += stride) {
                if (j + stride < q) {
#pragma unroll
                    for (int t = 0; t < T; ++t)
#pragma unroll
                        for (int s = 0; s < 4; ++s) ex.prefetch(a.tp.t[t] + j + stride + s * q);
                }
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
        } else {         // first round of a sumcheck: evaluations only (round_evals_kernel's body)
            const uint64_t half = len / 2;
            for (uint64_t j = first; j < half; j += stride) {
                if (j + stride < half) {
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        ex.prefetch(a.tp.t[t] + j + stride);
                        ex.prefetch(a.tp.t[t] + j + stride + half);
                    }
                }
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
            ex.template column_sums<RA::NC>(col, sh.tot);   // ends with a barrier
        }
        if (np > 1) {   // contribute to the grid totals; the fence inside arrive() also publishes this block's folded entries
            for (int c = tid; c < RA::NC; c += nt) ex.grid_add(&a.g->gacc[c], sh.tot[c]);
            ex.arrive(a.g);
        }
        stamp(1);
        return true;
    }

    // ---- sub-step 2, leader: the round's evaluations; sharded: hand the partial evaluations to every rank
    ZK_DEV bool step_post() {
        const int tid = ex.tid(), nt = ex.nthreads();
        if (np > 1) {
            arrived += np;
            if (!ex.wait_arrivals(a.g, arrived)) { failed = true; return false; }
            for (int c = tid; c < RA::NC; c += nt) sh.tot[c] = ex.grid_take(&a.g->gacc[c]);
            ex.sync();
        }
        stamp(2);
        for (int e = tid; e < NE; e += nt) RA::finalize(sh.evals[e], e, sh.tot);
        ex.sync();
        if (a.world > 1) {
            const uint32_t xs = a.xseq + round + 1;
            for (int i = tid; i < (int)a.world * NE; i += nt) {
                const int q = i / NE, e = i % NE;
                ex.store_peer(&a.peers[q][(xs & 1u) * a.world + a.rank].vals[e], sh.evals[e]);
            }
            ex.sync();
            for (int q = tid; q < (int)a.world; q += nt) ex.publish_peer(&a.peers[q][(xs & 1u) * a.world + a.rank].seq, xs);
        }
        return true;
    }

    // ---- sub-step 3, leader: (sum the ranks' partials,) coefficients, transcript, challenge, fold table, release
    ZK_DEV bool step_finish() {
        const int tid = ex.tid(), nt = ex.nthreads();
        if (a.world > 1) {
            const uint32_t xs = a.xseq + round + 1;
            PeerSlot* mine = a.peers[a.rank] + (xs & 1u) * a.world;
            if (!ex.wait_peers(mine, a.world, xs)) { failed = true; return false; }
            for (int e = tid; e < NE; e += nt) {
                Fe acc = ex.load_peer(&mine[0].vals[e]);
                for (uint32_t q = 1; q < a.world; ++q) {
                    Fe v = ex.load_peer(&mine[q].vals[e]);
                    Fp<FID>::add(acc, acc, v);
                }
                sh.evals[e] = acc;
            }
            ex.sync();
        }
        stamp(3);
        if (a.mode == kDevProduct) {
            // lagrange_interpolate on 0..D as a fixed matrix: coefficient i = sum_k M[i][k] s(k), accumulated UNREDUCED with one
            // Montgomery reduction.  2 (D+1) threads: with M in Montgomery form the reduction yields the coefficient as the
            // proof reports it; with M as plain integers it yields the coefficient's canonical integer -- the little-endian
            // bytes the transcript absorbs (sumcheck_gkr_protocol.rs:145-150) -- without a second pass.
            for (int j = tid; j < 2 * NE; j += nt) {
                const int i = j < NE ? j : j - NE;
                const Fe* m = (j < NE ? a.interp : a.interp_plain) + i * NE;
                uint32_t acc[17];
#pragma unroll
                for (int k = 0; k < 17; ++k) acc[k] = 0;
#pragma unroll
                for (int k = 0; k < NE; ++k) Fp<FID>::mul_acc(acc, m[k], sh.evals[k]);
                Fe c;
                Fp<FID>::redc_wide(c, acc);
                if (j < NE) {
                    a.out->round_vals[round][i] = c;
                } else {
#pragma unroll
                    for (int w = 0; w < 4; ++w) sh.words[i][w] = (uint64_t)c.v[2 * w] | ((uint64_t)c.v[2 * w + 1] << 32);
                }
            }
        } else {                       // plain sumcheck: the two half sums big-endian
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
        stamp(4);
        return true;
    }

    // advance the bookkeeping after a round; leader: hand the fold table to the blocks that go on
    ZK_DEV void advance() {
        const int tid = ex.tid(), nt = ex.nthreads();
        if (pending) len /= 2;
        pending = true;
        ++round;
        if (ex.bid() == 0 && rounds_left() && dev_participants(work_now(), ex.nblocks(), (uint32_t)ex.lanes()) > 1) {
            for (int i = tid; i < 64; i += nt) ex.store_word(&a.g->ft.w[i >> 3][i & 7], sh.ft.w[i >> 3][i & 7]);
            ex.release(a.g, round);   // barrier + fence + the release word
        }
    }

    // leader, after the last round: the last partial_evaluate, results, re-arm the global state, publish
    ZK_DEV void finish_launch(uint32_t status) {
        const int tid = ex.tid(), nt = ex.nthreads();
        if (status == kDevOk && a.final_fold && pending && len == 2) {
            for (int t = tid; t < T; t += nt) {
                Fe lo = ex.load(a.tp.t[t]), hi = ex.load(a.tp.t[t] + 1), o;
                FoldScalar<FID>::fold(o, lo, hi, sh.ft);
                ex.store(a.tp.t[t], o);
                a.out->finals[t] = o;
            }
        } else if (status == kDevOk && a.final_fold && !pending && len == 1) {
            for (int t = tid; t < T; t += nt) a.out->finals[t] = ex.load(a.tp.t[t]);
        }
        for (int i = tid; i < 25; i += nt) a.out->sponge.s[i] = sh.sponge.s[i];
        if (tid == 0) {
            a.out->sponge.pos = sh.sponge.pos;
            a.out->rounds = round;
            a.out->status = status;
        }
        ex.rearm(a.g, status != kDevOk);
        ex.sync();
        if (tid == 0) ex.publish(&a.out->seq, a.seq);
    }

    // one round of this block; false when the block has nothing more to do
    ZK_DEV bool step() {
        if (!rounds_left()) {
            if (ex.bid() == 0) finish_launch(kDevOk);
            return false;
        }
        if (ex.bid() >= dev_participants(work_now(), ex.nblocks(), (uint32_t)ex.lanes())) return false;
        bool ok = step_compute();
        if (ok && ex.bid() == 0) ok = step_post() && step_finish();
        if (!ok) {
            if (ex.bid() == 0) finish_launch(ex.failure());
            return false;
        }
        advance();
        return true;
    }
};

}  // namespace zk
