// Host emulation of zk::Fp (fp.cuh built with -DZK_HOST_EMU) checked against the C oracle.
// Test-only program (tests/test_host_emu.py builds and runs it); exits non-zero on mismatch.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "fp.cuh"
#include "round_acc.cuh"
#include "devrounds.cuh"
#include "host_field.h"
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

// RoundAcc (round_acc.cuh): several "threads" accumulate pairs, their 32-bit columns are summed as exact
// 64-bit integers (what the kernels' REDUX / RED stages do) and finalize() must return the reference's
// generate_round_univariate evaluations (oracle: zko_generate_round_univariate on the same tables).
template <int FID, int P, int D, bool SKIP1, int NLIN> static int run_round_acc(int n_threads, int pairs_per_thread, int edge) {
    typedef zk::RoundAcc<FID, P, D, SKIP1, NLIN> RA;
    constexpr int T = P * D + NLIN;
    const uint64_t half = (uint64_t)n_threads * pairs_per_thread, len = 2 * half;   // must be a power of two
    // oracle layout: P' products of D factors; a linear table l enters as the product (l, ones, ones...)
    constexpr int PO = P + NLIN;
    std::vector<uint64_t> tabs((size_t)PO * D * len * 4);
    uint64_t one[4]; zko_fe_from_u64(FID, 1, one);
    for (int p = 0; p < PO; ++p)
        for (int d = 0; d < D; ++d)
            for (uint64_t i = 0; i < len; ++i) {
                uint64_t* dst = &tabs[(((size_t)p * D + d) * len + i) * 4];
                if (p >= P && d > 0) memcpy(dst, one, 32);
                else rand_fe(FID, dst, edge ? (int)(rnd() % 6) : 0);
            }
    std::vector<unsigned long long> tot(RA::NC, 0);
    for (int th = 0; th < n_threads; ++th) {
        RA ra; ra.init();
        for (int it = 0; it < pairs_per_thread; ++it) {
            uint64_t j = (uint64_t)it * n_threads + th;
            zk::Fe lo[T], hi[T];
            for (int t = 0; t < T; ++t) {
                // table t of the kernel: products first, then the linear tables
                int p = t < P * D ? t / D : P + (t - P * D), d = t < P * D ? t % D : 0;
                const uint64_t* base = &tabs[(((size_t)p * D + d) * len) * 4];
                lo[t] = to_fe(base + j * 4); hi[t] = to_fe(base + (j + half) * 4);
            }
            ra.add_pair(lo, hi);
        }
        uint32_t col[RA::NC]; ra.columns(col);
        for (int c = 0; c < RA::NC; ++c) tot[c] += col[c];
    }
    std::vector<uint64_t> want((D + 1) * 4);
    zko_generate_round_univariate(FID, tabs.data(), PO, D, len, want.data());
    int bad = 0;
    for (int e = 0; e <= D; ++e) {
        if (SKIP1 && e == 1) continue;
        zk::Fe out; RA::finalize(out, e, tot.data());
        if (!eq(out, &want[4 * e])) { ++bad; printf("RoundAcc<%d,%d,%d,%d,%d> eval %d mismatch\n", FID, P, D, (int)SKIP1, NLIN, e); }
    }
    return bad;
}
template <int FID> static int run_round_accs() {
    int bad = 0;
    for (int edge = 0; edge < 2; ++edge) {
        bad += run_round_acc<FID, 1, 1, false, 0>(8, 4, edge);
        bad += run_round_acc<FID, 1, 2, false, 0>(4, 4, edge);
        bad += run_round_acc<FID, 1, 2, true, 0>(256, 2, edge);
        bad += run_round_acc<FID, 2, 2, false, 0>(4, 4, edge);
        bad += run_round_acc<FID, 1, 2, false, 1>(8, 2, edge);
        bad += run_round_acc<FID, 1, 2, true, 1>(1, 1, edge);
        bad += run_round_acc<FID, 2, 3, false, 0>(4, 2, edge);
        bad += run_round_acc<FID, 4, 2, true, 0>(2, 2, edge);
    }
    printf("round_acc field %d: %s\n", FID, bad ? "FAIL" : "ok");
    return bad;
}

// ---------------------------------------------------------------------------------------------------------------
// The device transcript and the whole tail body (tail.cuh) under a one-thread host policy, against the oracle.
// 32 lanes of a warp as an array: the host stand-in for one-value-per-thread registers, so that warp_keccak_f1600
// (dev_transcript.cuh) runs here exactly as written
struct LaneVec { uint32_t v[32]; };
#define LV_BIN(OP) \
    static LaneVec operator OP(const LaneVec& a, const LaneVec& b) { LaneVec r; for (int i = 0; i < 32; ++i) r.v[i] = a.v[i] OP b.v[i]; return r; } \
    static LaneVec operator OP(const LaneVec& a, uint32_t b) { LaneVec r; for (int i = 0; i < 32; ++i) r.v[i] = a.v[i] OP b; return r; }
LV_BIN(^) LV_BIN(&) LV_BIN(|) LV_BIN(+)
#undef LV_BIN
static LaneVec operator>>(const LaneVec& a, int n) { LaneVec r; for (int i = 0; i < 32; ++i) r.v[i] = a.v[i] >> n; return r; }
static LaneVec operator~(const LaneVec& a) { LaneVec r; for (int i = 0; i < 32; ++i) r.v[i] = ~a.v[i]; return r; }
// (global namespace, next to LaneVec: found by argument-dependent lookup from the template in dev_transcript.cuh)
static LaneVec wk_shfl(const LaneVec& v, const LaneVec& src) { LaneVec r; for (int i = 0; i < 32; ++i) r.v[i] = v.v[src.v[i] & 31]; return r; }
static uint32_t funnel1(uint32_t lo, uint32_t hi, uint32_t n) { n &= 31; return n ? (hi << n) | (lo >> (32 - n)) : hi; }
static LaneVec wk_funnel_l(const LaneVec& lo, const LaneVec& hi, const LaneVec& n) { LaneVec r; for (int i = 0; i < 32; ++i) r.v[i] = funnel1(lo.v[i], hi.v[i], n.v[i]); return r; }
static LaneVec wk_funnel_l(const LaneVec& lo, const LaneVec& hi, uint32_t n) { LaneVec r; for (int i = 0; i < 32; ++i) r.v[i] = funnel1(lo.v[i], hi.v[i], n); return r; }
static LaneVec wk_lane0_mask(const LaneVec&) { LaneVec r; for (int i = 0; i < 32; ++i) r.v[i] = i == 0 ? 0xffffffffu : 0u; return r; }
static void lanes_permute(uint64_t a[25]) {   // what CudaExec::permute does, on the lane array
    LaneVec lo, hi, A, B;
    for (int i = 0; i < 32; ++i) { uint64_t w = i < 25 ? a[i] : 0x1111111111111111ull * i; lo.v[i] = (uint32_t)w; hi.v[i] = (uint32_t)(w >> 32); A.v[i] = zk::kWkA[i]; B.v[i] = zk::kWkB[i]; }
    zk::warp_keccak_f1600<LaneVec>(lo, hi, A, B);
    for (int i = 0; i < 25; ++i) a[i] = (uint64_t)lo.v[i] | ((uint64_t)hi.v[i] << 32);
}

struct HostExec {   // one "thread" per block; blocks (and ranks) are stepped in an order that satisfies every wait
    uint32_t bid_ = 0, nblocks_ = 1;
    int lane() const { return 0; }
    int warp() const { return 0; }
    uint32_t lanes() const { return 1; }
    unsigned long long now_ns() const { return 0; }
    void permute(uint64_t* s) const { lanes_permute(s); }
    void finalize(const uint64_t* s, uint32_t pos, uint64_t* digest) const {
        uint64_t a[25];
        memcpy(a, s, sizeof a);
        a[pos >> 3] ^= 0x01ull << (8 * (pos & 7));
        a[16] ^= 0x8000000000000000ull;
        lanes_permute(a);
        memcpy(digest, a, 32);
    }
    int tid() const { return 0; }
    int nthreads() const { return 1; }
    uint32_t bid() const { return bid_; }
    uint32_t nblocks() const { return nblocks_; }
    uint32_t failure() const { return 9; }
    void sync() const {}
    template <int NC> void column_sums(const uint32_t (&col)[NC], unsigned long long* tot) const { for (int c = 0; c < NC; ++c) tot[c] = col[c]; }
    zk::Fe load(const zk::Fe* p) const { return *p; }
    void store(zk::Fe* p, const zk::Fe& v) const { *p = v; }
    void prefetch(const zk::Fe*) const {}
    uint32_t load_word(const uint32_t* p) const { return *p; }
    void store_word(uint32_t* p, uint32_t v) const { *p = v; }
    void grid_add(unsigned long long* p, unsigned long long v) const { *p += v; }
    unsigned long long grid_take(unsigned long long* p) const { unsigned long long v = *p; *p = 0; return v; }
    void arrive(zk::DevGlobal* g) const { g->arrive += 1; }
    // the stepping order makes every wait already satisfied; anything else is a bug in the round bookkeeping
    bool wait_arrivals(zk::DevGlobal* g, uint32_t target) const { return g->arrive == target; }
    void release(zk::DevGlobal* g, uint32_t round) const { g->release = round; }
    bool wait_release(zk::DevGlobal* g, uint32_t round) const { return g->release >= round; }
    void store_peer(zk::Fe* p, const zk::Fe& v) const { *p = v; }
    void publish_peer(uint32_t* seq, uint32_t v) const { *seq = v; }
    bool wait_peers(zk::PeerSlot* mine, uint32_t world, uint32_t xs) const {
        for (uint32_t q = 0; q < world; ++q) if (mine[q].seq != xs) return false;
        return true;
    }
    zk::Fe load_peer(const zk::Fe* p) const { return *p; }
    void rearm(zk::DevGlobal* g, bool failed) const { if (failed) g->abort_ = 1; else { g->arrive = 0; g->release = 0; } }
    void publish(uint32_t* seq, uint32_t v) const { *seq = v; }
};

// Runs one launch of the device-resident round loop on `nblocks` emulated blocks (one thread each, so a table of a few
// entries already spreads over several blocks).  Blocks are stepped round by round, highest block first: a non-leader's
// wait for the previous challenge and the leader's wait for the arrivals are then always satisfied.
template <int FID, int P, int D, int NLIN> static int emu_launch(zk::DevArgs& a, int nblocks) {
    typedef zk::DevRounds<FID, P, D, NLIN, HostExec> Blk;
    zk::DevGlobal g; memset(&g, 0, sizeof g);
    a.g = &g;
    std::vector<zk::DevShared> sh(nblocks);
    std::vector<HostExec> ex(nblocks);
    std::vector<Blk*> blk(nblocks);
    std::vector<char> alive(nblocks, 1);
    for (int b = 0; b < nblocks; ++b) { ex[b].bid_ = b; ex[b].nblocks_ = nblocks; blk[b] = new Blk(a, sh[b], ex[b]); blk[b]->init(); }
    for (int guard = 0; guard < 100 && alive[0]; ++guard)
        for (int b = nblocks - 1; b >= 0; --b)
            if (alive[b]) alive[b] = blk[b]->step();
    int bad = 0;
    for (int b = 0; b < nblocks; ++b) { if (alive[b] || blk[b]->failed) { ++bad; printf("emulated block %d did not finish cleanly\n", b); } delete blk[b]; }
    if (g.arrive != 0 || g.release != 0 || g.abort_ != 0) { ++bad; printf("the launch did not re-arm its global state\n"); }
    for (int c = 0; c < zk::kMaxCols; ++c) if (g.gacc[c] != 0) { ++bad; printf("grid accumulator column %d left non-zero\n", c); break; }
    return bad;
}

static int test_dev_sponge() {
    int bad = 0;
    // the warp-wide permutation (lane array) against the one-thread permutation
    for (int it = 0; it < 200; ++it) {
        uint64_t a[25], b[25];
        for (int i = 0; i < 25; ++i) a[i] = b[i] = it == 0 ? 0 : rnd();
        zk::keccak_f1600(a);
        lanes_permute(b);
        if (memcmp(a, b, sizeof a)) { ++bad; printf("warp keccak differs from the scalar permutation (it=%d)\n", it); break; }
    }
    // cooperative absorb / sample (what the tail kernel's warp 0 runs) at every alignment against the oracle
    for (int prefix = 0; prefix < 140; ++prefix) {
        std::vector<uint8_t> data(prefix + 8 * 64);
        for (auto& x : data) x = (uint8_t)rnd();
        zk::KeccakState st; memset(&st, 0, sizeof st);
        zko_transcript* ot = zko_transcript_new();
        for (int i = 0; i < prefix; ++i) zk::sponge_absorb_byte(&st, data[i]);
        zko_transcript_append(ot, data.data(), prefix);
        HostExec ex;
        uint32_t pos = st.pos;
        for (int rep = 0; rep < 4; ++rep) {
            uint64_t words[16]; memcpy(words, &data[prefix + 128 * rep], 128);
            int nw = 12 + rep;   // 96 .. 120 bytes per step
            zk::coop_absorb_words(ex, st.s, pos, words, nw);
            zko_transcript_append(ot, &data[prefix + 128 * rep], 8 * nw);
            uint64_t dg[4]; uint8_t want[32];
            zk::coop_sample(ex, st.s, pos, dg);
            zko_transcript_sample(ot, want);
            if (memcmp(dg, want, 32)) { ++bad; printf("cooperative sponge digest mismatch prefix=%d rep=%d\n", prefix, rep); }
        }
        zko_transcript_free(ot);
    }
    // byte-wise and word-wise absorbs at every alignment against the oracle's Keccak / transcript
    for (int prefix = 0; prefix < 300; prefix += (prefix < 20 ? 1 : 37)) {
        std::vector<uint8_t> data(prefix + 8 * 40);
        for (auto& b : data) b = (uint8_t)rnd();
        zk::KeccakState st; memset(&st, 0, sizeof st);
        zko_transcript* ot = zko_transcript_new();
        for (int i = 0; i < prefix; ++i) zk::sponge_absorb_byte(&st, data[i]);
        zko_transcript_append(ot, data.data(), prefix);
        for (int rep = 0; rep < 5; ++rep) {
            for (int w = 0; w < 8; ++w) { uint64_t word; memcpy(&word, &data[prefix + 8 * (8 * rep + w)], 8); zk::sponge_absorb_word(&st, word); }
            zko_transcript_append(ot, &data[prefix + 64 * rep], 64);
            uint64_t dg[4]; uint8_t want[32];
            zk::sponge_sample(&st, dg);
            zko_transcript_sample(ot, want);
            if (memcmp(dg, want, 32)) { ++bad; printf("device sponge digest mismatch prefix=%d rep=%d\n", prefix, rep); }
        }
        // the host transcript of the library takes the state over and continues identically
        zk::HostTranscript ht; ht.import_state(st.s, st.pos);
        uint8_t a[32], b[32]; ht.sample(a); zko_transcript_sample(ot, b);
        if (memcmp(a, b, 32)) { ++bad; printf("state handback mismatch prefix=%d\n", prefix); }
        zko_transcript_free(ot);
    }
    printf("dev_sponge: %s\n", bad ? "FAIL" : "ok");
    return bad;
}

static zk::FoldTable host_fold_table(const zk::HostField& f, const zk::HFe& r_mont) {
    zk::FoldTable ft; zk::HFe cur = f.from_mont(r_mont), m232 = f.from_u64(1ull << 32);
    for (int i = 0; i < 8; ++i) { memcpy(ft.w[i], cur.l, 32); cur = f.mul(cur, m232); }
    return ft;
}
template <int FID> static void fill_tail_consts(zk::DevArgs& a, const zk::HostField& f, int D) {
    zk::Interpolator ip(f, D);
    memcpy(a.interp, ip.matrix(), (size_t)(D + 1) * (D + 1) * 32);
    for (int i = 0; i < (D + 1) * (D + 1); ++i) { zk::HFe pl = f.from_mont(ip.matrix()[i]); memcpy(a.interp_plain[i].v, pl.l, 32); }
    zk::HFe cur = f.one(), m232 = f.from_u64(1ull << 32);
    for (int i = 0; i < 8; ++i) { memcpy(a.pow32[i].v, cur.l, 32); cur = f.mul(cur, m232); }
}

// product sumcheck: `host_rounds` rounds are run the way the host driver runs them, the rest by the tail body
template <int FID, int P, int D, int NLIN> static int run_tail_product(int n, int host_rounds, int prefix_bytes, int nblocks = 1) {
    constexpr int T = P * D + NLIN, PO = P + NLIN + (P + NLIN < 2 ? 1 : 0), NE = D + 1;
    const uint64_t len = 1ull << n;
    zk::HostField f(FID);
    // oracle tables: P products, NLIN products (l, 1, ..), and a 0*0 product when the oracle needs a second one
    std::vector<uint64_t> ot((size_t)PO * D * len * 4, 0);
    uint64_t one[4]; zko_fe_from_u64(FID, 1, one);
    std::vector<std::vector<zk::Fe>> tabs(T, std::vector<zk::Fe>(len));
    for (int t = 0; t < T; ++t) {
        int p = t < P * D ? t / D : P + (t - P * D), d = t < P * D ? t % D : 0;
        for (uint64_t i = 0; i < len; ++i) {
            uint64_t* dst = &ot[(((size_t)p * D + d) * len + i) * 4];
            rand_fe(FID, dst, (rnd() % 16 == 0) ? (int)(rnd() % 6) : 0);
            tabs[t][i] = to_fe(dst);
        }
    }
    for (int l = 0; l < NLIN; ++l)
        for (int d = 1; d < D; ++d)
            for (uint64_t i = 0; i < len; ++i) memcpy(&ot[(((size_t)(P + l) * D + d) * len + i) * 4], one, 32);
    // claimed sum and the oracle proof (transcript pre-loaded with some bytes so the sponge position is odd)
    std::vector<uint64_t> red(len * 4); uint64_t claim[4];
    zko_sumpoly_reduce(FID, ot.data(), PO, D, len, red.data()); zko_fe_sum(FID, red.data(), len, claim);
    std::vector<uint8_t> prefix(prefix_bytes); for (auto& b : prefix) b = (uint8_t)rnd();
    zko_transcript* otr = zko_transcript_new(); zko_transcript_append(otr, prefix.data(), prefix.size());
    std::vector<uint64_t> wc((size_t)n * NE * 4), wch((size_t)n * 4), wfin((size_t)PO * D * 4);
    if (zko_product_prove(FID, ot.data(), PO, D, len, claim, otr, wc.data(), wch.data(), wfin.data())) { printf("oracle refused\n"); return 1; }
    // the library's side: host transcript + host rounds (oracle pieces stand in for the big kernels), then the tail
    zk::HostTranscript tr; tr.append(prefix.data(), prefix.size());
    zk::HFe hclaim; memcpy(hclaim.l, claim, 32); tr.append_be(f, hclaim);
    zk::Interpolator ip(f, D);
    int bad = 0;
    zk::HFe r = f.zero();
    uint64_t cur_len = len;
    std::vector<uint64_t> cur_ot = ot;
    for (int k = 0; k < host_rounds; ++k) {
        zk::HFe ev[8], co[8];
        zko_generate_round_univariate(FID, cur_ot.data(), PO, D, cur_len, (uint64_t*)ev);
        ip.coefficients(ev, co);
        uint8_t bytes[32 * 8];
        for (int i = 0; i < NE; ++i) f.to_bytes_le(co[i], bytes + 32 * i);
        tr.append(bytes, 32 * NE);
        r = tr.challenge(f);
        if (memcmp(co, &wc[(size_t)k * NE * 4], 32 * NE) || memcmp(r.l, &wch[(size_t)k * 4], 32)) { ++bad; printf("host round %d mismatch\n", k); }
        if (k + 1 < host_rounds) {   // fold everything (the last host challenge stays pending for the tail)
            std::vector<uint64_t> nxt((size_t)PO * D * (cur_len / 2) * 4);
            for (int t = 0; t < PO * D; ++t)
                zko_mle_partial_evaluate(FID, &cur_ot[(size_t)t * cur_len * 4], cur_len, 0, r.l, &nxt[(size_t)t * (cur_len / 2) * 4]);
            cur_ot.swap(nxt); cur_len /= 2;
        }
    }
    // tables as the kernels would hold them at hand-over: length cur_len, fold by r pending (if any host round ran)
    for (int t = 0; t < T; ++t) {
        int p = t < P * D ? t / D : P + (t - P * D), d = t < P * D ? t % D : 0;
        tabs[t].resize(cur_len);
        for (uint64_t i = 0; i < cur_len; ++i) tabs[t][i] = to_fe(&cur_ot[(((size_t)p * D + d) * cur_len + i) * 4]);
    }
    zk::DevArgs a; memset(&a, 0, sizeof a);
    zk::DevOut out; memset(&out, 0, sizeof out);
    for (int t = 0; t < T; ++t) a.tp.t[t] = tabs[t].data();
    a.log_len = 0; while ((1ull << a.log_len) < cur_len) ++a.log_len;
    a.pending = host_rounds > 0; a.mode = zk::kDevProduct; a.seq = 77; a.world = 1; a.final_fold = 1;
    a.max_rounds = a.log_len - (a.pending ? 1 : 0);
    if (a.pending) a.ft = host_fold_table(f, r);
    fill_tail_consts<FID>(a, f, D);
    tr.export_state(a.sponge.s, &a.sponge.pos);
    a.out = &out;
    bad += emu_launch<FID, P, D, NLIN>(a, nblocks);
    if (out.seq != 77 || out.status != zk::kDevOk || (int)out.rounds != n - host_rounds) { ++bad; printf("tail rounds %u status %u (want %d)\n", out.rounds, out.status, n - host_rounds); }
    for (int k = host_rounds; k < n; ++k) {
        if (memcmp(out.round_vals[k - host_rounds], &wc[(size_t)k * NE * 4], 32 * NE)) { ++bad; printf("tail<%d,%d,%d,%d> n=%d coeffs of round %d mismatch\n", FID, P, D, NLIN, n, k); }
        if (memcmp(&out.challenges[k - host_rounds], &wch[(size_t)k * 4], 32)) { ++bad; printf("tail challenge of round %d mismatch\n", k); }
    }
    for (int t = 0; t < T; ++t) {
        int p = t < P * D ? t / D : P + (t - P * D), d = t < P * D ? t % D : 0;
        if (memcmp(&out.finals[t], &wfin[((size_t)p * D + d) * 4], 32) || memcmp(&tabs[t][0], &out.finals[t], 32)) { ++bad; printf("tail final value of table %d mismatch\n", t); }
    }
    // the transcript continues on the host exactly where the oracle's is
    tr.import_state(out.sponge.s, out.sponge.pos);
    uint8_t d1[32], d2[32]; tr.sample(d1); zko_transcript_sample(otr, d2);
    if (memcmp(d1, d2, 32)) { ++bad; printf("transcript after the tail differs\n"); }
    zko_transcript_free(otr);
    return bad;
}

// plain sumcheck (basic_sumcheck::Prover): the table absorb and the claimed sum on the host, the rounds in the tail
template <int FID> static int run_tail_plain(int n, int host_rounds, int nblocks = 1) {
    const uint64_t len = 1ull << n;
    zk::HostField f(FID);
    std::vector<uint64_t> table(len * 4);
    for (uint64_t i = 0; i < len; ++i) rand_fe(FID, &table[4 * i], (rnd() % 16 == 0) ? (int)(rnd() % 6) : 0);
    uint64_t claim[4], fin[4];
    std::vector<uint64_t> rp((size_t)n * 8), ch((size_t)n * 4);
    if (zko_basic_prove(FID, table.data(), len, claim, rp.data(), ch.data(), fin)) { printf("oracle refused\n"); return 1; }
    zk::HostTranscript tr;
    std::vector<uint8_t> bytes(len * 32);
    zko_mle_to_bytes(FID, table.data(), len, bytes.data());
    tr.append(bytes.data(), bytes.size());
    zk::HFe hclaim; memcpy(hclaim.l, claim, 32); tr.append_be(f, hclaim);
    int bad = 0;
    zk::HFe r = f.zero();
    std::vector<uint64_t> cur = table; uint64_t cur_len = len;
    for (int k = 0; k < host_rounds; ++k) {
        zk::HFe ev[2]; zko_split_and_sum(FID, cur.data(), cur_len, (uint64_t*)ev);
        uint8_t b[64]; f.to_bytes_be(ev[0], b); f.to_bytes_be(ev[1], b + 32); tr.append(b, 64);
        r = tr.challenge(f);
        if (memcmp(ev, &rp[(size_t)k * 8], 64) || memcmp(r.l, &ch[(size_t)k * 4], 32)) { ++bad; printf("plain host round %d mismatch\n", k); }
        if (k + 1 < host_rounds) {
            std::vector<uint64_t> nxt((cur_len / 2) * 4);
            zko_mle_partial_evaluate(FID, cur.data(), cur_len, 0, r.l, nxt.data());
            cur.swap(nxt); cur_len /= 2;
        }
    }
    std::vector<zk::Fe> tab(cur_len);
    for (uint64_t i = 0; i < cur_len; ++i) tab[i] = to_fe(&cur[4 * i]);
    zk::DevArgs a; memset(&a, 0, sizeof a);
    zk::DevOut out; memset(&out, 0, sizeof out);
    a.tp.t[0] = tab.data();
    a.log_len = 0; while ((1ull << a.log_len) < cur_len) ++a.log_len;
    a.pending = host_rounds > 0; a.mode = zk::kDevPlain; a.seq = 5; a.world = 1; a.final_fold = 1;
    a.max_rounds = a.log_len - (a.pending ? 1 : 0);
    if (a.pending) a.ft = host_fold_table(f, r);
    fill_tail_consts<FID>(a, f, 1);
    tr.export_state(a.sponge.s, &a.sponge.pos);
    a.out = &out;
    bad += emu_launch<FID, 1, 1, 0>(a, nblocks);
    if (out.seq != 5 || out.status != zk::kDevOk || (int)out.rounds != n - host_rounds) { ++bad; printf("plain tail rounds %u (want %d)\n", out.rounds, n - host_rounds); }
    for (int k = host_rounds; k < n; ++k) {
        if (memcmp(out.round_vals[k - host_rounds], &rp[(size_t)k * 8], 64)) { ++bad; printf("plain tail n=%d round %d sums mismatch\n", n, k); }
        if (memcmp(&out.challenges[k - host_rounds], &ch[(size_t)k * 4], 32)) { ++bad; printf("plain tail challenge %d mismatch\n", k); }
    }
    if (memcmp(&out.finals[0], fin, 32)) { ++bad; printf("plain tail final evaluation mismatch\n"); }
    return bad;
}

// Sharded product sumcheck: G ranks hold the low-index-bit shards of the tables (comm.cu), every rank runs the round loop
// on `nblocks` blocks and the leaders exchange their partial evaluations through each other's slot arrays; after
// `sharded_rounds` rounds the launch stops with the last challenge pending, the shards are folded, gathered and
// interleaved (what zk_prove_product_sharded's collapse does) and one more launch finishes on the full table.
template <int FID, int P, int D> static int run_sharded_product(int n, int g, int sharded_rounds, int nblocks) {
    constexpr int T = P * D, PO = P < 2 ? 2 : P, NE = D + 1;
    typedef zk::DevRounds<FID, P, D, 0, HostExec> Blk;
    const int G = 1 << g;
    const uint64_t len = 1ull << n, m = len >> g;
    zk::HostField f(FID);
    std::vector<uint64_t> ot((size_t)PO * D * len * 4, 0);
    for (int t = 0; t < T; ++t)
        for (uint64_t i = 0; i < len; ++i) rand_fe(FID, &ot[((size_t)t * len + i) * 4], (rnd() % 16 == 0) ? (int)(rnd() % 6) : 0);
    std::vector<uint64_t> red(len * 4); uint64_t claim[4];
    zko_sumpoly_reduce(FID, ot.data(), PO, D, len, red.data()); zko_fe_sum(FID, red.data(), len, claim);
    zko_transcript* otr = zko_transcript_new();
    std::vector<uint64_t> wc((size_t)n * NE * 4), wch((size_t)n * 4), wfin((size_t)PO * D * 4);
    if (zko_product_prove(FID, ot.data(), PO, D, len, claim, otr, wc.data(), wch.data(), wfin.data())) { printf("oracle refused\n"); return 1; }
    int bad = 0;
    // ---- stage 1: G ranks x nblocks blocks
    std::vector<std::vector<std::vector<zk::Fe>>> shard(G, std::vector<std::vector<zk::Fe>>(T, std::vector<zk::Fe>(m)));
    for (int q = 0; q < G; ++q)
        for (int t = 0; t < T; ++t)
            for (uint64_t j = 0; j < m; ++j) shard[q][t][j] = to_fe(&ot[((size_t)t * len + (j * G + q)) * 4]);
    std::vector<std::vector<zk::PeerSlot>> slots(G, std::vector<zk::PeerSlot>(2 * G));
    for (auto& v : slots) memset(v.data(), 0, v.size() * sizeof(zk::PeerSlot));
    std::vector<zk::DevArgs> args(G);
    std::vector<zk::DevOut> outs(G);
    std::vector<zk::DevGlobal> glob(G);
    std::vector<std::vector<zk::DevShared>> sh(G, std::vector<zk::DevShared>(nblocks));
    std::vector<std::vector<HostExec>> ex(G, std::vector<HostExec>(nblocks));
    std::vector<std::vector<Blk*>> blk(G, std::vector<Blk*>(nblocks));
    std::vector<std::vector<char>> alive(G, std::vector<char>(nblocks, 1));
    zk::HostTranscript tr0;
    zk::HFe hclaim; memcpy(hclaim.l, claim, 32); tr0.append_be(f, hclaim);
    for (int q = 0; q < G; ++q) {
        zk::DevArgs& a = args[q];
        memset(&a, 0, sizeof a); memset(&outs[q], 0, sizeof(zk::DevOut)); memset(&glob[q], 0, sizeof(zk::DevGlobal));
        for (int t = 0; t < T; ++t) a.tp.t[t] = shard[q][t].data();
        a.log_len = n - g; a.pending = 0; a.mode = zk::kDevProduct; a.seq = 11; a.max_rounds = sharded_rounds;
        a.world = G; a.rank = q; a.xseq = 1000;
        for (int r = 0; r < G; ++r) a.peers[r] = slots[r].data();
        fill_tail_consts<FID>(a, f, D);
        tr0.export_state(a.sponge.s, &a.sponge.pos);
        a.out = &outs[q]; a.g = &glob[q];
        for (int b = 0; b < nblocks; ++b) { ex[q][b].bid_ = b; ex[q][b].nblocks_ = nblocks; blk[q][b] = new Blk(a, sh[q][b], ex[q][b]); blk[q][b]->init(); }
    }
    for (int round = 0; round <= sharded_rounds; ++round) {
        for (int q = 0; q < G; ++q)
            for (int b = nblocks - 1; b >= 1; --b)
                if (alive[q][b]) alive[q][b] = blk[q][b]->step();
        if (!blk[0][0]->rounds_left()) {
            for (int q = 0; q < G; ++q) alive[q][0] = blk[q][0]->step();   // finish_launch
            break;
        }
        bool ok = true;
        for (int q = 0; q < G; ++q) ok = ok && blk[q][0]->step_compute();
        for (int q = 0; q < G; ++q) ok = ok && blk[q][0]->step_post();
        for (int q = 0; q < G; ++q) ok = ok && blk[q][0]->step_finish();
        if (!ok) { ++bad; printf("sharded emulation: a wait was not satisfied in round %d\n", round); break; }
        for (int q = 0; q < G; ++q) blk[q][0]->advance();
    }
    for (int q = 0; q < G; ++q) {
        for (int b = 0; b < nblocks; ++b) { if (alive[q][b]) { ++bad; printf("rank %d block %d still alive\n", q, b); } delete blk[q][b]; }
        if (outs[q].seq != 11 || outs[q].status != zk::kDevOk || (int)outs[q].rounds != sharded_rounds) { ++bad; printf("rank %d: rounds %u status %u\n", q, outs[q].rounds, outs[q].status); }
        for (int k = 0; k < sharded_rounds; ++k) {
            if (memcmp(outs[q].round_vals[k], &wc[(size_t)k * NE * 4], 32 * NE)) { ++bad; printf("sharded<%d,%d,%d> n=%d G=%d rank %d round %d coefficients mismatch\n", FID, P, D, n, G, q, k); }
            if (memcmp(&outs[q].challenges[k], &wch[(size_t)k * 4], 32)) { ++bad; printf("sharded rank %d round %d challenge mismatch\n", q, k); }
        }
    }
    // ---- collapse: fold the shards by the pending challenge, gather, interleave; stage 2 on one rank
    zk::HFe r; memcpy(r.l, &outs[0].challenges[sharded_rounds - 1], 32);
    uint64_t cur = m >> (sharded_rounds - 1);           // local length before the pending fold
    const zk::FoldTable ft = host_fold_table(f, r);
    std::vector<std::vector<zk::Fe>> full(T, std::vector<zk::Fe>((cur / 2) * G));
    for (int q = 0; q < G; ++q)
        for (int t = 0; t < T; ++t)
            for (uint64_t j = 0; j < cur / 2; ++j) {
                zk::Fe o;
                zk::FoldScalar<FID>::fold(o, shard[q][t][j], shard[q][t][j + cur / 2], ft);
                full[t][j * G + q] = o;
            }
    zk::DevArgs a; memset(&a, 0, sizeof a);
    zk::DevOut out; memset(&out, 0, sizeof out);
    for (int t = 0; t < T; ++t) a.tp.t[t] = full[t].data();
    const uint64_t flen = (cur / 2) * G;
    a.log_len = 0; while ((1ull << a.log_len) < flen) ++a.log_len;
    a.pending = 0; a.mode = zk::kDevProduct; a.seq = 12; a.world = 1; a.max_rounds = a.log_len; a.final_fold = 1;
    fill_tail_consts<FID>(a, f, D);
    memcpy(&a.sponge, &outs[G - 1].sponge, sizeof a.sponge);
    a.out = &out;
    bad += emu_launch<FID, P, D, 0>(a, nblocks);
    if ((int)out.rounds != n - sharded_rounds) { ++bad; printf("stage 2 rounds %u (want %d)\n", out.rounds, n - sharded_rounds); }
    for (int k = sharded_rounds; k < n; ++k)
        if (memcmp(out.round_vals[k - sharded_rounds], &wc[(size_t)k * NE * 4], 32 * NE) || memcmp(&out.challenges[k - sharded_rounds], &wch[(size_t)k * 4], 32)) { ++bad; printf("sharded stage 2 round %d mismatch\n", k); }
    for (int t = 0; t < T; ++t)
        if (memcmp(&out.finals[t], &wfin[(size_t)t * 4], 32)) { ++bad; printf("sharded final value %d mismatch\n", t); }
    zko_transcript_free(otr);
    return bad;
}

template <int FID> static int run_tails() {
    int bad = 0;
    for (int n = 1; n <= 7; ++n)
        for (int h = 0; h <= n && h <= 3; ++h) {
            bad += run_tail_product<FID, 1, 2, 0>(n, h, 3 * n + h);
            bad += run_tail_product<FID, 2, 2, 0>(n, h, 8 * h);
            bad += run_tail_product<FID, 1, 2, 1>(n, h, 0);
            bad += run_tail_plain<FID>(n, h);
        }
    bad += run_tail_product<FID, 2, 3, 0>(4, 1, 1);
    bad += run_tail_product<FID, 1, 3, 0>(3, 0, 2);
    bad += run_tail_product<FID, 4, 2, 0>(3, 2, 135);
    bad += run_tail_product<FID, 3, 2, 0>(5, 5, 136);
    // several blocks: the grid accumulator, arrivals / release, blocks leaving as the tables shrink
    for (int nb : {2, 3, 5, 8, 64})
        for (int n = 1; n <= 7; n += 2)
            for (int h = 0; h <= 2 && h <= n; ++h) {
                bad += run_tail_product<FID, 1, 2, 0>(n, h, n + h, nb);
                bad += run_tail_product<FID, 1, 2, 1>(n, h, 7, nb);
                bad += run_tail_plain<FID>(n, h, nb);
            }
    bad += run_tail_product<FID, 2, 2, 0>(6, 0, 0, 7);
    bad += run_tail_product<FID, 2, 3, 0>(5, 1, 1, 4);
    // several ranks exchanging partial evaluations through peer slots, then the collapse
    for (int g = 1; g <= 3; ++g)
        for (int nb : {1, 3})
            for (int sr = 1; sr <= 4; ++sr) {   // sr == 4: the sharded rounds exhaust the LOCAL tables (collapse_len == 1)
                bad += run_sharded_product<FID, 1, 2>(g + 4, g, sr, nb);
                if (g < 3) bad += run_sharded_product<FID, 2, 2>(g + 3, g, sr > 2 ? 2 : sr, nb);
            }
    printf("tail field %d: %s\n", FID, bad ? "FAIL" : "ok");
    return bad;
}

int main() {
    int bad = run<0>() + run<1>() + run<2>();
    bad += test_dev_sponge();
    bad += run_tails<0>() + run_tails<1>() + run_tails<2>();
    bad += run_round_accs<0>() + run_round_accs<1>() + run_round_accs<2>();
    return bad ? 1 : 0;
}
