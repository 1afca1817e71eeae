// internal.h -- engine.cu entry points shared with comm.cu
#pragma once
#include "../../include/zk_sumcheck.h"
#include "engine.h"
#include "kernels.cuh"
#include "tail.cuh"

int fail(zk_ctx* ctx, int code, const char* msg);
zk::FoldTable make_fold_table(const zk::HostField& f, const zk::HFe& r_mont);
const zk::Interpolator& interp_for(zk_ctx* ctx, int degree);
zk::TablePtrs ptrs_of(const zk_sumpoly* sp);
void set_len(zk_sumpoly* sp, uint64_t len);
int sync_len(zk_ctx* ctx, zk_sumpoly* sp);
int table_alloc(zk_ctx* ctx, uint64_t n, zk_table** out);
int ensure_scratch(zk_ctx* ctx, size_t bytes);
int ensure_pool(zk_ctx* ctx, size_t bytes);   // persistent table pool of the dense GKR prover
int tensor_into(zk_ctx* ctx, const zk::Fe* wb, const zk::Fe* wc, uint64_t n, zk::Fe* out, int op);
namespace zk {
int fetch_result(zk_ctx* ctx, HFe* out, int ne);
int launch_round_evals(zk_ctx* ctx, const TablePtrs& tp, int P, int D, uint64_t len, bool shared = false, int nlin = 0);
int launch_fold_evals(zk_ctx* ctx, const TablePtrs& tp, int P, int D, uint64_t len, const FoldTable& ft, bool skip1, bool shared = false, int nlin = 0);
// spin until `box` carries sequence number `seq` (watchdog: stream errors, 120 s wall clock)
int wait_mailbox(zk_ctx* ctx, const volatile Mailbox* box, unsigned seq, bool own);
int wait_seq(zk_ctx* ctx, const volatile unsigned* word, unsigned seq, bool own);
int launch_fold0(zk_ctx* ctx, const TablePtrs& tp, int ntables, uint64_t len, const FoldTable& ft);
// Device tail (tail.cuh): every remaining round of a sumcheck over T = P*D + nlin tables of `len` entries in ONE
// single-block launch, transcript included; `pending_r` (may be null) is a challenge the tables still have to be
// folded by.  mode: kTailProduct / kTailPlain.  vals_out receives (D+1) elements per round, chal_out (may be null)
// one; finals (may be null) the T single entries left.  The tables end with one entry each.
bool tail_applies(const zk_ctx* ctx, uint64_t len, int tables, uint32_t flags);
int run_tail(zk_ctx* ctx, const TablePtrs& tp, int P, int D, int nlin, int mode, uint64_t len, const HFe* pending_r, HostTranscript& tr,
             uint64_t* vals_out, uint64_t* chal_out, uint64_t* finals);
}
