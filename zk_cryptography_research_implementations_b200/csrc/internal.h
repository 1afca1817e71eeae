// internal.h -- engine.cu entry points shared with comm.cu
#pragma once
#include "../../include/zk_sumcheck.h"
#include "engine.h"
#include "kernels.cuh"
#include "devrounds.cuh"

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
int launch_round_evals(zk_ctx* ctx, const TablePtrs& tp, int P, int D, uint64_t len, bool shared = false, int nlin = 0, bool skip1 = false);
// the shapes for which a round-0 kernel without s(1) exists (ZK_FLAG_TRUSTED_CLAIM)
inline bool round_evals_skip1_supported(int P, int D, int nlin) { return P == 1 && D == 2 && nlin == 1; }
int launch_fold_evals(zk_ctx* ctx, const TablePtrs& tp, int P, int D, uint64_t len, const FoldTable& ft, bool skip1, bool shared = false, int nlin = 0);
// spin until `box` carries sequence number `seq` (watchdog: stream errors, 120 s wall clock)
int wait_mailbox(zk_ctx* ctx, const volatile Mailbox* box, unsigned seq, bool own);
int wait_seq(zk_ctx* ctx, const volatile unsigned* word, unsigned seq, bool own);
int launch_fold0(zk_ctx* ctx, const TablePtrs& tp, int ntables, uint64_t len, const FoldTable& ft);
// Device-resident rounds (devrounds.cuh): the remaining rounds of a sumcheck over T = P*D + nlin tables of `len` entries
// in ONE persistent launch, transcript included; `pending_r` (may be null) is a challenge the tables still have to be
// folded by.  mode: kDevProduct / kDevPlain.  max_rounds == 0: run to the end (the tables end with one entry each and
// `finals`, if not null, receives them); otherwise stop after max_rounds rounds with the last challenge left pending.
// sharded: the tables are this rank's shard and the partial evaluations are exchanged between the ranks' kernels over
// peer memory (comm.cu must have attached the peers).  vals_out receives (D+1) elements per round, chal_out (may be
// null) one; *rounds_run (may be null) the number of rounds the launch ran.
// comm.cu: recv[q * bytes ..] = rank q's `send` (host buffers), over the communicator of zk_comm_init; world == 1 copies
int allgather_host_bytes(zk_ctx* ctx, const void* send, size_t bytes, void* recv);
bool dev_rounds_apply(const zk_ctx* ctx, uint64_t len, int tables, uint32_t flags);
int run_dev_rounds(zk_ctx* ctx, const TablePtrs& tp, int P, int D, int nlin, int mode, uint64_t len, const HFe* pending_r, HostTranscript& tr,
                   uint64_t* vals_out, uint64_t* chal_out, uint64_t* finals, uint32_t max_rounds = 0, bool sharded = false,
                   uint32_t* rounds_run = nullptr);
}
