// engine.h -- the opaque handle types of include/zk_sumcheck.h.
#pragma once
#include <cuda_runtime.h>
#include <functional>
#include <map>
#include <string>
#include <vector>
#include "host_field.h"

namespace zk { struct Fe; struct Mailbox; struct DevOut; struct DevGlobal; struct PeerSlot; }

struct zk_ctx {
    explicit zk_ctx(int field_id) : fid(field_id), field(field_id) {}
    int fid;
    int device = 0;
    int sm_count = 148;
    int max_grid = 148 * 8;
    int grid_cap = 0;   // ZKB200_GRID_CAP: test hook, caps every grid (forces long grid-stride loops)
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    zk::HostField field;
    std::map<int, zk::Interpolator> interps;   // one inverse Vandermonde per degree
    // grid-wide reduction scratch (column totals, kernels.cuh) + the published round evaluations (mapped pinned host memory)
    unsigned long long* gacc = nullptr;
    unsigned* ticket = nullptr;
    zk::Mailbox* mail_host = nullptr;   // this context's own mailbox (unsharded operations)
    zk::Mailbox* mail_dev = nullptr;
    unsigned mail_seq = 0;
    // sharded runs: mailboxes [2][world] in a POSIX shared-memory segment mapped by every rank process
    zk::Mailbox* xmail_host = nullptr;
    zk::Mailbox* xmail_dev = nullptr;
    size_t xmail_bytes = 0;
    unsigned xmail_seq = 0;
    bool exchange_pending = false;      // the last round kernel published into the shared mailbox
    // device-resident rounds (devrounds.cuh): all remaining rounds in one persistent launch once tables x entries <= 2^tail_log
    zk::DevOut* dev_host = nullptr;     // mapped pinned host memory
    zk::DevOut* dev_dev = nullptr;
    zk::DevGlobal* dev_global = nullptr;   // device: grid accumulator, arrival / release words, the current fold table
    bool dev_global_dirty = false;      // a launch failed: re-zero before the next one
    unsigned dev_seq = 0;
    int tail_log = 20;                  // ZKB200_TAIL_LOG / zk_ctx_set_tail_log; 0 = always host-driven rounds
    std::map<int, int> dev_capacity;    // co-resident blocks of the round-loop kernel per (P, D, nlin)
    // Called right after the persistent round-loop launch has been queued (argument: where that launch's first challenge
    // will be stored in the caller's challenge array).  The GKR prover uses it to queue work that only needs the challenges
    // known so far on `side_stream`: it backfills the SMs the shrinking round loop leaves idle.
    std::function<void(const uint64_t*)> dev_hook;
    cudaStream_t side_stream = nullptr;   // created on first use
    cudaEvent_t side_event = nullptr;
    zk::HFe pow32[8];                   // Montgomery forms of 2^(32 i)
    // general scratch (evaluate / convert_to_bytes / out-of-place folds)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    void* pinned = nullptr;
    // table pool of the dense GKR prover (gkr.cu): kept across proves -- a 2 GiB cudaMalloc/cudaFree per call costs
    // milliseconds and varies from call to call
    void* pool = nullptr;
    size_t pool_bytes = 0;
    // accounting
    bool profiling = false;
    uint64_t launches = 0, round_launches = 0;
    double round_ms = 0, round_bytes = 0;
    std::vector<cudaEvent_t> events;
    size_t ev_used = 0;
    std::string err;
    // one-process-per-GPU sharding (comm.cu): NCCL communicator + exchange buffers
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
    zk::Fe* xchg_send = nullptr;   // device, kMaxEvals elements
    zk::Fe* xchg_recv = nullptr;   // device, world * kMaxEvals elements
    zk::Fe* xchg_host = nullptr;   // pinned, world * kMaxEvals elements
    // in-kernel exchange of the device-resident rounds: every rank's PeerSlot[2][world] array, peer-mapped (cudaIpc)
    zk::PeerSlot* peer_slots[16] = {nullptr};   // [q]: rank q's array as mapped in this process ([rank]: our own allocation)
    bool peers_attached = false;
    unsigned xseq = 0;             // exchange sequence number, re-agreed by all ranks at the start of every sharded prove
};

struct zk_table {
    zk::Fe* d = nullptr;
    uint64_t len = 0, cap = 0;
    bool owned = true;
};

struct zk_sumpoly {
    std::vector<zk_table*> tabs;   // [p * D + d]
    uint32_t P = 0, D = 0;
    uint32_t nlin = 0;             // trailing tables that enter the sum linearly (internal: sparse GKR phases)
    uint64_t len = 0;
};

struct zk_transcript {
    zk::HostTranscript t;
};
