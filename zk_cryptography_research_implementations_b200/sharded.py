"""One process per GPU: product sumcheck over tables sharded on the low index bits.

Two drivers over the same layout:

  * `prove_product_native` -- the whole round loop in C++ (csrc/comm.cu), NCCL all-gather of the d+1
    partial evaluations per round; this is what bench.py times.
  * `prove_product` -- the round loop here, transcript on the caller's side, `torch.distributed`
    all-gather for the exchange (NCCL on GPUs, gloo on CPU).  The per-rank work goes through a small
    engine interface; the product engine is `CudaShardEngine`.  tests/test_sharded_gloo.py drives the
    same loop with an oracle-backed engine on CPU to cover the layout, the field reduction of the
    partials, the replicated transcript and the collapse/interleave step without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .core import Context, DeviceTable, ReferencePanic, _ptr, as_elems, fe_binop
from .transcripts import Transcript


# ------------------------------------------------------------------ layout
def shard_of(global_table: np.ndarray, rank: int, world: int) -> np.ndarray:
    """rank q holds global entries q, q + G, q + 2G, ... (the variables bound LAST are the shard index)"""
    return np.ascontiguousarray(global_table[rank::world])


def interleave(shards: Sequence[np.ndarray]) -> np.ndarray:
    """inverse of shard_of: out[j * G + q] = shards[q][j]"""
    G = len(shards)
    m = shards[0].shape[0]
    out = np.zeros((m * G,) + shards[0].shape[1:], dtype=shards[0].dtype)
    for q, s in enumerate(shards):
        out[q::G] = s
    return out


# ------------------------------------------------------------------ communicator bootstrap
def init_comm(ctx: Context, group=None, shared_mailboxes: bool = True) -> None:
    """create the library's NCCL communicator for this process group (id broadcast through torch)"""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    buf = C.create_string_buffer(128)
    if rank == 0:
        if ctx.lib.zk_comm_unique_id(buf) != 0:
            raise RuntimeError("ncclGetUniqueId failed")
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0, group=group)
    ident = bytes(t.cpu().tolist())
    ctx.check(ctx.lib.zk_comm_init(ctx.h, rank, world, ident))
    if world > 1 and shared_mailboxes:
        # per-round exchange through a shared-memory segment all rank processes (one node) map
        name = [("/zkb200_%d_%s" % (os.getpid(), os.urandom(4).hex())) if rank == 0 else None]
        dist.broadcast_object_list(name, 0, group=group)
        if rank == 0:
            ctx.check(ctx.lib.zk_comm_attach_mailboxes(ctx.h, name[0].encode(), 1))
        dist.barrier(group=group)
        if rank != 0:
            ctx.check(ctx.lib.zk_comm_attach_mailboxes(ctx.h, name[0].encode(), 0))
        dist.barrier(group=group)
        if rank == 0:
            ctx.lib.zk_comm_unlink_mailboxes(name[0].encode())


def exchange_kind(ctx: Context) -> str:
    """how the sharded provers exchange a round's partial evaluations on this communicator"""
    if ctx.lib.zk_comm_world(ctx.h) < 2:
        return "none (one rank)"
    if ctx.lib.zk_comm_peer_exchange(ctx.h):
        return ("in-kernel over peer memory (NVLink): each rank's persistent round-loop kernel writes its (d+1) x 32 B partial evaluations "
                "into every peer's slot and runs the transcript itself; host mailboxes only for rounds above the hand-over size")
    return "shared-memory mailboxes written by the round kernels, summed on the host (no per-round collective)"


def prove_product_native(ctx: Context, sp_handle, P: int, D: int, n_global: int, claimed_sum, transcript: Transcript,
                         flags: int = 0, collapse_len: int = 1 << 12):
    coeffs = np.zeros((max(n_global, 1), D + 1, 4), dtype=np.uint64)
    chal = np.zeros((max(n_global, 1), 4), dtype=np.uint64)
    fin = np.zeros((P * D, 4), dtype=np.uint64)
    claimed = as_elems(claimed_sum).copy()
    ctx.check(ctx.lib.zk_prove_product_sharded(ctx.h, sp_handle, _ptr(claimed), transcript.h, _ptr(coeffs), _ptr(chal),
                                               _ptr(fin), flags, collapse_len))
    return coeffs[:n_global], chal[:n_global], fin


# ------------------------------------------------------------------ per-rank engines
class CudaShardEngine:
    """this rank's shard of the P*D tables on its GPU, stepped with the C-ABI round primitives"""

    def __init__(self, ctx: Context, shard_tables: np.ndarray):
        # shard_tables: (P, D, m, 4)
        self.ctx = ctx
        self.P, self.D = shard_tables.shape[0], shard_tables.shape[1]
        self._load([shard_tables[p, d] for p in range(self.P) for d in range(self.D)])

    def _load(self, tables: List[np.ndarray]) -> None:
        ctx = self.ctx
        tabs = [ctx.upload(t) for t in tables]
        arr = (C.c_void_p * len(tabs))(*[t.release() for t in tabs])
        h = C.c_void_p()
        ctx.check(ctx.lib.zk_sumpoly_create(ctx.h, arr, self.P, self.D, C.byref(h)))
        self.h = h

    def local_len(self) -> int:
        return int(self.ctx.lib.zk_sumpoly_len(self.h))

    def round_evals(self) -> np.ndarray:
        out = np.zeros((self.D + 1, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.zk_sumcheck_round_evals(self.ctx.h, self.h, _ptr(out)))
        return out

    def fold_and_evals(self, r) -> np.ndarray:
        out = np.zeros((self.D + 1, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.zk_sumcheck_fold_and_evals(self.ctx.h, self.h, _ptr(as_elems(r)), _ptr(out)))
        return out

    def fold(self, r) -> None:
        self.ctx.check(self.ctx.lib.zk_sumcheck_fold_and_evals(self.ctx.h, self.h, _ptr(as_elems(r)), None))

    def tables(self) -> List[np.ndarray]:
        res = []
        for i in range(self.P * self.D):
            th = self.ctx.lib.zk_sumpoly_table(self.h, i)
            out = np.zeros((self.local_len(), 4), dtype=np.uint64)
            self.ctx.check(self.ctx.lib.zk_table_download(self.ctx.h, th, _ptr(out)))
            res.append(out)
        return res

    def replace_tables(self, tables: List[np.ndarray]) -> None:
        self.close()
        self._load(tables)

    def close(self) -> None:
        if getattr(self, "h", None):
            self.ctx.lib.zk_sumpoly_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------ the round loop with the exchange
def _all_gather_elems(local: np.ndarray, group, device: str) -> List[np.ndarray]:
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.from_numpy(local.view(np.int64).copy()).to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return [o.cpu().numpy().view(np.uint64) for o in outs]


def _field_sum(field: int, parts: Sequence[np.ndarray]) -> np.ndarray:
    acc = parts[0].copy()
    for p in parts[1:]:
        for e in range(acc.shape[0]):
            acc[e] = fe_binop("add", field, acc[e], p[e])
    return acc


def prove_product(engine, field: int, n_global: int, claimed_sum, transcript, group=None, collapse_len: int = 1,
                  derive_s1: bool = True):
    """`sumcheck_gkr_protocol::prove` (sumcheck_gkr_protocol.rs:24-67) over sharded tables.
    `transcript` needs append(bytes) and random_challenge_as_field_element(field).
    Returns (coeffs (n, D+1, 4), challenges (n, 4), final table values (P*D, 4)); identical on all ranks."""
    import torch.distributed as dist
    lib = _lib.load()
    world = dist.get_world_size(group)
    device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    D = engine.D
    NE = D + 1
    from .core import fe_to_ints
    transcript.append(fe_to_ints(field, claimed_sum)[0].to_bytes(32, "big"))           # :35
    coeffs_all = np.zeros((n_global, NE, 4), dtype=np.uint64)
    chal_all = np.zeros((n_global, 4), dtype=np.uint64)
    running = as_elems(claimed_sum).copy()
    sharded = world > 1
    r = None
    for k in range(n_global):
        plain = (k == 0)
        if sharded and ((k == 0 and engine.local_len() <= collapse_len) or (k > 0 and engine.local_len() // 2 <= collapse_len)):
            if k > 0:
                engine.fold(r)
            gathered = [_all_gather_elems(t, group, device) for t in engine.tables()]
            engine.replace_tables([interleave(parts) for parts in gathered])
            sharded = False
            plain = True
        ev = engine.round_evals() if plain else engine.fold_and_evals(r)
        if sharded:
            ev = _field_sum(field, _all_gather_elems(ev, group, device))
        if derive_s1 and k > 0:
            assert np.array_equal(fe_binop("add", field, ev[0], ev[1]), running), "round sums do not telescope"
        c = np.zeros((NE, 4), dtype=np.uint64)
        lib.zk_interpolate_evals(field, NE, _ptr(np.ascontiguousarray(ev)), _ptr(c))   # :46-50
        transcript.append(b"".join(v.to_bytes(32, "little") for v in fe_to_ints(field, c)))   # :52
        r = transcript.random_challenge_as_field_element(field)                        # :55
        lib.zk_univariate_evaluate(field, _ptr(c), NE, _ptr(as_elems(r)), _ptr(running))
        coeffs_all[k] = c
        chal_all[k] = r
    if sharded:
        raise AssertionError("tables still sharded after the last round")
    if n_global > 0:
        engine.fold(r)                                                                  # :57 of the last round
    fin = np.stack([t[0] for t in engine.tables()])
    return coeffs_all, chal_all, fin
