"""world_size-2 (and 4) gloo runs of the sharded product-sumcheck round loop on CPU.

The loop under test is the product's own `sharded.prove_product`; the per-rank table work is done by an
oracle-backed engine (tests may use the oracle), so what is covered here is the N>1 host logic: the
low-bit shard layout, the field reduction of the all-gathered partial evaluations, the transcript
replicated on every rank, and the collapse (all-gather + interleave) to an unsharded table."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleShardEngine:
    def __init__(self, co, fid, shard_tables):
        self.co, self.fid = co, fid
        self.P, self.D = shard_tables.shape[0], shard_tables.shape[1]
        self.t = [np.ascontiguousarray(shard_tables[p, d]) for p in range(self.P) for d in range(self.D)]

    def local_len(self):
        return self.t[0].shape[0]

    def _stack(self):
        return np.stack(self.t).reshape(self.P, self.D, self.local_len(), 4)

    def round_evals(self):
        return self.co.generate_round_univariate(self.fid, self._stack())

    def fold(self, r):
        self.t = [self.co.mle_partial_evaluate(self.fid, t, 0, r) for t in self.t]

    def fold_and_evals(self, r):
        self.fold(r)
        return self.round_evals()

    def tables(self):
        return [t.copy() for t in self.t]

    def replace_tables(self, tables):
        self.t = [np.ascontiguousarray(t) for t in tables]


def _worker(rank, world, port, fid, n, P, D, collapse_len, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import coracle as co
    from zk_cryptography_research_implementations_b200 import sharded
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(1234)
        raw = rng.integers(0, 1 << 62, size=(P, D, 1 << n, 4), dtype=np.uint64)
        raw[..., 3] &= np.uint64((1 << 58) - 1)          # < p for all three fields: canonical limbs
        full = raw
        claimed = np.zeros(4, dtype=np.uint64)
        co.lib().zko_fe_sum(fid, co._p(co.sumpoly_reduce(fid, full)), 1 << n, co._p(claimed))
        want = co.product_prove(fid, full, claimed, co.Transcript())
        shard = np.stack([np.stack([sharded.shard_of(full[p, d], rank, world) for d in range(D)]) for p in range(P)])
        eng = OracleShardEngine(co, fid, shard)
        got = sharded.prove_product(eng, fid, n, claimed, co.Transcript(), collapse_len=collapse_len)
        ok = all(np.array_equal(a, b) for a, b in zip(got[:2], want[:2])) and np.array_equal(got[2].reshape(P, D, 4), want[2])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,P,D,collapse_len", [(2, 5, 2, 2, 1), (2, 6, 2, 2, 4), (4, 5, 2, 3, 2), (2, 1, 2, 2, 1), (4, 2, 2, 2, 8)])
def test_sharded_round_loop_matches_unsharded_oracle(world, n, P, D, collapse_len):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 0, n, P, D, collapse_len, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


def test_shard_layout_roundtrip():
    from zk_cryptography_research_implementations_b200 import sharded
    t = np.arange(64 * 4, dtype=np.uint64).reshape(64, 4)
    for G in (1, 2, 4, 8):
        shards = [sharded.shard_of(t, q, G) for q in range(G)]
        assert all(s.shape[0] == 64 // G for s in shards)
        assert np.array_equal(shards[G - 1][1], t[G + G - 1])
        assert np.array_equal(sharded.interleave(shards), t)
