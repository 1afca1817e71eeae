"""2, 4 and 8 ranks, one GPU each (skipped with fewer devices): the native sharded provers (csrc/comm.cu) with every
exchange path -- in-kernel over peer memory (the device-resident round loop, default), host mailboxes, ncclAllGather --
and the Python round loop over torch.distributed, all against the unsharded oracle, limb for limb."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ctypes as C
    import torch
    import torch.distributed as dist
    import coracle as co
    import zk_cryptography_research_implementations_b200 as zk
    from zk_cryptography_research_implementations_b200 import sharded
    from zk_cryptography_research_implementations_b200.core import _ptr
    from zk_cryptography_research_implementations_b200.transcripts import Transcript
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    ok = True
    notes = []
    try:
        fid = 0
        ctx = zk.Context(fid, rank)
        sharded.init_comm(ctx)
        peer = bool(ctx.lib.zk_comm_peer_exchange(ctx.h))
        for (P, D, n, collapse) in [(2, 2, 9, 1), (2, 2, 12, 64), (1, 2, 14, 256), (2, 3, 8, 4), (1, 2, 20, 4096), (1, 2, 18, 1)]:
            N = 1 << n
            if N // world < 2:
                continue
            co.set_threads(os.cpu_count() or 1)
            full = np.stack([np.stack([co.table_generate(fid, 5, p * D + d, N) for d in range(D)]) for p in range(P)])
            Pref = max(P, 2)
            ref_tabs = np.zeros((Pref, D, N, 4), dtype=np.uint64)
            ref_tabs[:P] = full
            claimed = np.zeros(4, dtype=np.uint64)
            co.lib().zko_fe_sum(fid, co._p(co.sumpoly_reduce(fid, ref_tabs)), N, co._p(claimed))
            want = co.product_prove(fid, ref_tabs, claimed, co.Transcript())
            co.set_threads(1)
            # native driver, shard generated on the device; flags: 0 = default exchange (in-kernel over peer memory when the
            # peers could be mapped), 16 = host-driven rounds with the shared mailboxes, 4 = ncclAllGather per round,
            # 1 = s(1) summed directly, 64 = host exchange for the sharded rounds, device rounds after the collapse
            for flags in (0, 16, 4, 1, 64):
                tabs = [ctx.generate(5, i, N // world, first=rank, step=world) for i in range(P * D)]
                arr = (C.c_void_p * (P * D))(*[t.release() for t in tabs])
                sp = C.c_void_p()
                ctx.check(ctx.lib.zk_sumpoly_create(ctx.h, arr, P, D, C.byref(sp)))
                tr = Transcript()
                got = sharded.prove_product_native(ctx, sp, P, D, n, claimed, tr, flags=flags, collapse_len=collapse)
                ctx.lib.zk_sumpoly_free(ctx.h, sp)
                same = np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2].reshape(P, D, 4), want[2][:P])
                if not same:
                    notes.append("product (P,D,n,collapse)=%s flags=%d differs from the oracle" % ((P, D, n, collapse), flags))
                ok &= same
            if n > 14:
                continue
            # Python round loop over torch.distributed (NCCL), CUDA engine
            shard = np.stack([np.stack([sharded.shard_of(full[p, d], rank, world) for d in range(D)]) for p in range(P)])
            eng = sharded.CudaShardEngine(ctx, shard)
            got2 = sharded.prove_product(eng, fid, n, claimed, Transcript(), collapse_len=collapse)
            eng.close()
            ok &= np.array_equal(got2[0], want[0]) and np.array_equal(got2[1], want[1])
            # sharded MLE evaluate
            local = ctx.generate(5, 0, N // world, first=rank, step=world)
            out = np.zeros(4, dtype=np.uint64)
            ctx.check(ctx.lib.zk_mle_evaluate_sharded(ctx.h, local.h, _ptr(np.ascontiguousarray(want[1])), n, _ptr(out)))
            ok &= np.array_equal(out, co.mle_evaluate(fid, full[0, 0], want[1]))
        # plain sumcheck over a sharded table: the caller absorbs the table, the rest must equal the reference proof
        for (n, collapse) in [(11, 1), (13, 128), (19, 2048)]:
            N = 1 << n
            full = co.table_generate(fid, 6, 0, N)
            claimed_w, rp_w, ch_w, fin_w = co.basic_prove(fid, full)
            for flags in (0, 16, 4):
                local = ctx.generate(6, 0, N // world, first=rank, step=world)
                tr = Transcript()
                tr.append(co.mle_to_bytes(fid, full))
                claimed = np.zeros(4, dtype=np.uint64); rp = np.zeros((n, 2, 4), dtype=np.uint64)
                ch = np.zeros((n, 4), dtype=np.uint64); fin = np.zeros(4, dtype=np.uint64)
                ctx.check(ctx.lib.zk_prove_basic_sharded(ctx.h, local.h, tr.h, _ptr(claimed), _ptr(rp), _ptr(ch), _ptr(fin), flags, collapse))
                same = np.array_equal(claimed, claimed_w) and np.array_equal(rp, rp_w) and np.array_equal(ch, ch_w) and np.array_equal(fin, fin_w)
                if not same:
                    notes.append("plain (n,collapse)=%s flags=%d differs from the oracle" % ((n, collapse), flags))
                ok &= same
        # GKR with every layer's phase tables and sumchecks spread over the ranks (gkr_protocol.rs:26-143): every
        # rank must return the gate-list oracle's proof, limb for limb, on every exchange path
        from zk_cryptography_research_implementations_b200 import gkr
        for (w, depth, collapse) in [(10, 3, 16), (13, 3, 256)]:
            rng = np.random.default_rng(40 + w)
            n = 1 << w
            g = np.arange(n, dtype=np.int64)
            layers = [np.stack([g, rng.integers(0, n, size=n), g & 1, rng.integers(0, 2, size=n)], axis=1)]
            for _ in range(depth - 1):
                layers.append(np.stack([rng.integers(0, n, size=n), rng.integers(0, n, size=n), g, rng.integers(0, 2, size=n)], axis=1))
            bits = [1] + [w] * depth
            dev_in = ctx.generate(9, 0, n)
            inputs = dev_in.download()
            sc = co.SparseCircuit(bits, layers)
            want = co.gkr_prove_sparse(fid, sc, inputs)
            wc = gkr.WideCircuit(ctx, bits, layers)
            for flags in (0, 16, 4, "peer"):
                if flags == "peer":          # the sharded phase rounds inside the persistent kernels, NVLink exchange
                    os.environ["ZKB200_GKR_PEER_EXCHANGE"] = "1"
                    flags = 0
                proof = gkr.prove_wide(ctx, wc, dev_in, flags=flags, sharded=True, collapse_len=collapse)
                os.environ.pop("ZKB200_GKR_PEER_EXCHANGE", None)
                got = np.concatenate([np.stack([p_.coefficients for p_ in sp_.round_univariate_polynomials]) for sp_ in proof.sumcheck_proofs])
                same = (np.array_equal(got, want.coeffs[: got.shape[0]]) and np.array_equal(proof.claimed_sum, want.claimed_sum)
                        and np.array_equal(proof.wb_evaluations, want.wb[: depth - 1]) and np.array_equal(proof.wc_evaluations, want.wc[: depth - 1])
                        and np.array_equal(np.concatenate([sp_.random_challenges for sp_ in proof.sumcheck_proofs]), want.challenges[: got.shape[0]]))
                if not same:
                    notes.append("sharded GKR w=%d flags=%d differs from the oracle" % (w, flags))
                ok &= same
                ok &= bool(gkr.verify_wide(ctx, wc, proof, dev_in))
            wc.close()
        q.put((rank, bool(ok), peer, notes))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, False, None, [repr(e), traceback.format_exc()[-1500:]]))
    finally:
        dist.destroy_process_group()


def _kzg_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import coracle as co
    import zk_cryptography_research_implementations_b200 as zk
    from zk_cryptography_research_implementations_b200 import gkr, sharded
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    ok = True
    notes = []
    try:
        # multilinear KZG and succinct GKR over the ranks (BLS12-381 Fr context): setup, polynomial and
        # circuit replicated, every rank sums its share of the points; all ranks must return the oracle's points
        from zk_cryptography_research_implementations_b200.multilinear_kzg import MultilinearKZG, TrustedSetup
        fr = 2
        ctx2 = zk.Context(fr, rank)
        sharded.init_comm(ctx2)
        co.set_threads(os.cpu_count() or 1)
        for nv in (6, 12, 18):
            taus = ctx2.generate(31, 98, 64).download()[:nv].copy()
            point = ctx2.generate(31, 99, 64).download()[:nv].copy()
            setup = TrustedSetup.initialize_setup(ctx2, taus)
            table = ctx2.generate(31, 0, 1 << nv)
            c = MultilinearKZG.commit_to_polynomial(table, setup, sharded=True)
            pr = MultilinearKZG.open_and_prove(table, setup, point, sharded=True)
            if nv <= 12:
                g1 = co.kzg_setup_g1(taus)
                vals = table.download()
                ev, prs = co.kzg_open(vals, g1, point)
                same = np.array_equal(c, co.kzg_commit(vals, g1)) and np.array_equal(pr.evaluation, ev) and np.array_equal(pr.proofs, prs)
            else:       # past what the oracle opens in seconds: equal to the single-GPU calls, and the verification equation
                c1 = MultilinearKZG.commit_to_polynomial(table, setup)
                pr1 = MultilinearKZG.open_and_prove(table, setup, point)
                same = (np.array_equal(c, c1) and np.array_equal(pr.proofs, pr1.proofs) and np.array_equal(pr.evaluation, pr1.evaluation)
                        and co.kzg_verify_trapdoor(taus, c, point, pr.evaluation, pr.proofs))
            if not same:
                notes.append("sharded KZG n=%d differs" % nv)
            ok &= bool(same)
            setup.release()
        # succinct GKR, sharded: GKR part against the gate-list oracle, KZG part against the KZG oracle, verify_succinct accepts
        w, depth = 10, 3
        rng = np.random.default_rng(77)
        n = 1 << w
        g = np.arange(n, dtype=np.int64)
        layers = [np.stack([g, rng.integers(0, n, size=n), g & 1, rng.integers(0, 2, size=n)], axis=1)]
        for _ in range(depth - 1):
            layers.append(np.stack([rng.integers(0, n, size=n), rng.integers(0, n, size=n), g, rng.integers(0, 2, size=n)], axis=1))
        bits = [1] + [w] * depth
        dev_in = ctx2.generate(33, 0, n)
        inputs = dev_in.download()
        taus = ctx2.generate(33, 98, 64).download()[:w].copy()
        setup = TrustedSetup.initialize_setup(ctx2, taus)
        wc = gkr.WideCircuit(ctx2, bits, layers)
        proof = gkr.prove_succinct(ctx2, wc, dev_in, setup, sharded=True, collapse_len=64)
        want = co.gkr_prove_sparse(fr, co.SparseCircuit(bits, layers), inputs)
        got = np.concatenate([np.stack([p_.coefficients for p_ in sp_.round_univariate_polynomials]) for sp_ in proof.sumcheck_proofs])
        g1 = co.kzg_setup_g1(taus)
        chal = proof.sumcheck_proofs[-1].random_challenges
        ev_b, prs_b = co.kzg_open(inputs, g1, chal[:w])
        ev_c, prs_c = co.kzg_open(inputs, g1, chal[w:])
        same = (np.array_equal(got, want.coeffs[: got.shape[0]]) and np.array_equal(proof.claimed_sum, want.claimed_sum)
                and np.array_equal(proof.input_polynomial_commitment, co.kzg_commit(inputs, g1))
                and np.array_equal(proof.input_rb_proof.proofs, prs_b) and np.array_equal(proof.input_rc_proof.proofs, prs_c)
                and np.array_equal(proof.input_rb_proof.evaluation, ev_b) and np.array_equal(proof.input_rc_proof.evaluation, ev_c))
        if not same:
            notes.append("sharded succinct GKR differs from the oracle")
        ok &= bool(same)
        ok &= bool(gkr.verify_succinct(ctx2, wc, proof, setup, bind_input_openings=True))
        co.set_threads(1)
        wc.close()
        setup.release()
        q.put((rank, bool(ok), None, notes))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, False, None, [repr(e), traceback.format_exc()[-1500:]]))
    finally:
        dist.destroy_process_group()


def _run_ranks(worker, world, budget_s):
    import queue as _queue
    import time as _time
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res, deadline = [], _time.time() + budget_s
    while len(res) < world and _time.time() < deadline:
        try:
            res.append(q.get(timeout=5))
        except _queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):      # a rank died (its peers would wait in a collective forever)
                break
    for p in procs:
        p.join(timeout=10 if len(res) == world else 1)
        if p.is_alive():
            p.kill()
    assert len(res) == world, "rank exit codes: %s" % [p.exitcode for p in procs]
    assert sorted((r[0], r[1]) for r in res) == [(r, True) for r in range(world)], [r[3] for r in res]
    return res


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_kzg_and_succinct_gkr_match_the_oracle(world):
    _run_ranks(_kzg_worker, world, 600)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_provers_match_the_oracle(world):
    res = _run_ranks(_worker, world, 900)
    print("peer-memory exchange attached:", [r[2] for r in res])
