#!/usr/bin/env python3
"""bench.py -- sumcheck prove throughput (field elements / s) on B200, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--log2 n]

A "step" is one complete sumcheck prove (all rounds, host Fiat-Shamir included) over one set of
synthetic tables.  Default workload (every N): BASELINE.json configs[2], the configuration the
north-star targets are quoted on -- degree-2 product sumcheck f*g over 2^30 entries (BN254 Fq,
64 GiB of tables), strong scaling over the ranks (tables sharded on the low index bits, one tiny
NCCL all-gather per round).  `--workload plain24` is configs[1] (plain sumcheck, 2^24, BLS12-381 Fr).

Printed by rank 0: ONE JSON line (contract in the task statement) with `roofline` (dominant kernel vs
the measured HBM peak), `cpu_baseline` (the oracle's single-thread restatement of the reference prover
on a bounded sample), `e2e` (same prove through the C-ABI from pinned HOST tables, copies inside the
timed region), `gpu_launches`, `clocks`.

`--impl reference` times the reference's own algorithm on the host: the reference is single-threaded
Rust that cannot be compiled in this image, so this runs the oracle's C restatement (oracle/zkoracle.c,
reference pass structure, 1 thread) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 0xB200
MEGA = 1e6      # BASELINE.json quotes the sumcheck metric in Melems/s
WORKLOADS = {
    # name: (field id, field name, P, D, default log2 N, description)
    "product30": (0, "BN254_FQ", 1, 2, 30, "degree-2 product sumcheck f*g (BASELINE.json configs[2])"),
    "plain24": (2, "BLS12_381_FR", 1, 1, 24, "plain sumcheck of one MLE (BASELINE.json configs[1])"),
    "plain32": (2, "BLS12_381_FR", 1, 1, 32, "plain sumcheck of one 2^32-entry MLE (128 GiB), rounds only -- the upper end of BASELINE's 2^24-2^32 range"),
    "gkr22": (0, "BN254_FQ", 2, 2, 22, "GKR-shaped 2x2 sumcheck add*(Wb+Wc)+mul*(Wb*Wc) tables"),
    # --log2 = log2 of the layer width; 16 layers of 2^log2 gates, sparse two-phase layer prover
    "gkr_wide": (0, "BN254_FQ", 2, 2, 22, "GKR prove of a synthetic layered add/mul circuit, depth 16, width 2^22 gates (BASELINE.json configs[3])"),
    "mle": (0, "BN254_FQ", 1, 1, 28, "MultilinearPolynomial::evaluate / partial_evaluate sweep point (BASELINE.json configs[4])"),
    # --log2 is the circuit depth L here: reference-shaped layered circuit, layer i has 2^i gates over 2^(i+1) wires,
    # 2^L inputs; the layer-i sumcheck runs over 4^(i+1) entries x 4 tables
    "gkr": (0, "BN254_FQ", 2, 2, 12, "GKR prove of a reference-shaped layered add/mul circuit (gkr_protocol::prove)"),
    # --log2 = number of variables of the committed polynomial (the input commitment of succinct GKR for the
    # 2^22-wide input layer of configs[3])
    # --log2 = log2 of the layer width: prove_succinct of the configs[3] circuit shape over the curve's scalar field
    "succinct": (2, "BLS12_381_FR", 2, 2, 22, "succinct GKR (gkr/src/succinct_gkr_protocol.rs): GKR prove of 16 layers x 2^22 gates over BLS12-381 Fr + multilinear KZG commitment of the 2^22 inputs + two openings"),
    "kzg": (2, "BLS12_381_FR", 1, 1, 22, "multilinear KZG over BLS12-381 G1: commit_to_polynomial + open_and_prove of a 2^22-entry polynomial (succinct GKR's input commitment)"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------- CPU baseline
def cpu_prove_once(field: int, P: int, D: int, log2: int, threads: int = 1):
    """one prove of the oracle (reference pass structure) on a 2^log2 sample; seconds.  threads == 1 is the reference's
    behaviour (single-threaded Rust, no rayon); threads > 1 runs the oracle's data-parallel loops on OpenMP threads --
    a stronger-than-the-reference baseline, reported separately and labelled."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import coracle as co
    co.set_threads(threads)
    try:
        return _cpu_prove_once(co, field, P, D, log2)
    finally:
        co.set_threads(1)


def _cpu_prove_once(co, field: int, P: int, D: int, log2: int):
    """the sample is the head of the GPU arm's own tables: entries 0 .. 2^log2 - 1 of table i under SEED, from the oracle's
    restatement of the device generator (zko_table_generate == zk_table_generate, tests/test_oracle.py)"""
    n = 1 << log2
    threads = co.get_threads()
    co.set_threads(os.cpu_count() or 1)          # generating the inputs is not part of the timed prove
    try:
        if D == 1:
            tab = co.table_generate(field, SEED, 0, n)
        else:
            # the reference panics on a single product (sum_polynomial.rs:58-61): f*g is posed as f*g + 0*0
            Pref = max(P, 2)
            tabs = np.zeros((Pref, D, n, 4), dtype=np.uint64)
            for i in range(P * D):
                tabs[i // D, i % D] = co.table_generate(field, SEED, i, n)
            claimed = np.zeros(4, dtype=np.uint64)
            co.lib().zko_fe_sum(field, co._p(co.sumpoly_reduce(field, tabs)), n, co._p(claimed))    # the true claimed sum
    finally:
        co.set_threads(threads)
    t0 = time.perf_counter()
    if D == 1:
        co.basic_prove(field, tab)
    else:
        co.product_prove(field, tabs, claimed, co.Transcript())
    return time.perf_counter() - t0


def synthetic_circuit(depth: int, seed: int = SEED):
    """reference-shaped circuit: layer i has one gate per output index 0..2^i-1 reading two seeded wires of the
    2^(i+1)-wide layer below, operator from the seed; duplicate-free by construction (distinct outputs)."""
    rng = np.random.default_rng(seed)
    layers = []
    for i in range(depth):
        n = 1 << i
        left = rng.integers(0, 2 << i, size=n)
        right = rng.integers(0, 2 << i, size=n)
        op = rng.integers(0, 2, size=n)
        layers.append([(int(left[o]), int(right[o]), o, int(op[o])) for o in range(n)])
    return layers


def gkr_inputs(field: int, depth: int):
    rng = np.random.default_rng(SEED + 1)
    tab = rng.integers(0, 1 << 62, size=(1 << depth, 4), dtype=np.uint64)
    tab[:, 3] &= np.uint64((1 << 58) - 1)
    return tab


def cpu_gkr_once(field: int, depth: int):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import coracle as co
    c = co.Circuit(synthetic_circuit(depth))
    I = gkr_inputs(field, depth)
    t0 = time.perf_counter()
    co.gkr_prove(field, c, I)
    return time.perf_counter() - t0


def cpu_succinct_once(depth: int):
    """prove_succinct the way the reference runs it (succinct_gkr_protocol.rs:35-169), oracle port, one thread: commitment to the
    2^depth inputs + gkr_protocol::prove of a reference-shaped depth-`depth` circuit + two openings; seconds"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import coracle as co
    field = 2
    c = co.Circuit(synthetic_circuit(depth))
    I = gkr_inputs(field, depth)
    taus = co.table_generate(field, SEED, 98, 64)[:depth].copy()
    g1 = co.kzg_setup_g1(taus)                       # the trusted setup is an input of prove_succinct, not part of it
    t0 = time.perf_counter()
    co.kzg_commit(I, g1)
    proof = co.gkr_prove(field, c, I)
    chal = proof.challenges[-2 * depth:]
    co.kzg_open(I, g1, chal[:depth])
    co.kzg_open(I, g1, chal[depth:])
    return time.perf_counter() - t0


def wide_circuit_arrays(width_log2: int, depth: int = 16, seed: int = SEED):
    """depth layers of 2^w gates each.  Layers 1..depth-1: gate g drives output g from two seeded wires of the layer below;
    layer 0 reduces the 2^w wires below it into TWO outputs (gate g reads wire g and a seeded wire, output g mod 2), so the
    circuit keeps the reference's output shape (one output bit -> one challenge r_a, 64 bytes absorbed).  Duplicate-free.
    Returned in the C-ABI's own flat layout (uint64 layer offsets, uint32 indices, uint8 operators): no conversion copies."""
    rng = np.random.default_rng(seed)
    n = 1 << width_log2
    left = np.empty(depth * n, dtype=np.uint32)
    right = np.empty(depth * n, dtype=np.uint32)
    out = np.empty(depth * n, dtype=np.uint32)
    op = np.empty(depth * n, dtype=np.uint8)
    g = np.arange(n, dtype=np.uint32)
    for li in range(depth):
        sl = slice(li * n, (li + 1) * n)
        right[sl] = rng.integers(0, n, size=n, dtype=np.uint32)
        op[sl] = rng.integers(0, 2, size=n, dtype=np.uint8)
        if li == 0:
            left[sl] = g
            out[sl] = g & 1
        else:
            left[sl] = rng.integers(0, n, size=n, dtype=np.uint32)
            out[sl] = g
    off = np.arange(depth + 1, dtype=np.uint64) * np.uint64(n)
    bits = [1] + [width_log2] * depth
    return bits, (off, left, right, out, op)


def keccak_digest(arrays) -> str:
    """Keccak-256 (the transcript's own hash) of the concatenated little-endian limb bytes of the proof arrays"""
    from zk_cryptography_research_implementations_b200.transcripts import Transcript
    tr = Transcript()
    for a in arrays:
        tr.append(np.ascontiguousarray(a, dtype=np.uint64).tobytes())
    return tr.sample_random_challenge().hex()


def run_gkr_wide(args, wl, steps=None, warmup=None):
    field, fname, P, D, w_default, desc = wl
    w = args.log2 if (args.log2 and args.workload == "gkr_wide") else w_default
    steps = steps or args.steps
    warmup = args.warmup if warmup is None else warmup
    depth = 16
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank != 0:
            return None
        args.workload = "gkr"
        return run_gkr(args, WORKLOADS["gkr"])
    import torch
    import zk_cryptography_research_implementations_b200 as zk
    from zk_cryptography_research_implementations_b200 import gkr
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    ctx = zk.Context(field, local_rank, stream=torch.cuda.current_stream().cuda_stream)
    if world > 1:   # N > 1: the circuit and the input layer are replicated, every layer's phase tables and sumchecks are sharded
        import torch.distributed as dist
        from zk_cryptography_research_implementations_b200 import sharded
        sharded.init_comm(ctx)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def prove():
        return gkr.prove_wide(ctx, circuit, dev_I, sharded=world > 1, collapse_len=args.collapse_len)
    t0 = time.perf_counter()
    bits, flat = wide_circuit_arrays(w, depth)
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    circuit = gkr.WideCircuit(ctx, bits, flat=flat)      # upload + range / duplicate checks + three CSR orderings per layer, on the GPU
    ctx.synchronize()
    setup_s = time.perf_counter() - t0
    dev_I = ctx.generate(SEED + 1, 0, 1 << w)            # value: the input layer is resident in HBM when the clock starts
    I = dev_I.download()
    proof = None
    for _ in range(warmup):
        barrier()
        proof = prove()
    ctx.set_profiling(True)
    ctx.reset_stats()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    times = []
    for _ in range(steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        proof = prove()                              # resident inputs -> proof on the host (every rank gets the same one)
        e1.record()
        barrier()
        times.append(e0.elapsed_time(e1))
    st = ctx.stats()
    ctx.set_profiling(False)
    ms = statistics.mean(times)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # e2e: the reference-facing call, inputs in pinned HOST memory, proof back on the host
    e2e_ms = None
    if not args.no_e2e and world == 1:
        pin = C.c_void_p()
        if ctx.lib.zk_pinned_alloc(C.c_size_t(I.nbytes), C.byref(pin)) == 0:
            host_I = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_uint64)), shape=(I.size,)).reshape(I.shape)
            host_I[:] = I
            gkr.prove_wide(ctx, circuit, host_I)
            ts = []
            for _ in range(max(1, min(steps, args.e2e_steps))):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                proof_e = gkr.prove_wide(ctx, circuit, host_I)
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
            e2e_ms = statistics.mean(ts)
            assert np.array_equal(proof_e.claimed_sum, proof.claimed_sum)
            del host_I
            ctx.lib.zk_pinned_free(pin)
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        barrier()
        dev_I.free()
        circuit.close()
        ctx.close()
        return None
    # after the timed region: the reference's verifier (gkr_protocol.rs:146-236) over this very proof, wiring predicates
    # evaluated from the gate list on the GPU, the input layer's W(u), W(v) by the evaluate kernels
    t0 = time.perf_counter()
    verified = bool(gkr.verify_wide(ctx, circuit, proof, dev_I))
    verify_ms = (time.perf_counter() - t0) * 1e3
    hbm_peak, peak_src = peaks()
    achieved = st["round_bytes"] / (st["round_ms"] * 1e-3) / 1e9 if st["round_ms"] > 0 else 0.0
    cpu = None
    if not args.no_cpu:
        d = args.cpu_depth
        t = cpu_gkr_once(field, d)
        cpu = {"value": t * 1e3, "unit": "ms", "cores": 1, "kind": "port",
               "sample": "reference-shaped depth-%d circuit (2^%d inputs, %d gates): the reference's dense 2^(3i+2) wiring tables cannot "
                         "express or hold a 2^%d-wide layer; oracle C restatement, 1 thread" % (d, d, (1 << d) - 1, w)}
    rounds = circuit.total_rounds()
    coeffs = np.stack([p.coefficients for sp in proof.sumcheck_proofs for p in sp.round_univariate_polynomials])
    # algorithmic HBM bytes of one prove: per layer 2 phases x 128 x 3 tables x 2^w, plus the gate passes
    alg_bytes = depth * (2 * 128.0 * 3 * (1 << w) + 3 * 48.0 * (1 << w))
    line = {"metric": "gkr_prove_ms", "value": ms, "unit": "ms", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u256 (8x u32 Montgomery limbs, integer IMAD arithmetic)", "data": "synthetic",
            "config": {"workload": "gkr_wide: " + desc, "field": fname, "depth": depth, "width_log2": w, "gates": depth << w,
                       "layer_bits": bits, "sumcheck_rounds": rounds, "prover": "sparse two-phase (csrc/gkr_wide.cu)",
                       "sharding": ("none" if world == 1 else "circuit, layer values and transcript replicated; every layer's phase tables and "
                                    "sumchecks sharded on the low index bits across %d ranks (zk_gkr_prove_wide_sharded)" % world),
                       "circuit_setup_s": setup_s,
                       "circuit_setup_note": "NOT inside `value`: zk_wide_circuit_create = upload of the gate lists + range / duplicate checks + the three "
                                             "CSR orderings of every layer, built on the GPU (the reference builds its wiring tables inside prove, "
                                             "arithmetic_circuit.rs:126-163); generating the synthetic gate lists with numpy took %.2f s more" % gen_s,
                       "l2": "per phase 3 tables x %d MiB" % ((32 << w) >> 20)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                         "peak_source": peak_src, "kernel": "sumcheck round kernels <%s,P=1,D=2,+1 linear table> (%d launches; most are on tables far smaller than L2: "
                                   "latency-bound, the fraction is not a bandwidth statement)" % (fname, st["round_launches"]),
                         "kernel_ms_per_step": st["round_ms"] / max(steps, 1),
                         "whole_prove": {"algorithmic_bytes": alg_bytes, "hbm_floor_ms": alg_bytes / (hbm_peak * 1e9) * 1e3,
                                         "frac": alg_bytes / (hbm_peak * 1e9) * 1e3 / ms}},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(I.nbytes), "d2h_bytes_per_step": int(rounds * 4 * 32),
                    "call": "zk_gkr_prove_wide: input layer in pinned host memory -> proof on the host (the circuit's CSR lives on the GPU); "
                            "`value` is zk_gkr_prove_wide_device with the input layer already in HBM"},
            "gpu_launches": st["launches"], "clocks": clocks, "tail_log": ctx.tail_log(),
            "exchange": None if world == 1 else ctx_exchange_name(ctx),
            "verified": verified, "verify_ms": verify_ms,
            "verified_by": "zk_gkr_verify_wide_device (gkr_protocol.rs:146-236) on the proof of the last timed step",
            "proof_digest": keccak_digest([coeffs, proof.claimed_sum, proof.wb_evaluations, proof.wc_evaluations])}
    barrier()
    dev_I.free()
    circuit.close()
    ctx.close()
    return line


def run_gkr(args, wl):
    """GKR prove ms (BASELINE metric, second half).  Single GPU: the dense (b,c) formulation of the reference."""
    field, fname, P, D, depth_default, desc = wl
    depth = args.log2 or depth_default
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    if args.impl == "reference":
        d = min(depth, args.cpu_depth)
        times = [cpu_gkr_once(field, d) for _ in range(max(args.steps, 1))]
        t = statistics.mean(times)
        line = {"impl": "reference", "metric": "gkr_prove_ms", "value": t * 1e3, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
                "dtype": "u256 (4x u64 Montgomery limbs)", "data": "synthetic",
                "config": {"workload": "gkr: " + desc, "field": fname, "depth": d, "requested_depth": depth},
                "cpu_baseline": {"value": t * 1e3, "unit": "ms", "cores": 1, "kind": "port",
                                 "sample": "depth-%d circuit (dense 2^(3i+2) wiring tables like the reference), oracle C restatement, 1 thread" % d},
                "e2e": {"value": t * 1e3, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        return line
    import torch
    import zk_cryptography_research_implementations_b200 as zk
    from zk_cryptography_research_implementations_b200 import gkr
    from zk_cryptography_research_implementations_b200.circuit import Circuit, Gate, Layer
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    ctx = zk.Context(field, local_rank, stream=torch.cuda.current_stream().cuda_stream)
    layers = synthetic_circuit(depth)
    circuit = Circuit.new(field, [Layer.new([Gate.new(*g) for g in l]) for l in layers])
    I = gkr_inputs(field, depth)
    for _ in range(args.warmup):
        gkr.prove(ctx, circuit, I)
    ctx.set_profiling(True)
    ctx.reset_stats()
    sampler = ClockSampler(local_rank)
    times = []
    for _ in range(args.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        proof = gkr.prove(ctx, circuit, I)     # host inputs -> proof on the host: this IS the end-to-end call
        torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop()
    st = ctx.stats()
    ms = statistics.mean(times)
    hbm_peak, peak_src = peaks()
    achieved = st["round_bytes"] / (st["round_ms"] * 1e-3) / 1e9 if st["round_ms"] > 0 else 0.0
    cpu = None
    if not args.no_cpu:
        d = min(depth, args.cpu_depth)
        t = cpu_gkr_once(field, d)
        cpu = {"value": t * 1e3, "unit": "ms", "cores": 1, "kind": "port",
               "sample": "depth-%d circuit (the GPU ran depth %d; the reference's dense 2^(3i+2) wiring tables make deeper ones "
                         "infeasible on the CPU), oracle C restatement, 1 thread" % (d, depth)}
    rounds = depth * (depth + 1)
    line = {"metric": "gkr_prove_ms", "value": ms, "unit": "ms", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u256 (8x u32 Montgomery limbs, integer IMAD arithmetic)", "data": "synthetic",
            "config": {"workload": "gkr: " + desc, "field": fname, "depth": depth, "inputs": 1 << depth, "sumcheck_rounds": rounds,
                       "largest_layer_entries": 4 ** depth, "tables_per_layer": 4, "l2": "largest layer 4 x %d MiB" % ((4 ** depth * 32) >> 20)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                         "peak_source": peak_src, "kernel": "fold_evals_kernel / round_evals_kernel <%s,P=2,D=2> (%d launches)" % (fname, st["round_launches"]),
                         "kernel_ms_per_step": st["round_ms"] / max(args.steps, 1)},
            "cpu_baseline": cpu,
            "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": int(I.nbytes), "d2h_bytes_per_step": int(rounds * 4 * 32),
                    "note": "value already is the host-to-host zk_gkr_prove call (circuit + inputs on the host, proof on the host)"},
            "gpu_launches": st["launches"], "clocks": clocks,
            "verified": bool(gkr.verify(ctx, circuit, proof, I)),
            "verified_by": "zk_gkr_verify (gkr_protocol.rs:146-236) on the proof of the last timed step",
            "proof_digest": keccak_digest([np.stack([p.coefficients for sp in proof.sumcheck_proofs for p in sp.round_univariate_polynomials]), proof.claimed_sum])}
    ctx.close()
    return line


def run_mle(args, wl, log2=None, sweep=None, steps=None, warmup=None):
    """configs[4]: MLE evaluate (all n challenges, one read of the table per 3 variables) and one partial_evaluate,
    against the HBM roofline.  N > 1: table sharded on the low index bits, zk_mle_evaluate_sharded."""
    field, fname, P, D, log2_default, desc = wl
    log2 = log2 or (args.log2 if args.workload == "mle" else 0) or log2_default
    sweep_arg = args.sweep if sweep is None else sweep
    n_steps = steps or args.steps
    n_warm = args.warmup if warmup is None else warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return None
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import coracle as co
        sl = min(log2, args.cpu_log2)
        tab = co.table_generate(field, SEED, 0, 1 << sl)
        rs = co.table_generate(field, SEED, 99, 64)[:sl].copy()
        times = []
        for _ in range(max(args.steps, 1)):
            t0 = time.perf_counter()
            co.mle_evaluate(field, tab, rs)
            times.append(time.perf_counter() - t0)
        t = statistics.mean(times)
        v = (1 << sl) / t / MEGA
        return {"impl": "reference", "metric": "mle_evaluate_Melems_per_s", "value": v, "unit": "Melems/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u256 (4x u64 Montgomery limbs)", "data": "synthetic",
                "config": {"workload": "mle: " + desc, "field": fname, "log2_entries": log2, "cpu_sample_log2": sl},
                "cpu_baseline": {"value": v, "unit": "Melems/s", "cores": 1, "kind": "port",
                                 "sample": "evaluate of the first 2^%d entries of the seeded table, oracle C restatement (n folds with fresh vectors), 1 thread" % sl},
                "e2e": {"value": v, "unit": "Melems/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    import torch
    import torch.distributed as dist
    import zk_cryptography_research_implementations_b200 as zk
    from zk_cryptography_research_implementations_b200 import sharded
    from zk_cryptography_research_implementations_b200.core import _ptr
    torch.cuda.set_device(local_rank)
    ctx = zk.Context(field, local_rank, stream=torch.cuda.current_stream().cuda_stream)
    if world > 1:
        sharded.init_comm(ctx)
    lib = ctx.lib
    N = 1 << log2
    m = N // world
    table = ctx.generate(SEED, 0, m, first=rank, step=world)
    rs = np.ascontiguousarray(ctx.generate(SEED, 99, 64).download()[:log2])
    out = np.zeros(4, dtype=np.uint64)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def ev():
        ctx.check(lib.zk_mle_evaluate_sharded(ctx.h, table.h, _ptr(rs), log2, _ptr(out)))

    def timed(fn, prep):
        for _ in range(n_warm):
            prep(); barrier(); fn(); barrier()
        ctx.reset_stats()
        ts = []
        for _ in range(n_steps):
            prep(); barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); barrier()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([statistics.mean(ts)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_eval = timed(ev, lambda: None)
    result = out.copy()
    launches = ctx.stats()["launches"]
    # after the timed region: the same value through the other implementation of evaluate (n folds, three variables per
    # pass: fold_multi_kernel, the reference's own pass structure) -- two independent kernels must agree limb for limb
    os.environ["ZKB200_EVAL_FOLDS"] = "1"
    try:
        ev()
    finally:
        del os.environ["ZKB200_EVAL_FOLDS"]
    folds_agree = bool(np.array_equal(out, result))
    # e2e (one GPU): the table in pinned host memory -> HBM -> the value back on the host
    e2e = None
    if world == 1 and not args.no_e2e:
        pin = C.c_void_p()
        if lib.zk_pinned_alloc(C.c_size_t(m * 32), C.byref(pin)) == 0:
            ctx.check(lib.zk_table_download(ctx.h, table.h, C.cast(pin, C.POINTER(C.c_uint64))))

            def up_and_eval():
                ctx.check(lib.zk_table_upload_into(ctx.h, table.h, pin, m))
                ev()
            ms_e2e = timed(up_and_eval, lambda: None)
            e2e = {"value": N / (ms_e2e * 1e-3) / MEGA, "unit": "Melems/s", "h2d_bytes_per_step": m * 32, "d2h_bytes_per_step": 32, "ms_per_step": ms_e2e,
                   "pcie_h2d_GBps_per_gpu": m * 32 / max(ms_e2e - ms_eval, 1e-6) / 1e6,
                   "call": "zk_table_upload_into (pinned host -> HBM) + zk_mle_evaluate"}
            assert np.array_equal(out, result)
            lib.zk_pinned_free(pin)
    # one partial_evaluate of variable 0 (in place: read N, write N/2), table refilled between steps
    r0 = np.ascontiguousarray(rs[0])
    ms_fold = timed(lambda: ctx.check(lib.zk_mle_partial_evaluate(ctx.h, table.h, 0, _ptr(r0))),
                    lambda: table.regenerate(SEED, 0, m, rank, world))
    clocks = sampler.stop() if sampler else None
    hbm_peak, peak_src = peaks()
    # --sweep a,b,c: the same two measurements at other table sizes (BASELINE.json configs[4] is a sweep), same context
    sweep = []
    for lg in [int(x) for x in sweep_arg.split(",") if x] if sweep_arg else []:
        if lg == log2 or (1 << lg) < world:
            continue
        mm = (1 << lg) // world
        t2 = ctx.generate(SEED, 0, mm, first=rank, step=world)
        rs2 = np.ascontiguousarray(ctx.generate(SEED, 99, 64).download()[:lg])
        ms_e = timed(lambda: ctx.check(lib.zk_mle_evaluate_sharded(ctx.h, t2.h, _ptr(rs2), lg, _ptr(out))), lambda: None)
        r02 = np.ascontiguousarray(rs2[0])
        ms_f = timed(lambda: ctx.check(lib.zk_mle_partial_evaluate(ctx.h, t2.h, 0, _ptr(r02))), lambda: t2.regenerate(SEED, 0, mm, rank, world))
        sweep.append({"log2_entries": lg, "evaluate_ms": ms_e, "evaluate_Melems_per_s": (1 << lg) / (ms_e * 1e-3) / MEGA,
                      "evaluate_frac_hbm": 32.0 * mm / (ms_e * 1e-3) / 1e9 / hbm_peak,
                      "partial_evaluate_ms": ms_f, "partial_evaluate_Melems_per_s": (1 << lg) / (ms_f * 1e-3) / MEGA,
                      "partial_evaluate_frac_hbm": 48.0 * mm / (ms_f * 1e-3) / 1e9 / hbm_peak,
                      "l2": "table larger than L2" if mm * 32 > 126e6 else "table fits L2 (timed back to back: L2-resident)"})
        t2.free()
    line = None
    if rank == 0:
        ach_eval = 32.0 * m / (ms_eval * 1e-3) / 1e9          # algorithmic: ONE read of the table
        ach_fold = 48.0 * m / (ms_fold * 1e-3) / 1e9
        cpu = None
        verified = None
        if not args.no_cpu:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import coracle as co
            sl = min(log2, args.cpu_log2)
            co.set_threads(os.cpu_count() or 1)
            tab = co.table_generate(field, SEED, 0, 1 << sl)
            co.set_threads(1)
            t0 = time.perf_counter()
            want = co.mle_evaluate(field, tab, rs[:sl].copy())
            t = time.perf_counter() - t0
            cpu = {"value": (1 << sl) / t / MEGA, "unit": "Melems/s", "cores": 1, "kind": "port",
                   "sample": "evaluate of the first 2^%d entries of the seeded table (%.2f s), oracle C restatement, 1 thread" % (sl, t)}
            if sl == log2:      # the oracle evaluated the very table the GPU did: compare the values
                verified = bool(np.array_equal(want, result))
        verified = folds_agree if verified is None else (verified and folds_agree)
        line = {
            "metric": "mle_evaluate_Melems_per_s", "value": N / (ms_eval * 1e-3) / MEGA, "unit": "Melems/s", "n_gpus": world,
            "steps": n_steps, "warmup": n_warm, "ms_per_step": ms_eval, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u256 (8x u32 Montgomery limbs, integer IMAD arithmetic)", "data": "synthetic",
            "config": {"workload": "mle: " + desc, "field": fname, "log2_entries": log2, "table_bytes_total": N * 32,
                       "sharding": "low index bits across %d rank(s)" % world, "l2": "table larger than L2" if m * 32 > 126e6 else "table fits L2"},
            "roofline": {"bound": "hbm", "achieved": ach_eval, "peak": hbm_peak, "unit": "GB/s", "frac": ach_eval / hbm_peak, "traffic": None,
                         "peak_source": peak_src, "kernel": "zk_mle_evaluate passes <%s> (algorithmic bytes = 32 N: one read)" % fname,
                         "note": "timed around the whole evaluate call (all passes + host fold tables)"},
            "partial_evaluate": {"ms": ms_fold, "Melems_per_s": N / (ms_fold * 1e-3) / MEGA, "achieved_GBps": ach_fold, "frac": ach_fold / hbm_peak,
                                 "kernel": "fold0_kernel (read N, write N/2: 48 N bytes)"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "sweep": sweep,
            "verified": verified,
            "verified_by": "after the timed region: the inner-product result equals the result of the fold passes (a second, independent kernel path); "
                           "and the oracle's mle_evaluate when the CPU sample is the whole table",
            "result_digest": keccak_digest([result])}
    barrier()
    ctx.close()
    return line


def run_kzg(args, wl, log2=None, steps=None, warmup=None):
    """MultilinearKZG::commit_to_polynomial (one 2^n-point G1 multi-scalar multiplication) and open_and_prove
    (n more over the folded setup).  value = points committed per second; the bucket kernel is integer-multiplier bound, so
    the roofline is the register-resident mixed-addition rate measured by zk_g1_arith_probe in the same run."""
    field, fname, P, D, log2_default, desc = wl
    log2 = log2 or (args.log2 if args.workload == "kzg" else 0) or log2_default
    n_steps = steps or args.steps
    n_warm = args.warmup if warmup is None else warmup
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    N = 1 << log2
    sys.path.insert(0, os.path.join(ROOT, "oracle"))

    def cpu_commit(sample_log2, threads):
        import coracle as co
        taus = co.table_generate(field, SEED, 98, 64)[:sample_log2].copy()
        vals = co.table_generate(field, SEED, 0, 1 << sample_log2)
        co.set_threads(os.cpu_count() or 1)
        try:
            g1 = co.kzg_setup_g1(taus)           # building the setup is not the reference's timed path
        finally:
            co.set_threads(threads)
        try:
            t0 = time.perf_counter()
            c = co.kzg_commit(vals, g1)
            return time.perf_counter() - t0, c, taus
        finally:
            co.set_threads(1)

    if args.impl == "reference":
        sl = min(log2, 11)
        th = max(1, args.cpu_threads)
        times = [cpu_commit(sl, th)[0] for _ in range(max(args.steps, 1))]
        t = statistics.mean(times)
        v = (1 << sl) / t / MEGA
        return {"impl": "reference", "metric": "kzg_commit_Mpoints_per_s", "value": v, "unit": "Mpoints/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u384 / u256 (6x u64 Montgomery base field, 4x u64 scalars)", "data": "synthetic",
                "config": {"workload": "kzg: " + desc, "field": fname, "log2_entries": log2, "cpu_sample_log2": sl},
                "cpu_baseline": {"value": v, "unit": "Mpoints/s", "cores": th, "kind": "port",
                                 "sample": "commit_to_polynomial of the first 2^%d entries: one double-and-add mul_bigint per point as the reference "
                                           "(multilinear_kzg.rs:39-43), oracle C restatement, %d thread(s)" % (sl, th)},
                "e2e": {"value": v, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    import torch
    import zk_cryptography_research_implementations_b200 as zk
    from zk_cryptography_research_implementations_b200.multilinear_kzg import MultilinearKZG, MultilinearKZGProof, TrustedSetup
    torch.cuda.set_device(0)
    ctx = zk.Context(field, 0, stream=torch.cuda.current_stream().cuda_stream)
    lib = ctx.lib
    taus = np.ascontiguousarray(ctx.generate(SEED, 98, 64).download()[:log2])
    opening = np.ascontiguousarray(ctx.generate(SEED, 99, 64).download()[:log2])
    t0 = time.perf_counter()
    setup = TrustedSetup.initialize_setup(ctx, taus)
    ctx.synchronize()
    setup_s = time.perf_counter() - t0
    table = ctx.generate(SEED, 0, N)
    ctx.synchronize()

    def timed(fn, k, w):
        for _ in range(w):
            fn()
        ts = []
        for _ in range(k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        return statistics.mean(ts)

    sampler = ClockSampler(0)
    ctx.reset_stats()
    box = {}
    ms_commit = timed(lambda: box.__setitem__("c", MultilinearKZG.commit_to_polynomial(table, setup)), n_steps, n_warm)
    launches = ctx.stats()["launches"] // max(n_steps + n_warm, 1)
    ms_open = timed(lambda: box.__setitem__("p", MultilinearKZG.open_and_prove(table, setup, opening)), max(1, n_steps // 2), 1)
    clocks = sampler.stop()
    commitment, proof = box["c"], box["p"]
    # e2e: evaluations in pinned host memory -> commitment on the host
    e2e = None
    if not args.no_e2e:
        pin = C.c_void_p()
        if lib.zk_pinned_alloc(C.c_size_t(N * 32), C.byref(pin)) == 0:
            host_vals = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_uint64)), shape=(N * 4,)).reshape(N, 4)
            ctx.check(lib.zk_table_download(ctx.h, table.h, C.cast(pin, C.POINTER(C.c_uint64))))
            ms_e2e = timed(lambda: box.__setitem__("ce", MultilinearKZG.commit_to_polynomial(host_vals, setup)), max(1, min(n_steps, 3)), 1)
            assert np.array_equal(box["ce"], commitment)
            e2e = {"value": N / (ms_e2e * 1e-3) / MEGA, "unit": "Mpoints/s", "h2d_bytes_per_step": N * 32, "d2h_bytes_per_step": 96,
                   "ms_per_step": ms_e2e, "call": "zk_kzg_commit (evaluations in pinned host memory -> HBM -> commitment on the host)"}
            del host_vals
            lib.zk_pinned_free(pin)
    # the multiplier ceiling: register-resident mixed additions and 381-bit products on the full grid
    peak_madd = peak_mul = None
    if not args.no_probe:
        ops, ms = C.c_double(), C.c_double()
        ctx.check(lib.zk_g1_arith_probe(ctx.h, 1, 2000, 4, C.byref(ops), C.byref(ms)))
        peak_madd = ops.value
        ctx.check(lib.zk_g1_arith_probe(ctx.h, 0, 4000, 4, C.byref(ops), C.byref(ms)))
        peak_mul = ops.value
    windows = (256 + 15) // 16 if N >= (1 << 20) else None
    madds = N * windows if windows else None
    ach = madds / (ms_commit * 1e-3) if madds else None
    # after the timed region: the pairing check of the reference's verify() on the commitment and the opening, and the
    # evaluation against zk's own evaluate
    t0 = time.perf_counter()
    verified = bool(MultilinearKZG.verify(setup, commitment, opening, proof))
    verify_s = time.perf_counter() - t0
    from zk_cryptography_research_implementations_b200.polynomials import MultilinearPolynomial
    verified = verified and bool(np.array_equal(MultilinearPolynomial(ctx, table).evaluate(opening), proof.evaluation))
    cpu = None
    if not args.no_cpu:
        sl = min(log2, 11)
        t, c_cpu, taus_cpu = cpu_commit(sl, 1)
        cpu = {"value": (1 << sl) / t / MEGA, "unit": "Mpoints/s", "cores": 1, "kind": "port",
               "sample": "commit_to_polynomial of a 2^%d-entry polynomial (%.2f s): one double-and-add mul_bigint per point as the reference, "
                         "oracle C restatement, 1 thread" % (sl, t)}
        # the same sample through the GPU path must give the oracle's commitment
        s2 = TrustedSetup.initialize_setup(ctx, taus_cpu)
        t2 = ctx.generate(SEED, 0, 1 << sl)
        verified = verified and bool(np.array_equal(MultilinearKZG.commit_to_polynomial(t2, s2), c_cpu))
        s2.release()
    line = {
        "metric": "kzg_commit_Mpoints_per_s", "value": N / (ms_commit * 1e-3) / MEGA, "unit": "Mpoints/s", "n_gpus": 1, "steps": n_steps,
        "warmup": n_warm, "ms_per_step": ms_commit, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u384 / u256 (12x u32 Montgomery base field, 8x u32 scalars; integer IMAD arithmetic)", "data": "synthetic",
        "config": {"workload": "kzg: " + desc, "field": fname, "curve": "BLS12-381 G1", "log2_entries": log2,
                   "timer": "host wall clock around the synchronous call (it ends with a host-side combine of the window sums)"},
        "kzg_commit_ms": ms_commit, "kzg_open_ms": ms_open, "trusted_setup_s": setup_s, "kzg_verify_s": verify_s,
        "roofline": {"bound": "imad", "achieved": ach, "peak": peak_madd, "unit": "mixed additions/s", "frac": (ach / peak_madd) if ach and peak_madd else None,
                     "traffic": None, "kernel": "msm_bucket_kernel (N x %s windows mixed additions per commit over the whole call's time)" % windows,
                     "peak_source": "zk_g1_arith_probe kind 1 in this run (register-resident XYZZ += affine, 4 blocks/SM); "
                                    "381-bit Montgomery products: %s /s" % ("%.3g" % peak_mul if peak_mul else "n/a")},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "verified": verified,
        "verified_by": "after the timed region: MultilinearKZG::verify (pairing check, host) accepts the commitment with the opening proof; "
                       "the opened value equals zk_mle_evaluate; the oracle's commitment of the CPU sample equals the GPU's",
        "result_digest": keccak_digest([np.ascontiguousarray(commitment.reshape(-1, 4)), np.ascontiguousarray(proof.proofs.reshape(-1, 4))])}
    setup.release()
    ctx.close()
    return line


def run_succinct(args, wl, log2=None, steps=None, warmup=None):
    """SURVEY 8 f4, whole: prove_succinct (succinct_gkr_protocol.rs:35-169) = commit to the input polynomial + the GKR layer
    sumchecks + two openings of the input polynomial at (rb, rc).  value = ms per proof, input layer resident in HBM.
    N > 1: circuit, setup and input layer replicated; the layer sumchecks and every large multi-scalar multiplication are
    spread over the ranks (zk_gkr_prove_wide_sharded, zk_kzg_commit_sharded / zk_kzg_open_sharded); same proof on every rank."""
    field, fname, P, D, w_default, desc = wl
    w = log2 or (args.log2 if args.workload == "succinct" else 0) or w_default
    n_steps = steps or args.steps
    n_warm = args.warmup if warmup is None else warmup
    depth = 16
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return None
        # the reference's dense wiring tables cannot hold a 2^22-wide layer; its own largest succinct test is 3 layers over 8
        # inputs.  Timed sample: the reference-shaped GKR part (run_gkr) -- the commitment part is `--workload kzg`.
        args.workload = "gkr"
        return run_gkr(args, WORKLOADS["gkr"])
    import torch
    import zk_cryptography_research_implementations_b200 as zk
    from zk_cryptography_research_implementations_b200 import gkr
    from zk_cryptography_research_implementations_b200.multilinear_kzg import TrustedSetup
    torch.cuda.set_device(local_rank)
    ctx = zk.Context(field, local_rank, stream=torch.cuda.current_stream().cuda_stream)
    if world > 1:
        import torch.distributed as dist
        from zk_cryptography_research_implementations_b200 import sharded
        sharded.init_comm(ctx)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    bits, flat = wide_circuit_arrays(w, depth)
    t0 = time.perf_counter()
    circuit = gkr.WideCircuit(ctx, bits, flat=flat)
    ctx.synchronize()
    circuit_s = time.perf_counter() - t0
    taus = np.ascontiguousarray(ctx.generate(SEED, 98, 64).download()[:w])
    t0 = time.perf_counter()
    setup = TrustedSetup.initialize_setup(ctx, taus)
    ctx.synchronize()
    setup_s = time.perf_counter() - t0
    dev_I = ctx.generate(SEED + 1, 0, 1 << w)

    def prove(inputs=None):
        return gkr.prove_succinct(ctx, circuit, dev_I if inputs is None else inputs, setup, sharded=world > 1, collapse_len=args.collapse_len)
    proof = None
    for _ in range(n_warm):
        barrier()
        proof = prove()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ctx.reset_stats()
    ts = []
    for _ in range(n_steps):
        barrier()
        t0 = time.perf_counter()
        proof = prove()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    ms = statistics.mean(ts)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = ctx.stats()["launches"] // max(n_steps, 1)
    clocks = sampler.stop() if sampler else None
    e2e_ms = None
    if not args.no_e2e and world == 1:
        pin = C.c_void_p()
        n_in = 1 << w
        if ctx.lib.zk_pinned_alloc(C.c_size_t(n_in * 32), C.byref(pin)) == 0:
            host_I = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_uint64)), shape=(n_in * 4,)).reshape(n_in, 4)
            ctx.check(ctx.lib.zk_table_download(ctx.h, dev_I.h, C.cast(pin, C.POINTER(C.c_uint64))))
            prove(host_I)
            t0 = time.perf_counter()
            pe = prove(host_I)
            torch.cuda.synchronize()
            e2e_ms = (time.perf_counter() - t0) * 1e3
            assert np.array_equal(pe.input_polynomial_commitment, proof.input_polynomial_commitment)
            del host_I
            ctx.lib.zk_pinned_free(pin)
    barrier()
    line = None
    if rank == 0:
        t0 = time.perf_counter()
        verified = bool(gkr.verify_succinct(ctx, circuit, proof, setup)) and bool(gkr.verify_succinct(ctx, circuit, proof, setup, bind_input_openings=True))
        verify_s = time.perf_counter() - t0
        rounds = circuit.total_rounds()
        cpu = None
        if not args.no_cpu and world == 1:
            try:
                d = min(args.cpu_depth, 7)
                t = cpu_succinct_once(d)
                cpu = {"value": t * 1e3, "unit": "ms", "cores": 1, "kind": "port",
                       "sample": "prove_succinct of a reference-shaped depth-%d circuit (2^%d inputs: commitment + GKR + two openings, the openings "
                                 "against blown-up quotients as the reference does); the reference's dense wiring tables cannot hold a 2^%d-wide "
                                 "layer; oracle C restatement, 1 thread" % (d, d, w)}
            except Exception as ex:       # the baseline must never cost the line
                cpu = {"error": repr(ex)}
        line = {"metric": "succinct_gkr_prove_ms", "value": ms, "unit": "ms", "n_gpus": world, "steps": n_steps, "warmup": n_warm, "ms_per_step": ms,
                "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
                "dtype": "u256 scalars / u384 curve coordinates (u32 Montgomery limbs, integer IMAD arithmetic)", "data": "synthetic",
                "config": {"workload": "succinct: " + desc, "field": fname, "curve": "BLS12-381 G1", "depth": depth, "width_log2": w,
                           "sumcheck_rounds": rounds, "circuit_setup_s": circuit_s, "trusted_setup_s": setup_s,
                           "sharding": ("none" if world == 1 else "circuit, setup, input layer and transcript replicated; layer sumchecks sharded on the low index bits, "
                                        "every large multi-scalar multiplication by contiguous shares of the points, across %d ranks" % world),
                           "timer": "host wall clock around prove_succinct (commit + GKR + two openings; each part ends on the host), max over ranks"},
                "roofline": None,
                "roofline_note": "three kernel families with different bounds: see the gkr_wide line (round latency / HBM) and the kzg line (integer multiplier)",
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": (32 << w), "d2h_bytes_per_step": int(rounds * 4 * 32 + (2 * w + 1) * 96),
                        "call": "gkr.prove_succinct with the input layer in pinned host memory"},
                "gpu_launches": launches, "clocks": clocks, "verified": verified, "verify_s": verify_s,
                "verified_by": "after the timed region: verify_succinct (succinct_gkr_protocol.rs:172-283): every layer's sumcheck and claim check on the GPU, "
                               "the two input openings by the pairing check on the host; and again with the opened values bound to the last sumcheck claim",
                "result_digest": keccak_digest([np.ascontiguousarray(proof.input_polynomial_commitment.reshape(-1, 4)),
                                                np.ascontiguousarray(proof.input_rb_proof.proofs.reshape(-1, 4)),
                                                np.ascontiguousarray(proof.input_rc_proof.proofs.reshape(-1, 4)), proof.claimed_sum])}
    barrier()
    setup.release()
    circuit.close()
    ctx.close()
    return line


def run_reference(args, wl):
    field, fname, P, D, log2_default, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    log2 = args.log2 or log2_default
    sample_log2 = min(log2, args.cpu_log2)
    th = max(1, args.cpu_threads)
    for _ in range(args.warmup):
        cpu_prove_once(field, P, D, min(sample_log2, 16), th)
    times = [cpu_prove_once(field, P, D, sample_log2, th) for _ in range(args.steps)]
    t = statistics.mean(times)
    value = (1 << sample_log2) / t / MEGA
    sample = "the first 2^%d entries of the 2^%d workload's own seeded tables, oracle C restatement of the reference prover%s, %s" % (
        sample_log2, log2, " (posed as f*g + 0*0, the only form the reference accepts)" if (P == 1 and D > 1) else "",
        "1 thread (the reference is single-threaded)" if th == 1 else "%d OpenMP threads (--cpu-threads: stronger than the single-threaded reference)" % th)
    line = {
        "impl": "reference", "metric": "sumcheck_prove_Melems_per_s", "value": value, "unit": "Melems/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u256 (4x u64 Montgomery limbs)",
        "data": "synthetic",
        "config": config_dict(args.workload, wl, log2, args.gpus, sample_log2),
        "cpu_baseline": {"value": value, "unit": "Melems/s", "cores": th, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Melems/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "extrapolation": "the CPU prover's passes are linear in the table length: Melems/s of the 2^%d sample stands for the 2^%d workload "
                         "(optimistic for the CPU: the sample is more cache-friendly)" % (sample_log2, log2),
    }
    return line


def config_dict(name, wl, log2, gpus, sample_log2):
    """identical in both arms (the driver compares them); sample_log2 = size of the CPU arm's bounded sample"""
    field, fname, P, D, _, desc = wl
    return {"workload": "%s: %s" % (name, desc), "field": fname, "log2_entries": log2, "sample_log2": sample_log2, "tables": P * D, "P": P, "D": D,
            "table_bytes_total": (P * D) << (log2 + 5), "sharding": "low index bits across %d rank(s)" % gpus,
            "l2": "inputs regenerated in HBM between steps (untimed); tables are %s than the 126 MB L2"
                  % ("larger" if ((P * D) << (log2 + 5)) // max(gpus, 1) > 126e6 else "NOT larger"),
            "seed": SEED, "claimed_sum": "the true sum of the seeded tables (computed before the timed region)"}


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region.  NVML in a background thread every ~5 ms (so that
    even millisecond-long timed regions see samples); falls back to an `nvidia-smi -lms 100` child process."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.sm, self.mx, self.pw, self.reasons = [], [], [], set()
        self.thread = self.p = None
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                         "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                         "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                         "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            self.stop_flag = threading.Event()
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        self.path = "/tmp/zk_clocks_%d.csv" % os.getpid()
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.f = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mx.append(self.max_sm)
                try:
                    self.pw.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:
                    pass
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    for nm, bit in self.bits.items():
                        if mask & bit:
                            self.reasons.add(nm)
                except Exception:
                    pass
            except Exception:
                break
            self.stop_flag.wait(0.005)

    def _summary(self, source):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "power_w_max": max(self.pw) if self.pw else None, "samples": len(self.sm), "reasons": sorted(self.reasons),
                "source": source}

    def stop(self):
        if self.thread:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            return self._summary("nvml, 5 ms period, sampled during the timed region")
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                self.sm.append(float(parts[0])); self.mx.append(float(parts[1])); self.pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(self.NAMES, parts[3:7]):
                if v.lower().startswith("active"):
                    self.reasons.add(nm)
        os.unlink(self.path)
        return self._summary("nvidia-smi -lms 100")


# ----------------------------------------------------------------------------------------- GPU arm
def run_ours(args, wl, name=None, log2=None, steps=None, warmup=None):
    import torch
    import torch.distributed as dist
    import zk_cryptography_research_implementations_b200 as zk
    from zk_cryptography_research_implementations_b200 import sharded
    from zk_cryptography_research_implementations_b200.core import _ptr
    from zk_cryptography_research_implementations_b200.transcripts import Transcript

    field, fname, P, D, log2_default, desc = wl
    name = name or args.workload
    log2 = log2 or (args.log2 if name == args.workload else 0) or log2_default
    n_steps = steps or args.steps
    n_warm = args.warmup if warmup is None else warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    ctx = zk.Context(field, local_rank, stream=stream)
    if world > 1:
        sharded.init_comm(ctx)
    lib = ctx.lib
    T = P * D
    N = 1 << log2
    m = N // world                     # local entries per table
    if m < 1:
        raise SystemExit("table smaller than the number of ranks")
    n_rounds = log2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident tables (this rank's shard), generated on the device
    tabs = [ctx.generate(SEED, i, m, first=rank, step=world) for i in range(T)]
    handles = [t.release() for t in tabs]
    arr = (C.c_void_p * T)(*handles)
    sp = C.c_void_p()
    ctx.check(lib.zk_sumpoly_create(ctx.h, arr, P, D, C.byref(sp)))

    def regenerate():
        for i in range(T):
            ctx.check(lib.zk_table_regenerate(ctx.h, lib.zk_sumpoly_table(sp, i), SEED, i, m, rank, world))

    coeffs = np.zeros((n_rounds, D + 1, 4), dtype=np.uint64)
    chal = np.zeros((n_rounds, 4), dtype=np.uint64)
    fin = np.zeros((T, 4), dtype=np.uint64)
    rpolys = np.zeros((n_rounds, 2, 4), dtype=np.uint64)
    claimed = np.zeros(4, dtype=np.uint64)

    def field_sum_over_ranks(elems):
        """sum over ranks, in the field, of an (k, 4) array of elements (host; only used outside the timed region)"""
        acc = np.ascontiguousarray(elems, dtype=np.uint64).reshape(-1, 4).copy()
        if world == 1:
            return acc
        t = torch.from_numpy(acc.view(np.int64)).cuda()
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        parts = [g.cpu().numpy().view(np.uint64) for g in gathered]
        out = np.zeros_like(acc)
        for part in parts:
            for k in range(out.shape[0]):
                out[k] = zk.fe_binop("add", field, out[k], part[k])
        return out

    if D > 1:
        # the TRUE claimed sum of the seeded tables, so that the reference's verifier accepts the timed proofs:
        # s(0) + s(1) of round 0 from the round-0 kernel, summed over the ranks' shards
        ev = np.zeros((D + 1, 4), dtype=np.uint64)
        ctx.check(lib.zk_sumcheck_round_evals(ctx.h, sp, _ptr(ev)))
        ev = field_sum_over_ranks(ev)
        claimed[:] = zk.fe_binop("add", field, ev[0], ev[1])

    def prove_resident():
        if D == 1:
            # plain sumcheck rounds; the 32*N-byte Keccak absorb of the table is host transcript work,
            # reported separately (absorb_ms) and included in e2e
            if world > 1:
                tr = Transcript()
                ctx.check(lib.zk_prove_basic_sharded(ctx.h, lib.zk_sumpoly_table(sp, 0), tr.h, _ptr(claimed), _ptr(rpolys), _ptr(chal),
                                                     _ptr(fin), 4 if args.nccl_exchange else 0, args.collapse_len))
            else:
                ctx.check(lib.zk_prove_basic_device(ctx.h, lib.zk_sumpoly_table(sp, 0), _ptr(claimed), _ptr(rpolys), _ptr(chal),
                                                    _ptr(fin), 2))
        else:
            tr = Transcript()
            ctx.check(lib.zk_prove_product_sharded(ctx.h, sp, _ptr(claimed), tr.h, _ptr(coeffs), _ptr(chal), _ptr(fin),
                                                   4 if args.nccl_exchange else 0, args.collapse_len))

    if D == 1 and world > 1:
        args.no_e2e = True   # the end-to-end plain prove absorbs the whole table through one host sponge: single-GPU only

    def timed_steps(fn, prep, steps, warmup, profile):
        for _ in range(warmup):
            prep(); barrier(); fn(); barrier()
        ctx.set_profiling(profile)
        ctx.reset_stats()
        times = []
        for _ in range(steps):
            prep()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            barrier()
            times.append(e0.elapsed_time(e1))
        st = ctx.stats()
        ctx.set_profiling(False)
        t = torch.tensor([statistics.mean(times)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), st

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, st = timed_steps(prove_resident, regenerate, n_steps, n_warm, True)
    clocks = sampler.stop() if sampler else None
    value = N / (ms * 1e-3) / MEGA
    proof_digest = keccak_digest([coeffs, chal, fin] if D > 1 else [claimed, rpolys, chal, fin])

    # ---- after the timed region: is the proof of the last timed step a proof?  (every rank holds the same one)
    #  1. the reference's verifier accepts it (sumcheck_gkr_protocol.rs:69-106 / verifier.rs:23-65 replayed on the host);
    #  2. the prover's final folded values are the tables' MLE evaluations at the challenges, recomputed from freshly
    #     regenerated tables by the separate evaluate kernels (sharded over the same ranks), and
    #  3. the verifier's last claim equals sum_p prod_d of them (the oracle check the reference's caller performs).
    verify = {}
    regenerate()
    evals_at_r = np.zeros((T, 4), dtype=np.uint64)
    for i in range(T):
        ctx.check(lib.zk_mle_evaluate_sharded(ctx.h, lib.zk_sumpoly_table(sp, i), _ptr(chal), n_rounds, _ptr(evals_at_r[i])))
    verify["final_values_are_mle_evaluations"] = bool(np.array_equal(evals_at_r, fin))
    if D > 1:
        ok = C.c_int(0)
        ch_v = np.zeros((n_rounds, 4), dtype=np.uint64)
        last = np.zeros(4, dtype=np.uint64)
        tr_v = Transcript()
        rc = lib.zk_verify_product(field, _ptr(claimed), _ptr(coeffs), n_rounds, D, tr_v.h, _ptr(ch_v), _ptr(last), C.byref(ok))
        verify["reference_verifier_accepts"] = bool(rc == 0 and ok.value == 1 and np.array_equal(ch_v, chal))
        total = np.zeros(4, dtype=np.uint64)
        for pi in range(P):
            prod = fin[pi * D]
            for d in range(1, D):
                prod = zk.fe_binop("mul", field, prod, fin[pi * D + d])
            total = zk.fe_binop("add", field, total, prod)
        verify["last_claim_matches_final_values"] = bool(np.array_equal(total, last))
    else:
        # plain sumcheck: the round sums telescope and the last claim is the table's evaluation (verifier.rs:44-70); the
        # challenges themselves are only reproducible with the table absorb, which `value` leaves out (see config.note)
        claim = claimed.copy()
        tele = True
        for k in range(n_rounds):
            tele = tele and np.array_equal(zk.fe_binop("add", field, rpolys[k, 0], rpolys[k, 1]), claim)
            diff = zk.fe_binop("sub", field, rpolys[k, 1], rpolys[k, 0])
            claim = zk.fe_binop("add", field, rpolys[k, 0], zk.fe_binop("mul", field, chal[k], diff))
        verify["round_sums_telescope"] = bool(tele)
        verify["last_claim_matches_final_values"] = bool(np.array_equal(claim, fin[0]))
    verified = all(verify.values())

    # ---- roofline of the dominant kernel family (round kernels), CUDA events on the launching stream
    hbm_peak, peak_src = peaks()
    achieved = st["round_bytes"] / (st["round_ms"] * 1e-3) / 1e9 if st["round_ms"] > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": None, "peak_source": peak_src,
                "kernel": "fold_evals_kernel / round_evals_kernel (+ the persistent sumcheck_rounds_kernel launch of each prove) <%s,P=%d,D=%d> (all %d launches of %d steps, rank 0)"
                          % (fname, P, D, st["round_launches"], args.steps),
                "algorithmic_bytes_per_step_per_rank": st["round_bytes"] / max(args.steps, 1),
                "kernel_ms_per_step": st["round_ms"] / max(args.steps, 1)}
    launches = st["launches"]
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture of this round (one fold_evals launch of
    # the f*g workload, old table length 2^26: 48 * T * 2^26 algorithmic bytes) -- bytes moved per launch, to set against them.
    # (ncu cannot run inside the timed run: a number measured under a profiler is never a bench value.)
    if P == 1 and D == 2:
        cap = os.path.join(ROOT, "profiles", "r02", "fold_evals_2p27_ncu_full_summary.csv")
        try:
            vals = {}
            import csv
            for row in csv.reader(open(cap)):
                if len(row) == 4 and row[1] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    vals[row[1]] = float(row[3]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[row[2]]
            if len(vals) == 2:
                roofline["traffic"] = sum(vals.values())
                roofline["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum of ONE fold_evals_kernel<BN254_FQ,1,2> launch folding 2 x 2^26 "
                                            "entries (ncu --set full of `bench.py --log2 27`, %s, not measured in this run): algorithmic bytes of that "
                                            "launch %.4g" % (os.path.relpath(cap, ROOT), 48.0 * 2 * (1 << 26)))
        except OSError:
            pass

    # ---- integer-multiply ceiling (register-resident probe, same clocks)
    integer = {}
    if rank == 0 and not args.no_probe:
        for kind, probe in ((0, "mont_mul"), (1, "fold_by_scalar"), (2, "mul_acc_unreduced")):
            ops, pms = C.c_double(), C.c_double()
            ctx.check(lib.zk_arith_probe(ctx.h, kind, 1500, 2, C.byref(ops), C.byref(pms)))
            integer[probe + "_Gops"] = ops.value / 1e9
        # IMAD.WIDE budget of one prove (per rank): round 0 evaluates (D+1) P products per pair (64 IMAD.WIDE each, D = 2)
        # or none (D = 1); each later round folds 2T entries per new pair (84 each) and evaluates D P products (s(1) is
        # derived) -- summed over the rounds: N/2 pairs in round 0, N/2 new pairs in all later rounds together.
        prod0 = (D + 1) * P * 64 if D >= 2 else 0
        prodk = D * P * 64 if D >= 2 else 0
        imad = (m / 2.0) * prod0 + (m / 2.0) * (2 * T * 84 + prodk)
        pipe = integer["mont_mul_Gops"] * 1e9 * 137          # IMAD.WIDE/s the probe sustains (137 per Montgomery product)
        integer.update({"imad_wide_per_prove_per_rank": imad, "imad_wide_per_s_measured": pipe, "imad_floor_ms": imad / pipe * 1e3,
                        "hbm_floor_ms": st["round_bytes"] / max(args.steps, 1) / (hbm_peak * 1e9) * 1e3,
                        "note": "frac is reported against the slower (larger-time) of the two floors: HBM here"})

    # ---- e2e: the same prove through the C-ABI from pinned HOST tables (H2D of the inputs every step)
    e2e = None
    if not args.no_e2e:
        host = []
        for i in range(T):
            p = C.c_void_p()
            if lib.zk_pinned_alloc(C.c_size_t(m * 32), C.byref(p)) != 0:
                for q in host:
                    lib.zk_pinned_free(q)
                host = None
                break
            host.append(p)
    if not args.no_e2e and host is None:
        e2e = {"value": None, "unit": "Melems/s", "unavailable": "cudaHostAlloc of %d x %d bytes of pinned host memory failed on this box" % (T, m * 32)}
    elif not args.no_e2e:
        regenerate()
        for i in range(T):
            ctx.check(lib.zk_table_download(ctx.h, lib.zk_sumpoly_table(sp, i), C.cast(host[i], C.POINTER(C.c_uint64))))

        def upload_and_prove():
            for i in range(T):
                ctx.check(lib.zk_table_upload_into(ctx.h, lib.zk_sumpoly_table(sp, i), host[i], m))
            if D == 1:
                ctx.check(lib.zk_prove_basic_device(ctx.h, lib.zk_sumpoly_table(sp, 0), _ptr(claimed), _ptr(rpolys), _ptr(chal),
                                                    _ptr(fin), 0))   # full prove(): table absorb included
            else:
                tr = Transcript()
                ctx.check(lib.zk_prove_product_sharded(ctx.h, sp, _ptr(claimed), tr.h, _ptr(coeffs), _ptr(chal), _ptr(fin), 0,
                                                       args.collapse_len))

        e_steps = max(1, min(n_steps, args.e2e_steps))
        ms_e2e, _ = timed_steps(upload_and_prove, lambda: None, e_steps, 1, False)
        d2h = (n_rounds * (D + 1 if D > 1 else 2) + T + 1) * 32
        e2e = {"value": N / (ms_e2e * 1e-3) / MEGA, "unit": "Melems/s", "h2d_bytes_per_step": T * m * 32 * world,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e, "steps": e_steps,
               # the step is the upload followed by the prove on one stream: what the host link delivered per GPU while it ran
               # (plain sumcheck: the e2e step is dominated by the serial Keccak absorb of the table, not by the link)
               "pcie_h2d_GBps_per_gpu": (T * m * 32 / max(ms_e2e - ms, 1e-6) / 1e6) if D > 1 else None,
               "pcie_note": "h2d bytes of one rank / (e2e ms - resident prove ms); all ranks upload concurrently from pinned host memory",
               "call": "zk_table_upload_into (pinned host -> HBM) + %s" % ("zk_prove_basic_device incl. the Keccak absorb of the table"
                                                                           if D == 1 else "zk_prove_product_sharded")}
        for p in host:
            lib.zk_pinned_free(p)

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    cpu_all = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample_log2 = min(log2, args.cpu_log2)
        t = cpu_prove_once(field, P, D, sample_log2)
        cpu = {"value": (1 << sample_log2) / t / MEGA, "unit": "Melems/s", "cores": 1, "kind": "port",
               "sample": "one prove of a 2^%d-entry sample (%.1f s) by oracle/zkoracle.c -- C restatement of the single-threaded "
                         "reference prover with its pass structure; host has %d cores" % (sample_log2, t, os.cpu_count() or 0)}
        try:   # extra, labelled line: must never cost the main line
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import coracle as co
            ncores = os.cpu_count() or 1
            if co.openmp_enabled() and ncores > 1:
                ta = cpu_prove_once(field, P, D, sample_log2, threads=ncores)
                cpu_all = {"value": (1 << sample_log2) / ta / MEGA, "unit": "Melems/s", "cores": ncores, "kind": "port",
                           "sample": "same 2^%d-entry sample (%.2f s) with the oracle's data-parallel loops on %d OpenMP threads -- STRONGER than the "
                                     "reference, which is single-threaded (no rayon, Cargo.lock:89-105); the per-round transcript stays serial"
                                     % (sample_log2, ta, ncores)}
        except Exception as ex:   # pragma: no cover
            cpu_all = {"value": None, "unavailable": repr(ex)}

    line = None
    if rank == 0:
        line = {
            "metric": "sumcheck_prove_Melems_per_s", "value": value, "unit": "Melems/s", "n_gpus": world,
            "steps": n_steps, "warmup": n_warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u256 (8x u32 Montgomery limbs, integer IMAD arithmetic)", "data": "synthetic",
            "config": config_dict(name, wl, log2, world, min(log2, args.cpu_log2)),
            "roofline": roofline, "integer_roofline": integer, "cpu_baseline": cpu, "cpu_baseline_all_cores": cpu_all, "e2e": e2e,
            "gpu_launches": launches,
            "exchange": (None if world == 1 else ("ncclAllGather per round" if args.nccl_exchange else ctx_exchange_name(ctx))),
            "clocks": clocks, "tail_log": ctx.tail_log(),
            "verified": verified, "verify": verify,
            "verified_by": "after the timed region, on the proof of the last timed step: zk_verify_product (the reference's verifier) + "
                           "zk_mle_evaluate_sharded of freshly regenerated tables at the challenges",
            "proof_digest": proof_digest, "proof_digest_kind": "Keccak-256 of coefficients | challenges | final values (identical at every N)",
        }
        if D == 1:
            line["note"] = ("value = the n fused rounds, table already absorbed (ZK_FLAG_SKIP_ABSORB: measurement only); "
                            "e2e = full Prover::prove from a host table incl. the 32*N-byte serial Keccak absorb")
    lib.zk_sumpoly_free(ctx.h, sp)
    barrier()
    ctx.close()
    return line


def ctx_exchange_name(ctx):
    try:
        from zk_cryptography_research_implementations_b200 import sharded
        return sharded.exchange_kind(ctx)
    except Exception:
        return "shared-memory mailboxes written by the round kernels (no per-round collective)"


def run_extras(args):
    """The rest of BASELINE.json's metric in the driver's own record: short runs of configs[3] (GKR 16 x 2^22), configs[1]
    (plain sumcheck 2^24) and configs[4] (MLE sweep), each a full line with its own roofline / cpu_baseline / e2e / clocks.
    N = 1: all three; N > 1: the sharded MLE point (2^32 at N = 8).  A failing extra is reported, never fatal."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    extras = []

    def attempt(label, fn):
        try:
            line = fn()
            if line is not None:
                extras.append(line)
        except Exception as ex:   # pragma: no cover -- must never cost the headline line
            if world > 1:
                raise             # a rank that bails out of a collective path alone would hang the others
            extras.append({"workload": label, "error": repr(ex)})

    if world == 1:
        attempt("gkr_wide", lambda: run_gkr_wide(args, WORKLOADS["gkr_wide"], steps=min(args.steps, 8), warmup=min(args.warmup, 3)))
        attempt("plain24", lambda: run_ours(args, WORKLOADS["plain24"], name="plain24", steps=min(args.steps, 10), warmup=min(args.warmup, 3)))
        attempt("mle", lambda: run_mle(args, WORKLOADS["mle"], log2=28, sweep="20,22,24,26,30", steps=min(args.steps, 5), warmup=min(args.warmup, 3)))
        attempt("kzg", lambda: run_kzg(args, WORKLOADS["kzg"], log2=22, steps=min(args.steps, 4), warmup=min(args.warmup, 2)))
        attempt("succinct", lambda: run_succinct(args, WORKLOADS["succinct"], log2=22, steps=min(args.steps, 3), warmup=1))
    else:
        lg = min(32, 29 + world.bit_length())       # 2^31 at 2 ranks, 2^32 at 4 and 8 (configs[4]: 2^32 sharded over 8 GPUs)
        attempt("mle", lambda: run_mle(args, WORKLOADS["mle"], log2=lg, sweep="", steps=min(args.steps, 5), warmup=min(args.warmup, 3)))
        attempt("gkr_wide", lambda: run_gkr_wide(args, WORKLOADS["gkr_wide"], steps=min(args.steps, 5), warmup=min(args.warmup, 2)))   # configs[3]: "on 8 x B200"
        attempt("succinct", lambda: run_succinct(args, WORKLOADS["succinct"], log2=22, steps=min(args.steps, 3), warmup=1))
    return extras


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="product30", choices=sorted(WORKLOADS))
    ap.add_argument("--log2", type=int, default=0, help="override log2(entries per table)")
    ap.add_argument("--collapse-len", type=int, default=1 << 12, dest="collapse_len")
    ap.add_argument("--cpu-log2", type=int, default=21, dest="cpu_log2", help="size of the CPU baseline sample")
    ap.add_argument("--cpu-threads", type=int, default=1, dest="cpu_threads",
                    help="--impl reference: OpenMP threads of the oracle (default 1: the reference is single-threaded)")
    ap.add_argument("--cpu-depth", type=int, default=7, dest="cpu_depth", help="circuit depth of the CPU GKR baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=3, dest="e2e_steps")
    ap.add_argument("--nccl-exchange", action="store_true", dest="nccl_exchange",
                    help="per-round partial exchange with ncclAllGather instead of the shared-memory mailboxes")
    ap.add_argument("--sweep", default="", help="mle workload: comma-separated extra log2 sizes measured in the same run")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-probe", action="store_true")
    ap.add_argument("--no-extras", action="store_true", dest="no_extras",
                    help="default workload only: skip the short GKR / plain-sumcheck / MLE runs appended as `extra_workloads`")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return                   # the CPU arm runs on rank 0 alone; the other ranks exit 0 without work
        fn = {"gkr": run_gkr, "gkr_wide": run_gkr_wide, "mle": run_mle, "kzg": run_kzg, "succinct": run_succinct}.get(args.workload, run_reference)
        print(json.dumps(fn(args, wl)), flush=True)
        return
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    fn = {"gkr": run_gkr, "gkr_wide": run_gkr_wide, "mle": run_mle, "kzg": run_kzg, "succinct": run_succinct}.get(args.workload, run_ours)
    line = fn(args, wl)
    if args.workload == "product30" and not args.log2 and not args.no_extras:
        extras = run_extras(args)
        if line is not None:
            line["extra_workloads"] = extras
    if line is not None:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
