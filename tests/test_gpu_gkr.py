"""GKR prover on the GPU against the oracle's dense restatement of gkr_protocol::prove, bit for bit, and
verified by the restated reference verifier."""
import random

import numpy as np
import pytest

from conftest import FIELDS

pytestmark = pytest.mark.gpu


def make_circuit(zk, fid, layers):
    from zk_cryptography_research_implementations_b200.circuit import Circuit, Gate, Layer
    return Circuit.new(fid, [Layer.new([Gate.new(*g) for g in l]) for l in layers])


def check_against_oracle(zk, co, ctx, fid, layers, inputs_ints, verify=True):
    from zk_cryptography_research_implementations_b200 import gkr
    circuit = make_circuit(zk, fid, layers)
    I = zk.fe_from_ints(fid, inputs_ints)
    proof = gkr.prove(ctx, circuit, I)
    oc = co.Circuit(layers)
    want = co.gkr_prove(fid, oc, I)
    assert np.array_equal(proof.circuit_output, want.circuit_output)
    assert np.array_equal(proof.claimed_sum, want.claimed_sum)
    got_coeffs = np.concatenate([np.stack([p.coefficients for p in sp.round_univariate_polynomials]) for sp in proof.sumcheck_proofs])
    assert np.array_equal(got_coeffs, want.coeffs)
    assert np.array_equal(np.concatenate([sp.random_challenges for sp in proof.sumcheck_proofs]), want.challenges)
    assert np.array_equal(np.stack([sp.claimed_sum for sp in proof.sumcheck_proofs]), want.layer_claims)
    L = len(layers)
    assert np.array_equal(proof.wb_evaluations, want.wb[: L - 1]) and np.array_equal(proof.wc_evaluations, want.wc[: L - 1])
    if verify:
        assert co.gkr_verify(fid, oc, want, I)      # the reference verifier accepts the (identical) proof
    return proof


def test_gkr_reference_circuits_and_golden(zk, co, ctx_for, golden):
    # gkr_protocol.rs:246-299, the reference's own two circuits
    for e in golden["reference_kats"]["gkr_round_trips"]:
        fid = FIELDS[e["field"]]
        check_against_oracle(zk, co, ctx_for(fid), fid, e["layers"], e["inputs"])
    # SURVEY appendix B vector
    g = golden["appendix_b"]["gkr"]
    fid = FIELDS[g["field"]]
    proof = check_against_oracle(zk, co, ctx_for(fid), fid, g["layers"], g["inputs"])
    assert zk.fe_to_ints(fid, proof.circuit_output) == g["output"]
    assert [zk.fe_to_ints(fid, p.coefficients) for p in proof.sumcheck_proofs[0].round_univariate_polynomials] == g["layer0_coeffs"]
    assert zk.fe_to_ints(fid, proof.wb_evaluations) == g["wb"] and zk.fe_to_ints(fid, proof.wc_evaluations) == g["wc"]
    assert zk.fe_to_ints(fid, proof.claimed_sum) == [g["claimed_sum"]]
    for e in golden["generated"]["gkr"]:
        fid = FIELDS[e["field"]]
        proof = check_against_oracle(zk, co, ctx_for(fid), fid, e["layers"], e["inputs"])
        assert zk.fe_to_ints(fid, proof.claimed_sum) == [e["claimed_sum"]]
        assert zk.fe_to_ints(fid, proof.wb_evaluations) == e["wb"]


@pytest.mark.parametrize("fid", [0, 2])
@pytest.mark.parametrize("depth", [1, 2, 4, 6])
def test_gkr_random_reference_shaped_circuits(zk, co, ctx_for, fid, depth):
    """layer i: 2^i outputs (2 at layer 0 for depth 1 variants), inputs of 2^(i+1) wires; several gates may feed one
    output, one (b,c) pair may feed several outputs, both operators; edge inputs 0 / p-1."""
    import pyoracle as po
    p = po.P[{v: k for k, v in FIELDS.items()}[fid]]
    rng = random.Random(100 * fid + depth)
    layers = []
    for i in range(depth):
        n_out = 1 << i
        gates, seen = [], set()
        for o in range(n_out):
            for _ in range(rng.choice([1, 1, 2, 3])):
                g = (rng.randrange(1 << (i + 1)), rng.randrange(1 << (i + 1)), o, rng.randrange(2))
                if g not in seen:
                    seen.add(g)
                    gates.append(g)
        # make sure the widest wire index is used so the layer below has full width
        layers.append(gates)
    # every layer must produce exactly 2^i values: guaranteed since each output index 0..2^i-1 has a gate
    inputs = [rng.randrange(p) for _ in range(1 << depth)]
    inputs[0] = 0
    inputs[-1] = p - 1
    check_against_oracle(zk, co, ctx_for(fid), fid, layers, inputs)


def test_gkr_two_outputs_and_duplicate_gates(zk, co, ctx_for):
    # layer 0 with two outputs; a duplicated gate (the reference's dense indicator stores `= one`, the evaluation `+=`)
    layers = [[(0, 1, 0, 1), (1, 0, 1, 0)], [(0, 1, 0, 0), (2, 3, 1, 1)]]
    check_against_oracle(zk, co, ctx_for(0), 0, layers, [2, 3, 4, 5])
    # with a duplicated gate the reference is unsound (its own verifier rejects its own proof); the prover output must
    # still be the reference's, limb for limb
    layers = [[(0, 1, 0, 1), (1, 0, 1, 0)], [(0, 1, 0, 0), (2, 3, 1, 1), (2, 3, 1, 1)]]
    check_against_oracle(zk, co, ctx_for(0), 0, layers, [2, 3, 4, 5], verify=False)


def test_circuit_kats_through_the_mirror_api(zk, co, ctx_for, golden):
    from zk_cryptography_research_implementations_b200.circuit import num_of_layer_variables
    k = golden["reference_kats"]
    ctx = ctx_for(0)
    for e in k["circuit_evaluate"]:
        c = make_circuit(zk, 0, e["layers"])
        res = c.evaluate(zk.fe_from_ints(0, e["inputs"]))
        if "layer_evaluations" in e:
            assert [zk.fe_to_ints(0, x) for x in res.layer_evaluations] == e["layer_evaluations"]
        else:
            assert zk.fe_to_ints(0, res.output) == e["output"]
    for i, v in k["num_of_layer_variables"]["values"]:
        assert num_of_layer_variables(i) == v
    for e in k["add_i_mul_i"]:
        c = make_circuit(zk, 0, e["layers"])
        a, m = c.add_i_and_mul_i_mle(ctx, e["layer"])
        av, mv = zk.fe_to_ints(0, a.evaluated_values), zk.fe_to_ints(0, m.evaluated_values)
        assert len(av) == e["size"] and [i for i, v in enumerate(av) if v] == e["add_ones"] and [i for i, v in enumerate(mv) if v] == e["mul_ones"]


# ------------------------------------------------------------------------------------------ sparse two-phase prover
def _random_layers(rng, bits, gates_per_out=(1, 1, 2)):
    layers = []
    for li in range(len(bits) - 1):
        gates, seen = [], set()
        for o in range(1 << bits[li]):
            for _ in range(rng.choice(gates_per_out)):
                g = (rng.randrange(1 << bits[li + 1]), rng.randrange(1 << bits[li + 1]), o, rng.randrange(2))
                if g not in seen:
                    seen.add(g)
                    gates.append(g)
        layers.append(gates)
    return layers


def _compare_wide(zk, fid, proof, want_coeffs, want_chal, want_claims, want_wb, want_wc, want_claimed):
    got_coeffs = np.concatenate([np.stack([p.coefficients for p in sp.round_univariate_polynomials]) for sp in proof.sumcheck_proofs])
    assert zk.fe_to_ints(fid, got_coeffs) == want_coeffs
    assert zk.fe_to_ints(fid, np.concatenate([sp.random_challenges for sp in proof.sumcheck_proofs])) == want_chal
    assert zk.fe_to_ints(fid, np.stack([sp.claimed_sum for sp in proof.sumcheck_proofs])) == want_claims
    assert zk.fe_to_ints(fid, proof.wb_evaluations) == want_wb and zk.fe_to_ints(fid, proof.wc_evaluations) == want_wc
    assert zk.fe_to_ints(fid, proof.claimed_sum) == [want_claimed]


@pytest.mark.parametrize("fid", [0, 2])
@pytest.mark.parametrize("depth", [1, 2, 3, 5, 6])
def test_wide_prover_equals_dense_reference_on_reference_shapes(zk, co, ctx_for, fid, depth):
    """the sparse two-phase prover must reproduce the reference's dense proof limb for limb (oracle: the C
    restatement of gkr_protocol::prove) and the reference verifier must accept it"""
    import pyoracle as po
    from zk_cryptography_research_implementations_b200 import gkr
    p = po.P[{v: k for k, v in FIELDS.items()}[fid]]
    rng = random.Random(500 + 10 * fid + depth)
    bits = [1] + list(range(1, depth + 1))
    layers = _random_layers(rng, [0] + bits[1:])      # layer 0: a single output gate set (index 0)
    inputs = [rng.randrange(p) for _ in range(1 << depth)]
    ctx = ctx_for(fid)
    I = zk.fe_from_ints(fid, inputs)
    wc = gkr.WideCircuit.reference_shaped(ctx, layers)
    proof = gkr.prove_wide(ctx, wc, I)
    oc = co.Circuit(layers)
    want = co.gkr_prove(fid, oc, I)
    assert co.gkr_verify(fid, oc, want, I)
    got_coeffs = np.concatenate([np.stack([q.coefficients for q in sp.round_univariate_polynomials]) for sp in proof.sumcheck_proofs])
    assert np.array_equal(got_coeffs, want.coeffs)
    assert np.array_equal(np.concatenate([sp.random_challenges for sp in proof.sumcheck_proofs]), want.challenges)
    assert np.array_equal(np.stack([sp.claimed_sum for sp in proof.sumcheck_proofs]), want.layer_claims)
    assert np.array_equal(proof.wb_evaluations, want.wb[: depth - 1]) and np.array_equal(proof.wc_evaluations, want.wc[: depth - 1])
    assert np.array_equal(proof.claimed_sum, want.claimed_sum)
    assert np.array_equal(proof.circuit_output[: want.circuit_output.shape[0]], want.circuit_output)
    # and the dense device prover gives the same proof
    dense = gkr.prove(ctx, make_circuit(zk, fid, layers), I)
    assert np.array_equal(np.concatenate([np.stack([q.coefficients for q in sp.round_univariate_polynomials]) for sp in dense.sumcheck_proofs]), got_coeffs)


@pytest.mark.parametrize("bits", [[2, 3, 3, 2], [3, 3, 3, 3], [1, 4, 2, 5], [4, 1, 3]])
def test_wide_prover_on_general_shapes(zk, co, ctx_for, bits):
    """layer shapes the reference cannot express: checked against the generalised dense Python model"""
    import pyoracle as po
    from zk_cryptography_research_implementations_b200 import gkr
    fid = 0
    p = po.P["BN254_FQ"]
    rng = random.Random(sum(b * 7 ** i for i, b in enumerate(bits)))
    layers = _random_layers(rng, bits)
    inputs = [rng.randrange(p) for _ in range(1 << bits[-1])]
    want = po.gkr_prove_general([[po.Gate(*g) for g in l] for l in layers], bits, inputs, p)
    ctx = ctx_for(fid)
    proof = gkr.prove_wide(ctx, gkr.WideCircuit(ctx, bits, layers), zk.fe_from_ints(fid, inputs))
    _compare_wide(zk, fid, proof,
                  [c for (_, polys, _) in want.sumcheck_proofs for poly in polys for c in poly],
                  [c for (_, _, ch) in want.sumcheck_proofs for c in ch],
                  [cl for (cl, _, _) in want.sumcheck_proofs], want.wb_evaluations, want.wc_evaluations, want.claimed_sum)
    assert zk.fe_to_ints(fid, proof.circuit_output) == want.circuit_output


def test_wide_prover_reduction_layer_with_heavy_fan_in(zk, co, ctx_for):
    """the shape of bench.py's synthetic wide circuit in miniature: layer 0 folds all 2^7 wires into TWO outputs (64 gates
    each: the sliced evaluation kernels), the layer below is full width; checked against the generalised dense Python model"""
    import pyoracle as po
    from zk_cryptography_research_implementations_b200 import gkr
    fid = 0
    p = po.P["BN254_FQ"]
    rng = random.Random(4242)
    w = 7
    bits = [1, w, w]
    layer0 = [(g, rng.randrange(1 << w), g & 1, rng.randrange(2)) for g in range(1 << w)]
    layer1 = [(rng.randrange(1 << w), rng.randrange(1 << w), g, rng.randrange(2)) for g in range(1 << w)]
    layers = [layer0, layer1]
    inputs = [rng.randrange(p) for _ in range(1 << w)]
    want = po.gkr_prove_general([[po.Gate(*g) for g in l] for l in layers], bits, inputs, p)
    ctx = ctx_for(fid)
    for tail_log in (24, 13, 0):
        ctx.set_tail_log(tail_log)
        try:
            proof = gkr.prove_wide(ctx, gkr.WideCircuit(ctx, bits, layers), zk.fe_from_ints(fid, inputs))
        finally:
            ctx.set_tail_log(20)
        _compare_wide(zk, fid, proof,
                      [c for (_, polys, _) in want.sumcheck_proofs for poly in polys for c in poly],
                      [c for (_, _, ch) in want.sumcheck_proofs for c in ch],
                      [cl for (cl, _, _) in want.sumcheck_proofs], want.wb_evaluations, want.wc_evaluations, want.claimed_sum)
        assert zk.fe_to_ints(fid, proof.circuit_output) == want.circuit_output


def _wide_arrays(rng, w, depth):
    """bench.py's synthetic wide circuit in miniature: layer 0 reduces 2^w wires to two outputs, the others are full width"""
    n = 1 << w
    g = np.arange(n, dtype=np.int64)
    layers = [np.stack([g, rng.integers(0, n, size=n), g & 1, rng.integers(0, 2, size=n)], axis=1)]
    for _ in range(depth - 1):
        layers.append(np.stack([rng.integers(0, n, size=n), rng.integers(0, n, size=n), g, rng.integers(0, 2, size=n)], axis=1))
    return [1] + [w] * depth, layers


def _assert_equals_sparse_oracle(proof, want):
    got_coeffs = np.concatenate([np.stack([p.coefficients for p in sp.round_univariate_polynomials]) for sp in proof.sumcheck_proofs])
    assert np.array_equal(got_coeffs, want.coeffs[: got_coeffs.shape[0]]), "round polynomials differ from the gate-list oracle"
    assert np.array_equal(np.concatenate([sp.random_challenges for sp in proof.sumcheck_proofs]), want.challenges[: got_coeffs.shape[0]])
    assert np.array_equal(np.stack([sp.claimed_sum for sp in proof.sumcheck_proofs]), want.layer_claims)
    L = len(proof.sumcheck_proofs)
    assert np.array_equal(proof.wb_evaluations, want.wb[: L - 1]) and np.array_equal(proof.wc_evaluations, want.wc[: L - 1])
    assert np.array_equal(proof.claimed_sum, want.claimed_sum)
    assert np.array_equal(proof.circuit_output[: want.circuit_output.shape[0]], want.circuit_output)


@pytest.mark.parametrize("w,depth,fid", [(12, 4, 0), (12, 3, 2), (16, 3, 0)])
def test_wide_prover_limb_for_limb_against_the_gate_list_oracle(zk, co, ctx_for, w, depth, fid):
    """widths the dense oracles cannot reach (2^12, 2^16: multi-block round kernels, the row-form eq tables, the sliced
    evaluation of the reduction layer, the CSR orderings built on the GPU): every limb of the proof against
    oracle/zkoracle.c zko_gkr_prove_sparse -- gkr_protocol.rs:57-134 with add_i / mul_i evaluated from the gate list --
    which is itself pinned to the dense restatement on reference shapes (tests/test_oracle.py).  Both eq-table forms
    (row-form / entry-wise outer products, ZKB200_EQ_ROWS) must give that proof, and both verifiers (the CUDA one and the
    oracle's) must accept it."""
    import os
    from zk_cryptography_research_implementations_b200 import gkr
    ctx = ctx_for(fid)
    rng = np.random.default_rng(1000 * w + 10 * depth + fid)
    bits, layers = _wide_arrays(rng, w, depth)
    inputs = ctx.generate(11, 0, 1 << w).download()
    sc = co.SparseCircuit(bits, layers)
    want = co.gkr_prove_sparse(fid, sc, inputs)
    assert co.gkr_verify_sparse(fid, sc, want, inputs)
    # tail_log 13 / 9: the phase sumchecks run 5 / 9 host-driven rounds before the persistent launch, so the overlapped
    # gate-wise phase-2 precomputation (phase2_pre_kernel on the side stream, phase2_fin_kernel) is what builds A and B;
    # ZKB200_GKR_OVERLAP=0: the one-kernel build.  All must give the oracle's proof.
    # ZKB200_GKR_SEG=0: the wire-per-thread table builders instead of the warp-segmented gate-parallel ones.
    for knobs, tail_log in (({"ZKB200_EQ_ROWS": "1"}, 20), ({"ZKB200_EQ_ROWS": "0"}, 20), ({}, 13), ({}, 9), ({"ZKB200_GKR_OVERLAP": "0"}, 13),
                            ({"ZKB200_GKR_SEG": "0"}, 20), ({"ZKB200_GKR_SEG": "0"}, 13)):
        os.environ.update(knobs)
        ctx.set_tail_log(tail_log)
        try:
            wc = gkr.WideCircuit(ctx, bits, layers)
            proof = gkr.prove_wide(ctx, wc, inputs)
            _assert_equals_sparse_oracle(proof, want)
            assert gkr.verify_wide(ctx, wc, proof, inputs)
            wc.close()
        finally:
            ctx.set_tail_log(20)
            for k in knobs:
                del os.environ[k]


def test_wide_verifier_accepts_honest_and_rejects_tampered_proofs(zk, co, ctx_for):
    """zk_gkr_verify_wide (gkr_protocol.rs:146-236 with utils.rs:84-135 evaluated from the gate list on the GPU): decision
    parity with the oracle's gate-list verifier on honest and tampered proofs, inputs from the host and from HBM"""
    import copy
    from zk_cryptography_research_implementations_b200 import gkr
    fid = 0
    ctx = ctx_for(fid)
    rng = np.random.default_rng(77)
    bits, layers = _wide_arrays(rng, 10, 3)
    dev_inputs = ctx.generate(5, 1, 1 << 10)
    inputs = dev_inputs.download()
    wc = gkr.WideCircuit(ctx, bits, layers)
    sc = co.SparseCircuit(bits, layers)
    proof = gkr.prove_wide(ctx, wc, dev_inputs)

    def both(pf, inp=inputs):
        got = gkr.verify_wide(ctx, wc, pf, inp)
        coeffs = np.concatenate([np.stack([p.coefficients for p in sp.round_univariate_polynomials]) for sp in pf.sumcheck_proofs])
        want = co.gkr_verify_sparse(fid, sc, co.make_gkr_proof(sc, pf.circuit_output, pf.claimed_sum, np.stack([sp.claimed_sum for sp in pf.sumcheck_proofs]),
                                                                coeffs, pf.wb_evaluations, pf.wc_evaluations), inp)
        assert got == want
        return got

    assert both(proof)
    assert gkr.verify_wide(ctx, wc, proof, dev_inputs)
    one = zk.fe_from_int(fid, 1)
    t = copy.deepcopy(proof); t.circuit_output[0] = zk.fe_binop("add", fid, t.circuit_output[0], one); assert not both(t)
    t = copy.deepcopy(proof); t.wb_evaluations[1] = zk.fe_binop("add", fid, t.wb_evaluations[1], one); assert not both(t)
    t = copy.deepcopy(proof); c = t.sumcheck_proofs[2].round_univariate_polynomials[7].coefficients; c[1] = zk.fe_binop("add", fid, c[1], one); assert not both(t)
    t = copy.deepcopy(proof); t.sumcheck_proofs[1].claimed_sum[:] = zk.fe_binop("add", fid, t.sumcheck_proofs[1].claimed_sum, one); assert not both(t)
    bad_inputs = inputs.copy(); bad_inputs[3] = zk.fe_binop("add", fid, bad_inputs[3], one)
    assert not both(proof, bad_inputs)
    wc.close()


def test_wide_circuit_single_output_is_the_reference_padding_and_bad_gate_lists_are_refused(zk, co, ctx_for):
    from zk_cryptography_research_implementations_b200 import gkr
    import pyoracle as po
    fid = 0
    ctx = ctx_for(fid)
    rng = random.Random(31)
    depth = 4
    layers = _random_layers(rng, [0] + list(range(1, depth + 1)))
    inputs = zk.fe_from_ints(fid, [rng.randrange(po.P["BN254_FQ"]) for _ in range(1 << depth)])
    # layer_bits[0] == 0: one output, padded to [out, 0] with one challenge exactly like gkr_protocol.rs:43-51
    wc = gkr.WideCircuit(ctx, [0] + list(range(1, depth + 1)), layers)
    assert wc.layer_bits[0] == 1
    proof = gkr.prove_wide(ctx, wc, inputs)
    want = co.gkr_prove(fid, co.Circuit(layers), inputs)
    assert np.array_equal(np.concatenate([np.stack([q.coefficients for q in sp.round_univariate_polynomials]) for sp in proof.sumcheck_proofs]), want.coeffs)
    assert np.array_equal(proof.claimed_sum, want.claimed_sum)
    assert np.array_equal(proof.circuit_output[0], want.circuit_output[0]) and not proof.circuit_output[1].any()
    assert gkr.verify_wide(ctx, wc, proof, inputs)
    # a duplicated gate, an index outside its layer, an operator that is neither add nor mul: refused at construction
    with pytest.raises(zk.ZkError, match="duplicate gate"):
        gkr.WideCircuit(ctx, [1, 2], [[(0, 1, 0, 1), (2, 3, 1, 0), (0, 1, 0, 1)]])
    gkr.WideCircuit(ctx, [1, 2], [[(0, 1, 0, 1), (2, 3, 1, 0), (0, 1, 0, 0)]]).close()      # same wires, other operator: distinct
    with pytest.raises(zk.ZkError, match="does not fit its layer width"):
        gkr.WideCircuit(ctx, [1, 2], [[(0, 4, 0, 1)]])
    with pytest.raises(zk.ZkError, match="does not fit its layer width"):
        gkr.WideCircuit(ctx, [0, 2], [[(0, 1, 1, 1)]])
    with pytest.raises(zk.ZkError, match="operator"):
        gkr.WideCircuit(ctx, [1, 2], [[(0, 1, 0, 2)]])


def _host_layer(co, fid, layer, n_out, values):
    """Circuit::evaluate for one layer on the host (arithmetic_circuit.rs:82-97), oracle field arithmetic"""
    out = np.zeros((n_out, 4), dtype=np.uint64)
    for l, r, o, op in layer:
        out[o] = co.fe_op("add", fid, out[o], co.fe_op("add" if op == 0 else "mul", fid, values[l], values[r]))
    return out
