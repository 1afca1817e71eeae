"""Succinct GKR (gkr/src/succinct_gkr_protocol.rs) on the GPU: the reference's own circuits (:293-405), the proof compared
part by part with the oracle -- the GKR transcript with oracle/zkoracle.c's restatement of gkr_protocol::prove (the
transcript flow of prove_succinct is the same: the commitment is not absorbed), commitment and openings with
oracle/zkoracle_kzg.c -- and verify_succinct's decision on honest and tampered proofs."""
import copy
import random

import numpy as np
import pytest

import pykzg as pk

pytestmark = pytest.mark.gpu
FR = 2
R = pk.R
ADD, MUL = 0, 1

REFERENCE = [
    # test_succinct_gkr_protocol1 (:293-318): (left, right, out, op) per layer, inputs, taus
    ([[(0, 1, 0, MUL)], [(0, 1, 0, ADD), (2, 3, 1, MUL)]], [2, 3, 4, 5], [5, 2]),
    # test_succinct_gkr_protocol2 (:320-361)
    ([[(0, 1, 0, ADD)], [(0, 1, 0, MUL), (2, 3, 1, ADD)], [(0, 1, 0, ADD), (2, 3, 1, ADD), (4, 5, 2, ADD), (6, 7, 3, ADD)]],
     [1, 2, 3, 4, 5, 6, 7, 8], [5, 2, 3]),
]


def _coeffs(proof):
    return np.concatenate([np.stack([p.coefficients for p in sp.round_univariate_polynomials]) for sp in proof.sumcheck_proofs])


def _check_against_oracle(co, proof, want, setup_g1, inputs, L):
    got = _coeffs(proof)
    assert np.array_equal(got, want.coeffs[: got.shape[0]])
    assert np.array_equal(np.concatenate([sp.random_challenges for sp in proof.sumcheck_proofs]), want.challenges[: got.shape[0]])
    assert np.array_equal(np.stack([sp.claimed_sum for sp in proof.sumcheck_proofs]), want.layer_claims)
    assert np.array_equal(proof.wb_evaluations, want.wb[: L - 1]) and np.array_equal(proof.wc_evaluations, want.wc[: L - 1])
    assert np.array_equal(proof.claimed_sum, want.claimed_sum)
    assert (proof.input_polynomial_commitment == co.kzg_commit(inputs, setup_g1)).all()
    chal = proof.sumcheck_proofs[-1].random_challenges
    mid = chal.shape[0] // 2
    for got, point in ((proof.input_rb_proof, chal[:mid]), (proof.input_rc_proof, chal[mid:])):
        ev, prs = co.kzg_open(inputs, setup_g1, point)
        assert (got.evaluation == ev).all() and (got.proofs == prs).all()


@pytest.mark.parametrize("case", range(len(REFERENCE)))
def test_reference_succinct_circuits(zk, co, ctx_for, case):
    from zk_cryptography_research_implementations_b200 import gkr
    from zk_cryptography_research_implementations_b200.multilinear_kzg import TrustedSetup
    layers, inputs, taus = REFERENCE[case]
    ctx = ctx_for(FR)
    I = zk.fe_from_ints(FR, inputs)
    setup = TrustedSetup.initialize_setup(ctx, zk.fe_from_ints(FR, taus))
    wc = gkr.WideCircuit.reference_shaped(ctx, layers)
    proof = gkr.prove_succinct(ctx, wc, I, setup)
    want = co.gkr_prove(FR, co.Circuit(layers), I)
    _check_against_oracle(co, proof, want, setup.g1_powers_of_tau, I, len(layers))
    assert gkr.verify_succinct(ctx, wc, proof, setup)                                   # the reference's assertion
    assert gkr.verify_succinct(ctx, wc, proof, setup, bind_input_openings=True)
    # the independent pairing model accepts the two openings as well
    psetup = pk.TrustedSetup.initialize(taus)
    chal = proof.sumcheck_proofs[-1].random_challenges
    mid = chal.shape[0] // 2
    for pr, point in ((proof.input_rb_proof, chal[:mid]), (proof.input_rc_proof, chal[mid:])):
        assert pk.verify(psetup, co.g1_to_ints(proof.input_polynomial_commitment)[0], co.to_ints(FR, point),
                         co.to_ints(FR, pr.evaluation)[0], co.g1_to_ints(pr.proofs))


def test_random_taus_and_tampering(zk, co, ctx_for):
    """test_succinct_gkr_protocol_with_random_values_of_tau (:363-405) and what verify_succinct must reject"""
    from zk_cryptography_research_implementations_b200 import gkr
    from zk_cryptography_research_implementations_b200.multilinear_kzg import TrustedSetup
    layers, inputs, _ = REFERENCE[1]
    rnd = random.Random(2024)
    taus = [rnd.randrange(R) for _ in range(3)]
    ctx = ctx_for(FR)
    I = zk.fe_from_ints(FR, inputs)
    setup = TrustedSetup.initialize_setup(ctx, zk.fe_from_ints(FR, taus))
    wc = gkr.WideCircuit.reference_shaped(ctx, layers)
    proof = gkr.prove_succinct(ctx, wc, I, setup)
    assert gkr.verify_succinct(ctx, wc, proof, setup)
    one = zk.fe_from_int(FR, 1)
    t = copy.deepcopy(proof); t.circuit_output[0] = zk.fe_binop("add", FR, t.circuit_output[0], one); assert not gkr.verify_succinct(ctx, wc, t, setup)
    t = copy.deepcopy(proof); t.wb_evaluations[0] = zk.fe_binop("add", FR, t.wb_evaluations[0], one); assert not gkr.verify_succinct(ctx, wc, t, setup)
    t = copy.deepcopy(proof); t.input_rb_proof.evaluation = zk.fe_binop("add", FR, t.input_rb_proof.evaluation, one); assert not gkr.verify_succinct(ctx, wc, t, setup)
    t = copy.deepcopy(proof); t.input_rc_proof.proofs[1] = co.g1_add(t.input_rc_proof.proofs[1], co.g1_generator()); assert not gkr.verify_succinct(ctx, wc, t, setup)
    t = copy.deepcopy(proof); t.input_polynomial_commitment = co.g1_add(t.input_polynomial_commitment, co.g1_generator()); assert not gkr.verify_succinct(ctx, wc, t, setup)
    # a commitment to OTHER inputs with honest openings of those inputs: the reference's verifier accepts it (it never ties
    # the opened values to the last sumcheck claim); the bound form rejects it
    from zk_cryptography_research_implementations_b200.multilinear_kzg import MultilinearKZG
    other = zk.fe_from_ints(FR, [9, 9, 9, 9, 1, 2, 3, 4])
    chal = proof.sumcheck_proofs[-1].random_challenges
    t = copy.deepcopy(proof)
    t.input_polynomial_commitment = MultilinearKZG.commit_to_polynomial(other, setup)
    t.input_rb_proof = MultilinearKZG.open_and_prove(other, setup, chal[:3])
    t.input_rc_proof = MultilinearKZG.open_and_prove(other, setup, chal[3:])
    assert gkr.verify_succinct(ctx, wc, t, setup)
    assert not gkr.verify_succinct(ctx, wc, t, setup, bind_input_openings=True)


def test_wide_succinct_proof_against_the_gate_list_oracle(zk, co, ctx_for):
    """a 2^10-wide, 3-layer circuit over BLS12-381 Fr: GKR part against zko_gkr_prove_sparse, KZG part against zkoracle_kzg"""
    from zk_cryptography_research_implementations_b200 import gkr
    from zk_cryptography_research_implementations_b200.multilinear_kzg import TrustedSetup
    from test_gpu_gkr import _wide_arrays
    ctx = ctx_for(FR)
    w, depth = 10, 3
    rng = np.random.default_rng(4242)
    bits, layers = _wide_arrays(rng, w, depth)
    dev_inputs = ctx.generate(21, 0, 1 << w)
    inputs = dev_inputs.download()
    rnd = random.Random(8)
    taus = zk.fe_from_ints(FR, [rnd.randrange(R) for _ in range(w)])
    setup = TrustedSetup.initialize_setup(ctx, taus)
    wc = gkr.WideCircuit(ctx, bits, layers)
    proof = gkr.prove_succinct(ctx, wc, dev_inputs, setup)
    want = co.gkr_prove_sparse(FR, co.SparseCircuit(bits, layers), inputs)
    co.set_threads(8)
    try:
        _check_against_oracle(co, proof, want, setup.g1_powers_of_tau, inputs, depth)
    finally:
        co.set_threads(1)
    assert gkr.verify_succinct(ctx, wc, proof, setup)
    assert gkr.verify_succinct(ctx, wc, proof, setup, bind_input_openings=True)
    wc.close()
