"""The reference's own criterion harnesses, shape for shape (prove with the GPU library, verify with the restated
reference verifier), plus BASELINE.json configs[0]: the basic-sumcheck bench on a seeded random 2^16-entry table,
bit-exact against the CPU oracle.

  sumcheck_protocol/benches/basic_sumcheck_benchmark.rs:5-28
  sumcheck_protocol/benches/gkr_sumcheck_benchmark.rs:13-40   (transcripts persist across iterations!)
  gkr/benches/gkr_protocol_benchmark.rs:6-24
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_basic_sumcheck_benchmark_shape_and_config0(zk, co, ctx_for):
    from zk_cryptography_research_implementations_b200.sumcheck_protocol import Prover
    fid = zk.BLS12_381_FR
    ctx = ctx_for(fid)
    Fr = lambda v: zk.fe_from_ints(fid, [v])[0]
    polynomial_evaluated_values = np.stack([Fr(0), Fr(0), Fr(2), Fr(7), Fr(3), Fr(3), Fr(6), Fr(11)])
    for _ in range(3):   # b.iter(|| { init; prove; init; verify })
        prover = Prover.init(ctx, polynomial_evaluated_values)
        proof = prover.prove()
        assert co.basic_verify(fid, proof.initial_polynomial, proof.initial_claimed_sum, proof.round_univariate_polynomials)
    # configs[0]: "prove+verify the sum of a random 2^16-entry multilinear polynomial on CPU (bit-exact ref)"
    table = ctx.generate(0xB200, 0, 1 << 16).download()
    proof = Prover.init(ctx, table).prove()
    claimed, rp, ch, fin = co.basic_prove(fid, table)
    assert np.array_equal(proof.initial_claimed_sum, claimed)
    assert np.array_equal(proof.round_univariate_polynomials, rp)
    assert np.array_equal(proof.challenges, ch) and np.array_equal(proof.final_evaluation, fin)
    assert co.basic_verify(fid, table, proof.initial_claimed_sum, proof.round_univariate_polynomials)
    # a tampered round polynomial is rejected
    bad = proof.round_univariate_polynomials.copy()
    bad[5, 1, 0] ^= np.uint64(1)
    assert not co.basic_verify(fid, table, proof.initial_claimed_sum, bad)


def test_gkr_sumcheck_benchmark_shape(zk, co, ctx_for):
    from zk_cryptography_research_implementations_b200 import sumcheck_protocol as scp
    from zk_cryptography_research_implementations_b200.polynomials import MultilinearPolynomial as MLE, ProductPolynomial, SumPolynomial
    from zk_cryptography_research_implementations_b200.transcripts import Transcript
    fid = zk.BN254_FQ
    ctx = ctx_for(fid)
    Fq = lambda v: zk.fe_from_ints(fid, [v])[0]

    def sum_polynomial():
        poly1a = MLE.new(ctx, np.stack([Fq(0), Fq(0), Fq(0), Fq(2)]))
        poly2a = MLE.new(ctx, np.stack([Fq(0), Fq(0), Fq(0), Fq(3)]))
        poly1b = MLE.new(ctx, np.stack([Fq(0), Fq(0), Fq(0), Fq(2)]))
        poly2b = MLE.new(ctx, np.stack([Fq(0), Fq(0), Fq(0), Fq(3)]))
        return SumPolynomial([ProductPolynomial([poly1a, poly2a]), ProductPolynomial([poly1b, poly2b])])

    prover_transcript = Transcript()          # created once, outside b.iter: they keep absorbing
    verifier_transcript = co.Transcript()
    oracle_prover_transcript = co.Transcript()
    tabs = np.stack([np.stack([zk.fe_from_ints(fid, [0, 0, 0, 2]), zk.fe_from_ints(fid, [0, 0, 0, 3])])] * 2)
    for _ in range(4):
        result = scp.prove(sum_polynomial(), Fq(12), prover_transcript)     # sum_polynomial.clone()
        coeffs = np.stack([p.coefficients for p in result.round_univariate_polynomials])
        ok, chal, _ = co.product_verify(fid, Fq(12), coeffs, verifier_transcript)
        assert ok and np.array_equal(chal, result.random_challenges)
        want_coeffs, want_chal, _ = co.product_prove(fid, tabs, Fq(12), oracle_prover_transcript)
        assert np.array_equal(coeffs, want_coeffs) and np.array_equal(result.random_challenges, want_chal)


def test_gkr_protocol_benchmark_shape(zk, co, ctx_for):
    from zk_cryptography_research_implementations_b200 import gkr
    from zk_cryptography_research_implementations_b200.circuit import Circuit, Gate, Layer, Operator
    fid = zk.BLS12_381_FR
    ctx = ctx_for(fid)
    gate1 = Gate.new(0, 1, 0, Operator.Mul)
    gate2 = Gate.new(0, 1, 0, Operator.Add)
    gate3 = Gate.new(2, 3, 1, Operator.Mul)
    layer0 = Layer.new([gate1])
    layer1 = Layer.new([gate2, gate3])
    circuit = Circuit.new(fid, [layer0, layer1])
    inputs = zk.fe_from_ints(fid, [2, 3, 4, 5])
    oc = co.Circuit([[(0, 1, 0, 1)], [(0, 1, 0, 0), (2, 3, 1, 1)]])
    for _ in range(2):
        proof = gkr.prove(ctx, circuit, inputs)
        want = co.gkr_prove(fid, oc, inputs)
        got_coeffs = np.concatenate([np.stack([p.coefficients for p in sp.round_univariate_polynomials]) for sp in proof.sumcheck_proofs])
        assert np.array_equal(got_coeffs, want.coeffs) and np.array_equal(proof.claimed_sum, want.claimed_sum)
        assert zk.fe_to_ints(fid, proof.circuit_output) == [100]
        assert co.gkr_verify(fid, oc, want, inputs)
        wide = gkr.prove_wide(ctx, gkr.WideCircuit.reference_shaped(ctx, [[(0, 1, 0, 1)], [(0, 1, 0, 0), (2, 3, 1, 1)]]), inputs)
        assert np.array_equal(np.concatenate([np.stack([p.coefficients for p in sp.round_univariate_polynomials]) for sp in wide.sumcheck_proofs]), want.coeffs)
