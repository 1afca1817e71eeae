"""The verifiers against the oracle's restatement of the reference verifiers: same accept / reject
decisions on honest and tampered proofs, same replayed challenges; and the reference's benches as they are written
(prove + verify)."""
import random

import numpy as np
import pytest

from conftest import FIELDS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fid", [0, 2])
def test_basic_verifier_decisions(zk, co, ctx_for, fid):
    from zk_cryptography_research_implementations_b200.sumcheck_protocol import Prover, Verifier, SumcheckProof
    ctx = ctx_for(fid)
    for n in (0, 1, 3, 8, 14):
        table = ctx.generate(3, n, 1 << n).download()
        proof = Prover.init(ctx, table).prove()
        assert Verifier.init(ctx).verify(proof) is True
        assert co.basic_verify(fid, table, proof.initial_claimed_sum, proof.round_univariate_polynomials)
        if n == 0:
            continue
        rng = random.Random(n)
        for what in ("round", "claim", "table", "length"):
            rp = proof.round_univariate_polynomials.copy()
            claimed = proof.initial_claimed_sum.copy()
            tab = table.copy()
            if what == "round":
                rp[rng.randrange(n), rng.randrange(2), 0] ^= np.uint64(1)
            elif what == "claim":
                claimed = zk.fe_binop("add", fid, claimed, zk.fe_from_int(fid, 1))
            elif what == "table":
                tab[rng.randrange(1 << n), 0] ^= np.uint64(1)
            else:
                rp = rp[:-1]
            bad = SumcheckProof(tab, claimed, rp)
            want = co.basic_verify(fid, tab, claimed, rp)
            assert Verifier.init(ctx).verify(bad) == want, (n, what)
            assert want is False


@pytest.mark.parametrize("fid", [0, 2])
def test_product_verifier_decisions_and_challenges(zk, co, ctx_for, fid):
    from zk_cryptography_research_implementations_b200 import sumcheck_protocol as scp
    from zk_cryptography_research_implementations_b200.polynomials import DenseUnivariatePolynomial, MultilinearPolynomial as MLE, ProductPolynomial, SumPolynomial
    from zk_cryptography_research_implementations_b200.transcripts import Transcript
    ctx = ctx_for(fid)
    for n in (1, 4, 9):
        tabs = np.stack([np.stack([ctx.generate(9, 2 * p + d, 1 << n).download() for d in range(2)]) for p in range(2)])
        claimed = np.zeros(4, dtype=np.uint64)
        co.lib().zko_fe_sum(fid, co._p(co.sumpoly_reduce(fid, tabs)), 1 << n, co._p(claimed))
        sp = SumPolynomial([ProductPolynomial([MLE.new(ctx, t) for t in prod]) for prod in tabs])
        result = scp.prove(sp, claimed, Transcript())
        verified = scp.verify(fid, result, Transcript())
        assert verified.is_proof_valid and np.array_equal(verified.random_challenges, result.random_challenges)
        coeffs = np.stack([p.coefficients for p in result.round_univariate_polynomials])
        ok, ch, last = co.product_verify(fid, claimed, coeffs, co.Transcript())
        assert ok and np.array_equal(ch, verified.random_challenges) and np.array_equal(last, verified.last_claimed_sum)
        # tamper with one coefficient: both verifiers reject at the same round with the same last claim
        k = n // 2
        bad_coeffs = coeffs.copy()
        bad_coeffs[k, 1, 0] ^= np.uint64(1)
        bad = scp.SumcheckProverProof(claimed, [DenseUnivariatePolynomial(fid, c) for c in bad_coeffs], result.random_challenges)
        v2 = scp.verify(fid, bad, Transcript())
        ok2, _, last2 = co.product_verify(fid, claimed, bad_coeffs, co.Transcript())
        assert v2.is_proof_valid == ok2 == False and np.array_equal(v2.last_claimed_sum, last2)
        assert v2.random_challenges.shape[0] == 0


def test_gkr_verifier_decisions(zk, co, ctx_for, golden):
    from zk_cryptography_research_implementations_b200 import gkr
    from zk_cryptography_research_implementations_b200.circuit import Circuit, Gate, Layer
    for e in golden["reference_kats"]["gkr_round_trips"] + [{"field": g["field"], "layers": g["layers"], "inputs": g["inputs"]} for g in golden["generated"]["gkr"]]:
        fid = FIELDS[e["field"]]
        ctx = ctx_for(fid)
        circuit = Circuit.new(fid, [Layer.new([Gate.new(*g) for g in l]) for l in e["layers"]])
        I = zk.fe_from_ints(fid, e["inputs"])
        proof = gkr.prove(ctx, circuit, I)                       # gkr_protocol.rs:246-299: prove, then verify == true
        assert gkr.verify(ctx, circuit, proof, I) is True
        # wrong inputs, a tampered coefficient, a tampered W evaluation, a wrong output
        I2 = I.copy(); I2[0] = zk.fe_binop("add", fid, I2[0], zk.fe_from_int(fid, 1))
        assert gkr.verify(ctx, circuit, proof, I2) is False
        proof.sumcheck_proofs[-1].round_univariate_polynomials[0].coefficients[2, 0] ^= np.uint64(1)
        assert gkr.verify(ctx, circuit, proof, I) is False
        proof.sumcheck_proofs[-1].round_univariate_polynomials[0].coefficients[2, 0] ^= np.uint64(1)
        if proof.wb_evaluations.shape[0]:
            proof.wb_evaluations[0, 0] ^= np.uint64(1)
            assert gkr.verify(ctx, circuit, proof, I) is False
            proof.wb_evaluations[0, 0] ^= np.uint64(1)
        proof.circuit_output[0] = zk.fe_binop("add", fid, proof.circuit_output[0], zk.fe_from_int(fid, 1))
        assert gkr.verify(ctx, circuit, proof, I) is False
