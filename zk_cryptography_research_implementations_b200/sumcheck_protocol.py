"""`sumcheck_protocol` crate mirror: the two provers, running on the GPU.

  basic_sumcheck.prover      sumcheck_protocol/src/basic_sumcheck/prover.rs
  basic_sumcheck.verifier    sumcheck_protocol/src/basic_sumcheck/verifier.rs
  gkr_sumcheck               sumcheck_protocol/src/gkr_sumcheck/sumcheck_gkr_protocol.rs (prove and verify)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _lib
from .core import Context, DeviceTable, ReferencePanic, _ptr, as_elems
from .polynomials import DenseUnivariatePolynomial, MultilinearPolynomial, SumPolynomial
from .transcripts import Transcript


# =============================================================== basic_sumcheck
@dataclass
class SumcheckProof:
    """`SumcheckProof<F>` (prover.rs:15-19); round polynomials are 2-entry evaluation tables."""
    initial_polynomial: np.ndarray            # (N, 4)
    initial_claimed_sum: np.ndarray           # (4,)
    round_univariate_polynomials: np.ndarray  # (n, 2, 4)
    # extras (not part of the reference proof; handy for parity checks)
    challenges: Optional[np.ndarray] = None
    final_evaluation: Optional[np.ndarray] = None


class Prover:
    """`Prover<F>` (prover.rs:7-13)."""

    def __init__(self):
        self.is_initialized = False

    @classmethod
    def init(cls, ctx: Context, polynomial_evaluated_values) -> "Prover":   # prover.rs:22-33
        self = cls()
        ev = as_elems(polynomial_evaluated_values).reshape(-1, 4)
        if ev.shape[0] == 0 or ev.shape[0] & (ev.shape[0] - 1):
            raise ReferencePanic("Evaluated values must be a power of 2")
        self.ctx = ctx
        self.initial_polynomial = ev.copy()
        self.is_initialized = True
        return self

    def prove(self) -> SumcheckProof:                                        # prover.rs:35-71
        if not self.is_initialized:
            raise ReferencePanic("Can't prove without init")
        ctx = self.ctx
        N = self.initial_polynomial.shape[0]
        n = N.bit_length() - 1
        claimed = np.zeros(4, dtype=np.uint64)
        rounds = np.zeros((max(n, 1), 2, 4), dtype=np.uint64)
        chal = np.zeros((max(n, 1), 4), dtype=np.uint64)
        fin = np.zeros(4, dtype=np.uint64)
        ctx.check(ctx.lib.zk_prove_basic(ctx.h, _ptr(self.initial_polynomial), N, _ptr(claimed), _ptr(rounds),
                                         _ptr(chal), _ptr(fin), 0))
        self.initial_claimed_sum = claimed
        return SumcheckProof(self.initial_polynomial.copy(), claimed, rounds[:n], chal[:n], fin)


class Verifier:
    """`Verifier<F>` (verifier.rs:8-12)."""

    def __init__(self):
        self.is_initialized = False

    @classmethod
    def init(cls, ctx: Context) -> "Verifier":                              # verifier.rs:15-21
        self = cls()
        self.ctx = ctx
        self.is_initialized = True
        return self

    def verify(self, proof: SumcheckProof) -> bool:                         # verifier.rs:23-71
        if not self.is_initialized:
            raise ReferencePanic("Can't verify without init")
        ctx = self.ctx
        table = ctx.upload(as_elems(proof.initial_polynomial).reshape(-1, 4))
        rp = np.ascontiguousarray(as_elems(proof.round_univariate_polynomials).reshape(-1, 2, 4))
        ok = C.c_int(0)
        ctx.check(ctx.lib.zk_verify_basic(ctx.h, table.h, _ptr(as_elems(proof.initial_claimed_sum)),
                                          _ptr(rp) if rp.size else None, rp.shape[0], C.byref(ok)))
        return bool(ok.value)


def split_polynomial_and_sum_each(ctx: Context, polynomial_evaluated_values) -> np.ndarray:   # prover.rs:74-89
    t = ctx.upload(as_elems(polynomial_evaluated_values).reshape(-1, 4))
    out = np.zeros((2, 4), dtype=np.uint64)
    ctx.check(ctx.lib.zk_sum_halves(ctx.h, t.h, _ptr(out)))
    return out


# =============================================================== gkr_sumcheck
@dataclass
class SumcheckProverProof:
    """`SumcheckProverProof<F>` (sumcheck_gkr_protocol.rs:9-13)."""
    claimed_sum: np.ndarray
    round_univariate_polynomials: List[DenseUnivariatePolynomial]
    random_challenges: np.ndarray              # (n, 4)
    final_values: Optional[np.ndarray] = None  # extra: tables after the last fold, (P*D, 4)


def generate_round_univariate(current_polynomial: SumPolynomial) -> np.ndarray:   # sumcheck_gkr_protocol.rs:113-143
    ctx = current_polynomial.ctx
    sp = current_polynomial._device_sumpoly(clone=True)
    try:
        out = np.zeros((current_polynomial.degree() + 1, 4), dtype=np.uint64)
        ctx.check(ctx.lib.zk_sumcheck_round_evals(ctx.h, sp, _ptr(out)))
        return out
    finally:
        ctx.lib.zk_sumpoly_free(ctx.h, sp)


def prove(sum_polynomial: SumPolynomial, claimed_sum, transcript: Transcript, flags: int = 0, consume: bool = False) -> SumcheckProverProof:
    """`prove` (sumcheck_gkr_protocol.rs:24-67).  The transcript is borrowed and advanced.  By default the tables are
    copied on the device and `sum_polynomial` stays usable (a Rust caller that wants to keep its polynomial clones it
    before the by-value call).  consume=True is the reference's by-value move without the copy: the tables are folded
    in place and every polynomial of `sum_polynomial` is dead afterwards (use raises)."""
    ctx = sum_polynomial.ctx
    P, D = len(sum_polynomial.product_polynomials), sum_polynomial.degree()
    if P < 2:
        raise ReferencePanic("more than one product polynomial required for add operation")
    if D < 2:
        raise ReferencePanic("more than one polynomial required for mul operation")
    n = sum_polynomial.number_of_variables()
    sp = sum_polynomial._device_sumpoly(clone=not consume)
    try:
        coeffs = np.zeros((max(n, 1), D + 1, 4), dtype=np.uint64)
        chal = np.zeros((max(n, 1), 4), dtype=np.uint64)
        fin = np.zeros((P * D, 4), dtype=np.uint64)
        claimed = as_elems(claimed_sum).copy()
        ctx.check(ctx.lib.zk_prove_product(ctx.h, sp, _ptr(claimed), transcript.h, _ptr(coeffs), _ptr(chal), _ptr(fin),
                                           flags))
    finally:
        ctx.lib.zk_sumpoly_free(ctx.h, sp)
    polys = [DenseUnivariatePolynomial(ctx.field, coeffs[k]) for k in range(n)]
    return SumcheckProverProof(claimed, polys, chal[:n], fin)


@dataclass
class SumcheckVerifierProof:
    """`SumcheckVerifierProof<F>` (sumcheck_gkr_protocol.rs:15-20)."""
    is_proof_valid: bool
    random_challenges: np.ndarray
    last_claimed_sum: np.ndarray


def verify(field: int, proof: SumcheckProverProof, transcript: Transcript) -> SumcheckVerifierProof:
    """`verify` (sumcheck_gkr_protocol.rs:69-106)"""
    lib = _lib.load()
    polys = proof.round_univariate_polynomials
    n = len(polys)
    D = (polys[0].coefficients.shape[0] - 1) if n else 2
    coeffs = np.ascontiguousarray(np.stack([p.coefficients for p in polys])) if n else np.zeros((1, D + 1, 4), dtype=np.uint64)
    chal = np.zeros((max(n, 1), 4), dtype=np.uint64)
    last = np.zeros(4, dtype=np.uint64)
    ok = C.c_int(0)
    rc = lib.zk_verify_product(field, _ptr(as_elems(proof.claimed_sum)), _ptr(coeffs), n, D, transcript.h, _ptr(chal), _ptr(last),
                               C.byref(ok))
    if rc:
        raise ValueError("zk_verify_product failed (%d)" % rc)
    if not ok.value:
        return SumcheckVerifierProof(False, np.zeros((0, 4), dtype=np.uint64), last)     # :85-89 `random_challenges: vec![]`
    return SumcheckVerifierProof(True, chal[:n], last)


def univariate_to_bytes(field: int, univariate_poly) -> bytes:      # sumcheck_gkr_protocol.rs:145-150 (little-endian)
    from .core import fe_to_ints
    return b"".join(v.to_bytes(32, "little") for v in fe_to_ints(field, univariate_poly))


def field_element_to_bytes(field: int, field_element) -> bytes:     # sumcheck_gkr_protocol.rs:152-154 / prover.rs:91-93
    from .core import fe_to_ints
    return fe_to_ints(field, field_element)[0].to_bytes(32, "big")
