"""`gkr` crate mirror: gkr/src/gkr_protocol.rs `prove` + `Proof` on the GPU (verifier: parity harness only)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List

import numpy as np

from .circuit import Circuit
from .core import Context, _ptr, as_elems
from .polynomials import DenseUnivariatePolynomial
from .sumcheck_protocol import SumcheckProverProof


@dataclass
class Proof:                                   # gkr_protocol.rs:17-23
    circuit_output: np.ndarray
    claimed_sum: np.ndarray
    sumcheck_proofs: List[SumcheckProverProof]
    wb_evaluations: np.ndarray
    wc_evaluations: np.ndarray


def prove(ctx: Context, circuit: Circuit, inputs) -> Proof:      # gkr_protocol.rs:26-143
    lib = ctx.lib
    inputs = as_elems(inputs).reshape(-1, 4)
    L = len(circuit.layers)
    R = int(lib.zk_gkr_total_rounds(L))
    out_cap = max((g.output_index for g in circuit.layers[0].gates), default=0) + 1
    output = np.zeros((out_cap, 4), dtype=np.uint64)
    n_out = C.c_uint64()
    claimed = np.zeros(4, dtype=np.uint64)
    claims = np.zeros((L, 4), dtype=np.uint64)
    coeffs = np.zeros((R, 3, 4), dtype=np.uint64)
    chal = np.zeros((R, 4), dtype=np.uint64)
    wb = np.zeros((max(L - 1, 1), 4), dtype=np.uint64)
    wc = np.zeros((max(L - 1, 1), 4), dtype=np.uint64)
    ctx.check(lib.zk_gkr_prove(ctx.h, C.byref(circuit.desc), _ptr(inputs), inputs.shape[0], _ptr(output), out_cap,
                               C.cast(C.byref(n_out), C.POINTER(C.c_uint64)), _ptr(claimed), _ptr(claims), _ptr(coeffs),
                               _ptr(chal), _ptr(wb), _ptr(wc)))
    proofs, o = [], 0
    for i in range(L):
        r = 2 * (i + 1)
        polys = [DenseUnivariatePolynomial(ctx.field, coeffs[o + k]) for k in range(r)]
        proofs.append(SumcheckProverProof(claims[i].copy(), polys, chal[o:o + r].copy()))
        o += r
    return Proof(output[: n_out.value].copy(), claimed, proofs, wb[: L - 1].copy(), wc[: L - 1].copy())
