"""`gkr` crate mirror: gkr/src/gkr_protocol.rs `prove` / `verify` + `Proof`, and gkr/src/succinct_gkr_protocol.rs
`prove_succinct` / `verify_succinct` + `SuccinctProof` (the input layer behind a multilinear KZG commitment)."""
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


def verify(ctx: Context, circuit: Circuit, proof: Proof, inputs) -> bool:      # gkr_protocol.rs:146-236
    lib = ctx.lib
    inputs = as_elems(inputs).reshape(-1, 4)
    L = len(circuit.layers)
    out = np.ascontiguousarray(as_elems(proof.circuit_output).reshape(-1, 4))
    claims = np.ascontiguousarray(np.stack([sp.claimed_sum for sp in proof.sumcheck_proofs]))
    coeffs = np.ascontiguousarray(np.concatenate([np.stack([p.coefficients for p in sp.round_univariate_polynomials])
                                                  for sp in proof.sumcheck_proofs]))
    wb = np.ascontiguousarray(np.concatenate([as_elems(proof.wb_evaluations).reshape(-1, 4), np.zeros((1, 4), dtype=np.uint64)]))
    wc = np.ascontiguousarray(np.concatenate([as_elems(proof.wc_evaluations).reshape(-1, 4), np.zeros((1, 4), dtype=np.uint64)]))
    ok = C.c_int(0)
    ctx.check(lib.zk_gkr_verify(ctx.h, C.byref(circuit.desc), _ptr(out), out.shape[0], _ptr(claims), _ptr(coeffs), _ptr(wb), _ptr(wc),
                                _ptr(inputs), inputs.shape[0], C.byref(ok)))
    return bool(ok.value)


class WideCircuit:
    """A layered circuit with explicit layer widths, resident on the GPU (three CSR orderings of every layer's
    gates) for the sparse two-phase layer prover (csrc/gkr_wide.cu).  layer_bits[li] = log2(#values of layer li),
    li = 0..L (last entry = inputs).  `layers` as for `Circuit`: lists of (left, right, out, op)."""

    def __init__(self, ctx: Context, layer_bits, layers=None, *, flat=None):
        """`layers`: per layer a list of (left, right, out, op) or an (n, 4) array.  `flat=(layer_off, left, right, out, op)`:
        the C-ABI's own layout (uint64 offsets, uint32 indices, uint8 operators) handed through without a copy."""
        self.ctx = ctx
        self.layer_bits = [int(b) for b in layer_bits]
        if flat is not None:
            self._off, left, right, out, op = flat
            self._off = np.ascontiguousarray(self._off, dtype=np.uint64)
            left, right, out = (np.ascontiguousarray(a, dtype=np.uint32) for a in (left, right, out))
            op = np.ascontiguousarray(op, dtype=np.uint8)
            self.n_layers = len(self._off) - 1
        else:
            self.n_layers = len(layers)
            off = [0]
            for l in layers:
                off.append(off[-1] + len(l))
            def col(k, dt):
                if isinstance(layers[0], np.ndarray):
                    return np.ascontiguousarray(np.concatenate([l[:, k] for l in layers]).astype(dt))
                return np.array([g[k] for l in layers for g in l], dtype=dt)
            self._off = np.array(off, dtype=np.uint64)
            left, right, out, op = col(0, np.uint32), col(1, np.uint32), col(2, np.uint32), col(3, np.uint8)
        assert len(self.layer_bits) == self.n_layers + 1
        bits = np.array(self.layer_bits, dtype=np.uint32)
        h = C.c_void_p()
        u32p = C.POINTER(C.c_uint32)
        ctx.check(ctx.lib.zk_wide_circuit_create(ctx.h, self.n_layers, bits.ctypes.data_as(u32p), _ptr(self._off),
                                                 left.ctypes.data_as(u32p), right.ctypes.data_as(u32p), out.ctypes.data_as(u32p),
                                                 op.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(h)))
        self.h = h
        # a single output is the reference's padded [out, 0] layer (gkr_protocol.rs:43-51): the prover sees one output bit
        self.layer_bits[0] = int(ctx.lib.zk_wide_circuit_output_bits(h))

    @classmethod
    def reference_shaped(cls, ctx: Context, layers):
        """the reference's rigid shape: layer 0 has one output bit, layer i >= 1 has i, inputs have L bits"""
        L = len(layers)
        return cls(ctx, [1] + list(range(1, L + 1)), layers)

    def total_rounds(self) -> int:
        return int(self.ctx.lib.zk_wide_circuit_total_rounds(self.h))

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.zk_wide_circuit_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def prove_wide(ctx: Context, circuit: WideCircuit, inputs, flags: int = 0, want_output: bool = True, sharded: bool = False,
               collapse_len: int = 1 << 12) -> Proof:
    """gkr_protocol::prove with the sparse two-phase layer sumcheck (same proof as `prove` on reference shapes).
    sharded=True (after sharded.init_comm, every rank calls with the same circuit and the whole input layer as a
    DeviceTable): the layer sumchecks are spread over the ranks; every rank returns the same proof."""
    lib = ctx.lib
    from .core import DeviceTable
    resident = isinstance(inputs, DeviceTable)      # input layer already in HBM: no host-to-device copy in the call
    if not resident:
        inputs = as_elems(inputs).reshape(-1, 4)
    L = circuit.n_layers
    R = circuit.total_rounds()
    n_out = 1 << circuit.layer_bits[0]
    output = np.zeros((n_out, 4), dtype=np.uint64) if want_output else None
    claimed = np.zeros(4, dtype=np.uint64)
    claims = np.zeros((L, 4), dtype=np.uint64)
    coeffs = np.zeros((R, 3, 4), dtype=np.uint64)
    chal = np.zeros((R, 4), dtype=np.uint64)
    wb = np.zeros((max(L - 1, 1), 4), dtype=np.uint64)
    wc = np.zeros((max(L - 1, 1), 4), dtype=np.uint64)
    if sharded:
        if not resident:
            raise ValueError("sharded GKR takes the input layer as a DeviceTable (replicated on every rank)")
        ctx.check(lib.zk_gkr_prove_wide_sharded(ctx.h, circuit.h, inputs.h, _ptr(output) if want_output else None, _ptr(claimed),
                                                _ptr(claims), _ptr(coeffs), _ptr(chal), _ptr(wb), _ptr(wc), flags, collapse_len))
    elif resident:
        ctx.check(lib.zk_gkr_prove_wide_device(ctx.h, circuit.h, inputs.h, _ptr(output) if want_output else None,
                                               _ptr(claimed), _ptr(claims), _ptr(coeffs), _ptr(chal), _ptr(wb), _ptr(wc), flags))
    else:
        ctx.check(lib.zk_gkr_prove_wide(ctx.h, circuit.h, _ptr(inputs), inputs.shape[0], _ptr(output) if want_output else None,
                                        _ptr(claimed), _ptr(claims), _ptr(coeffs), _ptr(chal), _ptr(wb), _ptr(wc), flags))
    proofs, o = [], 0
    for i in range(L):
        r = 2 * circuit.layer_bits[i + 1]
        polys = [DenseUnivariatePolynomial(ctx.field, coeffs[o + k]) for k in range(r)]
        proofs.append(SumcheckProverProof(claims[i].copy(), polys, chal[o:o + r].copy()))
        o += r
    return Proof(output, claimed, proofs, wb[: L - 1].copy(), wc[: L - 1].copy())


def verify_wide(ctx: Context, circuit: WideCircuit, proof: Proof, inputs, flags: int = 0) -> bool:
    """gkr_protocol::verify (gkr_protocol.rs:146-236) for explicit layer widths; wiring predicates from the gate list"""
    lib = ctx.lib
    from .core import DeviceTable
    L = circuit.n_layers
    out = np.ascontiguousarray(as_elems(proof.circuit_output).reshape(-1, 4))
    if out.shape[0] != (1 << circuit.layer_bits[0]):
        return False
    claims = np.ascontiguousarray(np.stack([sp.claimed_sum for sp in proof.sumcheck_proofs]))
    coeffs = np.ascontiguousarray(np.concatenate([np.stack([p.coefficients for p in sp.round_univariate_polynomials])
                                                  for sp in proof.sumcheck_proofs]))
    pad = np.zeros((1, 4), dtype=np.uint64)
    wb = np.ascontiguousarray(np.concatenate([as_elems(proof.wb_evaluations).reshape(-1, 4), pad]))
    wc = np.ascontiguousarray(np.concatenate([as_elems(proof.wc_evaluations).reshape(-1, 4), pad]))
    ok = C.c_int(0)
    if isinstance(inputs, DeviceTable):
        ctx.check(lib.zk_gkr_verify_wide_device(ctx.h, circuit.h, _ptr(out), _ptr(claims), _ptr(coeffs), _ptr(wb), _ptr(wc),
                                                inputs.h, flags, C.byref(ok)))
    else:
        inputs = as_elems(inputs).reshape(-1, 4)
        ctx.check(lib.zk_gkr_verify_wide(ctx.h, circuit.h, _ptr(out), _ptr(claims), _ptr(coeffs), _ptr(wb), _ptr(wc),
                                         _ptr(inputs), inputs.shape[0], flags, C.byref(ok)))
    return bool(ok.value)


# ------------------------------------------------------------------ succinct GKR (gkr/src/succinct_gkr_protocol.rs)
@dataclass
class SuccinctProof:                           # succinct_gkr_protocol.rs:22-32
    circuit_output: np.ndarray
    claimed_sum: np.ndarray
    sumcheck_proofs: List[SumcheckProverProof]
    wb_evaluations: np.ndarray
    wc_evaluations: np.ndarray
    input_polynomial_commitment: np.ndarray    # (12,) affine G1
    input_rb_proof: "MultilinearKZGProof"
    input_rc_proof: "MultilinearKZGProof"


def prove_succinct(ctx: Context, circuit: WideCircuit, inputs, trusted_setup, flags: int = 0, sharded: bool = False,
                   collapse_len: int = 1 << 12) -> SuccinctProof:
    """succinct_gkr_protocol.rs:35-169.  The transcript flow is gkr_protocol::prove's (the commitment is not absorbed, :46-66);
    on top of it the input polynomial is committed (:43-44) and opened at rb and rc, the two halves of the input layer's
    sumcheck challenges (:118-123, :151-154).  `ctx` must be a BLS12-381 Fr context; `inputs` host limbs or a DeviceTable.
    sharded=True (after sharded.init_comm; circuit, setup and inputs replicated on every rank): the layer sumchecks and the
    multi-scalar multiplications are spread over the ranks; every rank returns the same proof."""
    from .core import DeviceTable
    from .multilinear_kzg import MultilinearKZG
    table = inputs if isinstance(inputs, DeviceTable) else ctx.upload(as_elems(inputs).reshape(-1, 4))
    commitment = MultilinearKZG.commit_to_polynomial(table, trusted_setup, sharded=sharded)
    base = prove_wide(ctx, circuit, table, flags, sharded=sharded, collapse_len=collapse_len)
    chal = base.sumcheck_proofs[-1].random_challenges
    mid = chal.shape[0] // 2
    rb_proof = MultilinearKZG.open_and_prove(table, trusted_setup, chal[:mid], sharded=sharded)
    rc_proof = MultilinearKZG.open_and_prove(table, trusted_setup, chal[mid:], sharded=sharded)
    return SuccinctProof(base.circuit_output, base.claimed_sum, base.sumcheck_proofs, base.wb_evaluations, base.wc_evaluations,
                         commitment, rb_proof, rc_proof)


def verify_succinct(ctx: Context, circuit: WideCircuit, proof: SuccinctProof, trusted_setup, flags: int = 0,
                    bind_input_openings: bool = False) -> bool:
    """succinct_gkr_protocol.rs:172-283: the layer sumchecks and claim checks on the GPU (wiring predicates from the gate
    list), the two input openings by the pairing check on the host.  bind_input_openings=True additionally requires the
    opened values W(rb), W(rc) to satisfy the input layer's last sumcheck claim -- a check the reference does not make."""
    from .multilinear_kzg import MultilinearKZG
    lib = ctx.lib
    L = circuit.n_layers
    out = np.ascontiguousarray(as_elems(proof.circuit_output).reshape(-1, 4))
    if out.shape[0] != (1 << circuit.layer_bits[0]):
        return False
    claims = np.ascontiguousarray(np.stack([sp.claimed_sum for sp in proof.sumcheck_proofs]))
    coeffs = np.ascontiguousarray(np.concatenate([np.stack([p.coefficients for p in sp.round_univariate_polynomials])
                                                  for sp in proof.sumcheck_proofs]))
    pad = np.zeros((1, 4), dtype=np.uint64)
    wb = np.ascontiguousarray(np.concatenate([as_elems(proof.wb_evaluations).reshape(-1, 4), pad]))
    wc = np.ascontiguousarray(np.concatenate([as_elems(proof.wc_evaluations).reshape(-1, 4), pad]))
    m = circuit.layer_bits[L]
    last = np.zeros((2 * m, 4), dtype=np.uint64)
    evals = None
    if bind_input_openings:
        evals = np.ascontiguousarray(np.stack([as_elems(proof.input_rb_proof.evaluation).reshape(4),
                                               as_elems(proof.input_rc_proof.evaluation).reshape(4)]))
    ok = C.c_int(0)
    ctx.check(lib.zk_gkr_verify_wide_succinct(ctx.h, circuit.h, _ptr(out), _ptr(claims), _ptr(coeffs), _ptr(wb), _ptr(wc),
                                              _ptr(evals) if evals is not None else None, flags, _ptr(last), C.byref(ok)))
    if not ok.value:
        return False
    rb, rc = last[:m], last[m:]                                                            # :262-264
    return (MultilinearKZG.verify(trusted_setup, proof.input_polynomial_commitment, rb, proof.input_rb_proof)
            and MultilinearKZG.verify(trusted_setup, proof.input_polynomial_commitment, rc, proof.input_rc_proof))
