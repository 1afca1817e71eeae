"""`multilinear_kzg` crate mirror: the input commitment of succinct GKR on the GPU.

  TrustedSetup            multilinear_kzg/src/trusted_setup.rs:5-24   (g1_powers_of_tau in HBM, g2_powers_of_tau on the host)
  MultilinearKZG          multilinear_kzg/src/multilinear_kzg.rs:10-13, commit_to_polynomial :25-46, open_and_prove :51-127,
                          verify :132-159 (host only: pairing check)
  MultilinearKZGProof     multilinear_kzg/src/multilinear_kzg.rs:15-19

Same names, argument meaning and panic messages as the reference.  The curve is BLS12-381 (the only pairing the
reference instantiates); scalars are BLS12-381 Fr elements ((..., 4) uint64 Montgomery limbs) and G1 points are
(..., 12) uint64 arrays: affine x then y, Montgomery limbs, all zero = the point at infinity -- `G1Affine`'s coordinates.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from ._lib import vp
from .core import BLS12_381_FR, Context, DeviceTable, ReferencePanic, _ptr, as_elems
from .polynomials import MultilinearPolynomial


def as_points(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.shape[-1] != 12:
        raise ValueError("G1 points are (..., 12) uint64 limb arrays (affine x, y)")
    return a


class TrustedSetup:
    """`TrustedSetup<P>`: g1_powers_of_tau lives in HBM together with its partial sums over the leading variables."""

    def __init__(self, ctx: Context, handle, g2_powers_of_tau: Optional[np.ndarray] = None):
        self.ctx = ctx
        self._h = handle
        self.g2_powers_of_tau = g2_powers_of_tau     # (n, 24) or None (a prover-only setup)

    @property
    def h(self):
        if not self._h:
            raise ReferencePanic("use of a released TrustedSetup")
        return self._h

    # initialize_setup(taus) -- trusted_setup.rs:12-24
    @classmethod
    def initialize_setup(cls, ctx: Context, taus) -> "TrustedSetup":
        t = as_elems(taus).reshape(-1, 4)
        h = vp()
        ctx.check(ctx.lib.zk_kzg_setup_create(ctx.h, _ptr(t) if t.size else None, t.shape[0], C.byref(h)))
        g2 = np.zeros((t.shape[0], 24), dtype=np.uint64)                                    # trusted_setup.rs:65-78
        if ctx.lib.zk_kzg_g2_powers_of_tau(_ptr(t), t.shape[0], _ptr(g2)) != 0:
            raise ReferencePanic("requires at least one variable")
        return cls(ctx, h, g2)

    @classmethod
    def from_g1_powers_of_tau(cls, ctx: Context, points, g2_powers_of_tau=None) -> "TrustedSetup":
        """an existing setup (e.g. a ceremony's): 2^n affine G1 points (and, for verification, the n G2 points)"""
        pts = as_points(points).reshape(-1, 12)
        n = pts.shape[0].bit_length() - 1
        if pts.shape[0] == 0 or pts.shape[0] != 1 << n:
            raise ReferencePanic("Evaluated values must be a power of 2")
        h = vp()
        ctx.check(ctx.lib.zk_kzg_setup_from_points(ctx.h, _ptr(pts), n, C.byref(h)))
        g2 = None if g2_powers_of_tau is None else np.ascontiguousarray(g2_powers_of_tau, dtype=np.uint64).reshape(-1, 24)
        return cls(ctx, h, g2)

    def number_of_variables(self) -> int:
        return int(self.ctx.lib.zk_kzg_setup_num_vars(self.h))

    @property
    def g1_powers_of_tau(self) -> np.ndarray:
        return self.level(0)

    def level(self, k: int) -> np.ndarray:
        """the setup summed over its first k variables: 2^(n-k) points"""
        n = self.number_of_variables()
        out = np.zeros((1 << max(n - k, 0), 12), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.zk_kzg_setup_points(self.ctx.h, self.h, k, _ptr(out)))
        return out

    def release(self) -> None:
        if self._h:
            self.ctx.lib.zk_kzg_setup_free(self.ctx.h, self._h)
            self._h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.release()
        except Exception:
            pass


@dataclass
class MultilinearKZGProof:
    """multilinear_kzg.rs:15-19"""
    evaluation: np.ndarray     # (4,)
    proofs: np.ndarray         # (n, 12)


def _table_of(polynomial) -> Optional[DeviceTable]:
    if isinstance(polynomial, MultilinearPolynomial):
        return polynomial.table
    if isinstance(polynomial, DeviceTable):
        return polynomial
    return None


class MultilinearKZG:
    """`MultilinearKZG<F, P>`; `polynomial` is a MultilinearPolynomial / DeviceTable (resident) or host evaluations"""

    @staticmethod
    def commit_to_polynomial(polynomial, trusted_setup: TrustedSetup, sharded: bool = False) -> np.ndarray:
        """sharded=True (after sharded.init_comm; setup and polynomial replicated on every rank, polynomial resident): every rank
        sums its share of the points, all return the same commitment"""
        ctx = trusted_setup.ctx
        out = np.zeros(12, dtype=np.uint64)
        t = _table_of(polynomial)
        if sharded:
            if t is None:
                raise ValueError("the sharded calls take the polynomial as a DeviceTable / MultilinearPolynomial")
            ctx.check(ctx.lib.zk_kzg_commit_sharded(ctx.h, trusted_setup.h, t.h, _ptr(out)))
        elif t is not None:
            ctx.check(ctx.lib.zk_kzg_commit_device(ctx.h, trusted_setup.h, t.h, _ptr(out)))
        else:
            ev = as_elems(polynomial).reshape(-1, 4)
            ctx.check(ctx.lib.zk_kzg_commit(ctx.h, trusted_setup.h, _ptr(ev), ev.shape[0], _ptr(out)))
        return out

    @staticmethod
    def open_and_prove(polynomial, trusted_setup: TrustedSetup, opening_values, sharded: bool = False) -> MultilinearKZGProof:
        ctx = trusted_setup.ctx
        op = as_elems(opening_values).reshape(-1, 4)
        ev_out = np.zeros(4, dtype=np.uint64)
        proofs = np.zeros((max(op.shape[0], 1), 12), dtype=np.uint64)
        t = _table_of(polynomial)
        if sharded:
            if t is None:
                raise ValueError("the sharded calls take the polynomial as a DeviceTable / MultilinearPolynomial")
            ctx.check(ctx.lib.zk_kzg_open_sharded(ctx.h, trusted_setup.h, t.h, _ptr(op) if op.size else None, op.shape[0],
                                                  _ptr(ev_out), _ptr(proofs)))
        elif t is not None:
            ctx.check(ctx.lib.zk_kzg_open_device(ctx.h, trusted_setup.h, t.h, _ptr(op) if op.size else None, op.shape[0],
                                                 _ptr(ev_out), _ptr(proofs)))
        else:
            ev = as_elems(polynomial).reshape(-1, 4)
            ctx.check(ctx.lib.zk_kzg_open(ctx.h, trusted_setup.h, _ptr(ev), ev.shape[0], _ptr(op) if op.size else None,
                                          op.shape[0], _ptr(ev_out), _ptr(proofs)))
        return MultilinearKZGProof(ev_out, proofs[:op.shape[0]])


    @staticmethod
    def verify(trusted_setup: TrustedSetup, commitment, opening_values, proof: MultilinearKZGProof) -> bool:
        """multilinear_kzg.rs:132-159: e(C - v G1, G2) == prod e(Q_i, tau_i G2 - r_i G2); host only"""
        lib = trusted_setup.ctx.lib
        g2 = trusted_setup.g2_powers_of_tau
        if g2 is None:
            raise ValueError("this TrustedSetup carries no g2_powers_of_tau (prover-only)")
        op = as_elems(opening_values).reshape(-1, 4)
        pr = as_points(proof.proofs).reshape(-1, 12)
        ok = C.c_int(0)
        rc = lib.zk_kzg_verify(_ptr(g2), g2.shape[0], _ptr(as_points(commitment).reshape(12)), _ptr(op) if op.size else None,
                               op.shape[0], _ptr(as_elems(proof.evaluation).reshape(4)), _ptr(pr) if pr.size else None,
                               pr.shape[0], C.byref(ok))
        if rc == -1:
            raise ReferencePanic("Number of opening values must match number of proofs")
        if rc:
            return False            # a point that is not on its curve
        return bool(ok.value)


def g1_msm(ctx: Context, scalars, points) -> np.ndarray:
    """sum_i scalars[i] * points[i] (`.map(|(v, p)| p.mul_bigint(v.into_bigint())).sum()`)"""
    s = as_elems(scalars).reshape(-1, 4)
    p = as_points(points).reshape(-1, 12)
    if s.shape[0] != p.shape[0]:
        raise ValueError("one scalar per point")
    out = np.zeros(12, dtype=np.uint64)
    ctx.check(ctx.lib.zk_g1_msm(ctx.h, _ptr(s) if s.size else None, _ptr(p) if p.size else None, s.shape[0], _ptr(out)))
    return out


assert BLS12_381_FR == 2
