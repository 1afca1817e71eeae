"""`polynomials` crate mirror on device tables.

  MultilinearPolynomial   polynomials/src/multilinear/evaluation_form.rs
  ProductPolynomial       polynomials/src/composed/product_polynomial.rs
  SumPolynomial           polynomials/src/composed/sum_polynomial.rs
  DenseUnivariatePolynomial (host, coefficient form)  polynomials/src/univariate/dense_univariate.rs

Same names, argument meaning and panic messages as the reference; `evaluated_values` lives in HBM.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np

from ._lib import u8p, vp
from .core import Context, DeviceTable, ReferencePanic, _ptr, as_elems, fe_binop, fe_from_int


class MultilinearPolynomial:
    """`MultilinearPolynomial<F>` (evaluation_form.rs:7-9) -- evaluations over {0,1}^n, variable 0 = MSB."""

    def __init__(self, ctx: Context, table: DeviceTable):
        self.ctx = ctx
        self.table = table

    # new(&[F]) -- evaluation_form.rs:12-18
    @classmethod
    def new(cls, ctx: Context, evaluated_values) -> "MultilinearPolynomial":
        ev = as_elems(evaluated_values).reshape(-1, 4)
        if ev.shape[0] == 0 or ev.shape[0] & (ev.shape[0] - 1):
            raise ReferencePanic("Evaluated values must be a power of 2")
        return cls(ctx, ctx.upload(ev))

    @property
    def evaluated_values(self) -> np.ndarray:
        return self.table.download()

    def __len__(self) -> int:
        return len(self.table)

    def clone(self) -> "MultilinearPolynomial":
        return MultilinearPolynomial(self.ctx, self.table.clone())

    # evaluate(&self, &[F]) -> F -- evaluation_form.rs:21-33
    def evaluate(self, values) -> np.ndarray:
        vals = as_elems(values).reshape(-1, 4)
        out = np.zeros(4, dtype=np.uint64)
        self.ctx.check(self.ctx.lib.zk_mle_evaluate(self.ctx.h, self.table.h, _ptr(vals) if vals.size else None,
                                                    vals.shape[0], _ptr(out)))
        return out

    # convert_to_bytes -- evaluation_form.rs:35-43
    def convert_to_bytes(self) -> bytes:
        buf = (C.c_uint8 * (32 * len(self)))()
        self.ctx.check(self.ctx.lib.zk_mle_to_bytes(self.ctx.h, self.table.h, C.cast(buf, u8p)))
        return bytes(buf)

    # number_of_variables -- evaluation_form.rs:45-47
    def number_of_variables(self) -> int:
        return len(self).bit_length() - 1

    # scalar_mul -- evaluation_form.rs:49-57
    def scalar_mul(self, scalar) -> "MultilinearPolynomial":
        h = vp()
        self.ctx.check(self.ctx.lib.zk_mle_scalar_mul(self.ctx.h, self.table.h, _ptr(as_elems(scalar)), C.byref(h)))
        return MultilinearPolynomial(self.ctx, DeviceTable(self.ctx, h))

    # partial_evaluate(&Vec<F>, evaluating_variable, value) -> Self -- evaluation_form.rs:61-106
    @staticmethod
    def partial_evaluate(polynomial: "MultilinearPolynomial", evaluating_variable: int, value) -> "MultilinearPolynomial":
        out = polynomial.clone()
        out.partial_evaluate_in_place(evaluating_variable, value)
        return out

    def partial_evaluate_in_place(self, evaluating_variable: int, value) -> None:
        self.ctx.check(self.ctx.lib.zk_mle_partial_evaluate(self.ctx.h, self.table.h, evaluating_variable,
                                                            _ptr(as_elems(value))))

    # polynomial_tensor_add / polynomial_tensor_mul -- evaluation_form.rs:108-143
    @staticmethod
    def polynomial_tensor_add(w_b: "MultilinearPolynomial", w_c: "MultilinearPolynomial") -> "MultilinearPolynomial":
        h = vp()
        w_b.ctx.check(w_b.ctx.lib.zk_mle_tensor_add(w_b.ctx.h, w_b.table.h, w_c.table.h, C.byref(h)))
        return MultilinearPolynomial(w_b.ctx, DeviceTable(w_b.ctx, h))

    @staticmethod
    def polynomial_tensor_mul(w_b: "MultilinearPolynomial", w_c: "MultilinearPolynomial") -> "MultilinearPolynomial":
        h = vp()
        w_b.ctx.check(w_b.ctx.lib.zk_mle_tensor_mul(w_b.ctx.h, w_b.table.h, w_c.table.h, C.byref(h)))
        return MultilinearPolynomial(w_b.ctx, DeviceTable(w_b.ctx, h))

    # add_polynomials -- evaluation_form.rs:145-163
    @staticmethod
    def add_polynomials(poly1: "MultilinearPolynomial", poly2: "MultilinearPolynomial") -> "MultilinearPolynomial":
        h = vp()
        poly1.ctx.check(poly1.ctx.lib.zk_mle_add(poly1.ctx.h, poly1.table.h, poly2.table.h, C.byref(h)))
        return MultilinearPolynomial(poly1.ctx, DeviceTable(poly1.ctx, h))


class ProductPolynomial:
    """`ProductPolynomial<F>` (product_polynomial.rs:6-8): the pointwise product of its members."""

    def __init__(self, polynomials: Sequence[MultilinearPolynomial]):      # new :11-24
        polynomials = list(polynomials)
        n = polynomials[0].number_of_variables()
        if any(p.number_of_variables() != n for p in polynomials):
            raise ReferencePanic("different number of variables")
        self.polynomials = polynomials

    def evaluate(self, values) -> np.ndarray:                              # :26-34
        field = self.polynomials[0].ctx.field
        result = fe_from_int(field, 1)
        for poly in self.polynomials:
            result = fe_binop("mul", field, result, poly.evaluate(values))
        return result

    def partial_evaluate(self, evaluating_variable: int, value) -> List[MultilinearPolynomial]:   # :36-54
        return [MultilinearPolynomial.partial_evaluate(p, evaluating_variable, value) for p in self.polynomials]

    def multiply_polynomials_element_wise(self) -> MultilinearPolynomial:  # :58-73
        if len(self.polynomials) < 2:
            raise ReferencePanic("more than one polynomial required for mul operation")
        ctx = self.polynomials[0].ctx
        # P = 1 sum of one product: reuse the SumPolynomial reducer on [this, 1*1]?  Simpler: chain tensor-free muls
        acc = self.polynomials[0].clone()
        for p in self.polynomials[1:]:
            acc = _elementwise_mul(acc, p)
        return acc

    def convert_to_bytes(self) -> bytes:                                   # :75-83
        return b"".join(p.convert_to_bytes() for p in self.polynomials)

    def degree(self) -> int:                                               # :85-87
        return len(self.polynomials)


def _elementwise_mul(a: MultilinearPolynomial, b: MultilinearPolynomial) -> MultilinearPolynomial:
    """a[i] * b[i] through the SumPolynomial reducer kernel: (a*b) + (0*0) has P = 2, D = 2."""
    ctx = a.ctx
    zeros = np.zeros((len(a), 4), dtype=np.uint64)
    z1, z2 = ctx.upload(zeros), ctx.upload(zeros)
    sp = SumPolynomial([ProductPolynomial([a.clone(), b.clone()]),
                        ProductPolynomial([MultilinearPolynomial(ctx, z1), MultilinearPolynomial(ctx, z2)])])
    return sp.add_polynomials_element_wise()


class SumPolynomial:
    """`SumPolynomial<F>` (sum_polynomial.rs:7-9): the sum of its product polynomials."""

    def __init__(self, product_polynomials: Sequence[ProductPolynomial]):  # new :12-28
        product_polynomials = list(product_polynomials)
        n = product_polynomials[0].polynomials[0].number_of_variables()
        if any(p.number_of_variables() != n for prod in product_polynomials for p in prod.polynomials):
            raise ReferencePanic("different number of variables")
        self.product_polynomials = product_polynomials

    @property
    def ctx(self) -> Context:
        return self.product_polynomials[0].polynomials[0].ctx

    def evaluate(self, values) -> np.ndarray:                              # :30-38
        field = self.ctx.field
        result = np.zeros(4, dtype=np.uint64)
        for prod in self.product_polynomials:
            result = fe_binop("add", field, result, prod.evaluate(values))
        return result

    def partial_evaluate(self, evaluating_variable: int, value) -> "SumPolynomial":   # :40-53
        return SumPolynomial([ProductPolynomial(prod.partial_evaluate(evaluating_variable, value))
                              for prod in self.product_polynomials])

    def add_polynomials_element_wise(self) -> MultilinearPolynomial:       # :57-76
        if len(self.product_polynomials) < 2:
            raise ReferencePanic("more than one product polynomial required for add operation")
        if any(len(prod.polynomials) < 2 for prod in self.product_polynomials):
            raise ReferencePanic("more than one polynomial required for mul operation")
        ctx = self.ctx
        sp = self._device_sumpoly(clone=True)
        try:
            h = vp()
            ctx.check(ctx.lib.zk_sumpoly_reduce(ctx.h, sp, C.byref(h)))
            return MultilinearPolynomial(ctx, DeviceTable(ctx, h))
        finally:
            ctx.lib.zk_sumpoly_free(ctx.h, sp)

    def convert_to_bytes(self) -> bytes:                                   # :78-86
        return b"".join(p.convert_to_bytes() for p in self.product_polynomials)

    def degree(self) -> int:                                               # :88-90
        return self.product_polynomials[0].degree()

    def number_of_variables(self) -> int:                                  # :92-94
        return self.product_polynomials[0].polynomials[0].number_of_variables()

    # ---- device handle (zk_sumpoly); the handle owns its tables
    def _device_sumpoly(self, clone: bool):
        """clone=True: the sumpoly gets copies (the reference clones: `self` stays usable).  clone=False: the tables of
        `self` are MOVED into the sumpoly -- a table that sits in two slots (f*f, W_b == W_c) is cloned for its second
        slot, and afterwards every MultilinearPolynomial of `self` holds a dead handle (any later use raises instead of
        touching freed memory)."""
        ctx = self.ctx
        D = self.degree()
        if any(len(prod.polynomials) != D for prod in self.product_polynomials):
            raise ValueError("all products must have the same number of factors on the device path")
        polys = [p for prod in self.product_polynomials for p in prod.polynomials]
        tabs, made, seen = [], [], set()
        for p in polys:
            p.table.h               # raises if the table was consumed by an earlier prove
            if clone or id(p.table) in seen:
                t = p.table.clone()
                made.append(t)
            else:
                t = p.table
                seen.add(id(t))
            tabs.append(t)
        arr = (vp * len(tabs))(*[t.h for t in tabs])
        h = vp()
        try:
            ctx.check(ctx.lib.zk_sumpoly_create(ctx.h, arr, len(self.product_polynomials), D, C.byref(h)))
        except Exception:
            for t in made:          # nothing was transferred: drop the copies, the caller keeps its own tables
                t.free()
            raise
        for t in tabs:              # ownership is now the sumpoly's: the copies and (clone=False) the caller's tables die here
            t.release()
        return h


class DenseUnivariatePolynomial:
    """`DenseUnivariatePolynomial<F>` (dense_univariate.rs:4-6), coefficient form, host side."""

    def __init__(self, field: int, coefficients):
        self.field = field
        self.coefficients = as_elems(coefficients).reshape(-1, 4).copy()

    def degree(self) -> int:                                               # :15-17
        return self.coefficients.shape[0] - 1

    def evaluate(self, value) -> np.ndarray:                               # :57-68
        result = np.zeros(4, dtype=np.uint64)
        power = fe_from_int(self.field, 1)
        for c in self.coefficients:
            result = fe_binop("add", self.field, result, fe_binop("mul", self.field, c, power))
            power = fe_binop("mul", self.field, power, value)
        return result
