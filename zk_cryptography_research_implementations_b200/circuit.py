"""`circuit` crate mirror: circuit/src/arithmetic_circuit.rs (the GKR input model, host side)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from enum import IntEnum
from typing import List, Sequence

import numpy as np

from . import _lib
from ._lib import u64p
from .core import Context, ReferencePanic, _ptr, as_elems, fe_from_int
from .polynomials import MultilinearPolynomial


class Operator(IntEnum):            # arithmetic_circuit.rs:5-8
    Add = 0
    Mul = 1


@dataclass
class Gate:                         # arithmetic_circuit.rs:10-16,33-47
    left_index: int
    right_index: int
    output_index: int
    operator: Operator

    @classmethod
    def new(cls, left_index, right_index, output_index, operator):
        return cls(left_index, right_index, output_index, Operator(operator))


@dataclass
class Layer:                        # arithmetic_circuit.rs:18-20,50-54
    gates: List[Gate]

    @classmethod
    def new(cls, gates):
        return cls(list(gates))


class _Desc(C.Structure):
    _fields_ = [("n_layers", C.c_uint32), ("layer_off", u64p), ("left", C.POINTER(C.c_uint32)),
                ("right", C.POINTER(C.c_uint32)), ("out", C.POINTER(C.c_uint32)), ("op", C.POINTER(C.c_uint8))]


@dataclass
class CircuitEvaluationResult:      # arithmetic_circuit.rs:27-30
    output: np.ndarray
    layer_evaluations: List[np.ndarray]


def num_of_layer_variables(layer_index: int) -> int:            # arithmetic_circuit.rs:166-178
    return 3 if layer_index == 0 else layer_index + 2 * (layer_index + 1)


def convert_to_binary_and_to_decimal(layer_index: int, a: int, b: int, c: int) -> int:   # :180-200
    s = format(a, "0>%db" % layer_index) + format(b, "0>%db" % (layer_index + 1)) + format(c, "0>%db" % (layer_index + 1))
    return int(s, 2)


class Circuit:
    """`Circuit<F>` (arithmetic_circuit.rs:22-25); layers[0] is the output layer."""

    def __init__(self, field: int, layers: Sequence[Layer]):
        self.field = field
        self.layers = list(layers)
        flat = [g for l in self.layers for g in l.gates]
        off = [0]
        for l in self.layers:
            off.append(off[-1] + len(l.gates))
        self._off = np.array(off, dtype=np.uint64)
        self._left = np.array([g.left_index for g in flat], dtype=np.uint32)
        self._right = np.array([g.right_index for g in flat], dtype=np.uint32)
        self._out = np.array([g.output_index for g in flat], dtype=np.uint32)
        self._op = np.array([int(g.operator) for g in flat], dtype=np.uint8)
        self.desc = _Desc(len(self.layers), self._off.ctypes.data_as(u64p), self._left.ctypes.data_as(C.POINTER(C.c_uint32)),
                          self._right.ctypes.data_as(C.POINTER(C.c_uint32)), self._out.ctypes.data_as(C.POINTER(C.c_uint32)),
                          self._op.ctypes.data_as(C.POINTER(C.c_uint8)))

    @classmethod
    def new(cls, field: int, layers):
        return cls(field, layers)

    def evaluate(self, values) -> CircuitEvaluationResult:      # arithmetic_circuit.rs:65-109
        lib = _lib.load()
        values = as_elems(values).reshape(-1, 4)
        L = len(self.layers)
        sizes = np.zeros(L + 1, dtype=np.uint64)
        cap = values.shape[0] + sum(max((g.output_index for g in l.gates), default=0) + 1 for l in self.layers)
        out = np.zeros((cap, 4), dtype=np.uint64)
        rc = lib.zk_circuit_evaluate(self.field, C.byref(self.desc), _ptr(values), values.shape[0], _ptr(sizes), _ptr(out), cap)
        if rc == _lib.ZK_ERR_ASSERT:
            raise ReferencePanic("index out of bounds")
        if rc:
            raise ValueError("zk_circuit_evaluate failed (%d)" % rc)
        evs, o = [], 0
        for s in sizes:
            evs.append(out[o:o + int(s)].copy())
            o += int(s)
        return CircuitEvaluationResult(evs[0], evs)

    @staticmethod
    def w_i_polynomial(ctx: Context, circuit_evaluation: CircuitEvaluationResult, layer_index: int) -> MultilinearPolynomial:
        if layer_index >= len(circuit_evaluation.layer_evaluations):      # arithmetic_circuit.rs:118-121
            raise ReferencePanic("layer index out of bounds")
        return MultilinearPolynomial.new(ctx, circuit_evaluation.layer_evaluations[layer_index])

    def add_i_and_mul_i_mle(self, ctx: Context, layer_index: int):        # arithmetic_circuit.rs:126-163 (dense)
        n = 1 << num_of_layer_variables(layer_index)
        one = fe_from_int(self.field, 1)
        add = np.zeros((n, 4), dtype=np.uint64)
        mul = np.zeros((n, 4), dtype=np.uint64)
        for g in self.layers[layer_index].gates:
            pos = convert_to_binary_and_to_decimal(layer_index, g.output_index, g.left_index, g.right_index)
            (add if g.operator == Operator.Add else mul)[pos] = one
        return MultilinearPolynomial.new(ctx, add), MultilinearPolynomial.new(ctx, mul)
