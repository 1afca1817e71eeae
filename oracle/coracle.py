"""ctypes binding of oracle/libzkoracle.so (the C restatement of the reference path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs; never by the product package.  Elements cross this binding as numpy uint64 arrays of shape
(..., 4): little-endian limbs, Montgomery form -- the same layout the CUDA library uses.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libzkoracle.so")

BN254_FQ, BN254_FR, BLS12_381_FR = 0, 1, 2
u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    deps = [os.path.join(_HERE, f) for f in ("zkoracle.c", "zkoracle.h", "field_consts.h", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["make", "-C", _HERE, "libzkoracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def set_threads(n: int) -> None:
    """all-core mode of the data-parallel loops (NOT the reference's behaviour -- it is single-threaded; used for the
    labelled 'stronger than the reference' CPU baseline of bench.py).  1 restores the restated reference loops."""
    try:
        lib().zko_set_threads(C.c_int(int(n)))
    except AttributeError:      # a library built before the all-core mode existed: single-threaded by construction
        if int(n) != 1:
            raise


def get_threads() -> int:
    return int(lib().zko_get_threads())


def openmp_enabled() -> bool:
    try:
        return bool(lib().zko_openmp_enabled())
    except AttributeError:
        return False


class _Circuit(C.Structure):
    _fields_ = [("n_layers", C.c_uint32), ("layer_off", u64p), ("left", u32p), ("right", u32p),
                ("out", u32p), ("op", u8p)]


class _GkrProof(C.Structure):
    _fields_ = [("output", u64p), ("n_output", C.c_uint64), ("claimed_sum", C.c_uint64 * 4),
                ("layer_claims", u64p), ("coeffs", u64p), ("challenges", u64p), ("wb", u64p), ("wc", u64p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.zko_transcript_new.restype = C.c_void_p
        _lib.zko_transcript_clone.restype = C.c_void_p
        _lib.zko_transcript_clone.argtypes = [C.c_void_p]
        _lib.zko_transcript_free.argtypes = [C.c_void_p]
        _lib.zko_transcript_append.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        _lib.zko_transcript_sample.argtypes = [C.c_void_p, C.c_char_p]
        _lib.zko_transcript_challenge.argtypes = [C.c_void_p, C.c_int, u64p]
        _lib.zko_gkr_total_rounds.restype = C.c_uint64
        _lib.zko_gate_position.restype = C.c_uint64
        _lib.zko_gate_position.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64]
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def _arr(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


# ------------------------------------------------------------------ field helpers
def from_ints(fid: int, vals: Sequence[int]) -> np.ndarray:
    """canonical Python ints -> Montgomery limb array (len, 4)"""
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    tmp = np.zeros(4, dtype=np.uint64)
    for i, v in enumerate(vals):
        for k in range(4):
            tmp[k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
        lib().zko_fe_from_canonical(fid, _p(tmp), _p(out[i]))
    return out


def to_ints(fid: int, a: np.ndarray) -> List[int]:
    a = _arr(a).reshape(-1, 4)
    tmp = np.zeros(4, dtype=np.uint64)
    res = []
    for i in range(a.shape[0]):
        lib().zko_fe_to_canonical(fid, _p(a[i]), _p(tmp))
        res.append(sum(int(tmp[k]) << (64 * k) for k in range(4)))
    return res


def fe_op(name: str, fid: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    out = np.zeros(4, dtype=np.uint64)
    getattr(lib(), "zko_fe_" + name)(fid, _p(_arr(a)), _p(_arr(b)), _p(out))
    return out


def keccak256(data: bytes) -> bytes:
    out = C.create_string_buffer(32)
    lib().zko_keccak256(data, C.c_size_t(len(data)), out)
    return out.raw


class Transcript:
    def __init__(self, handle=None):
        self.h = C.c_void_p(handle if handle is not None else lib().zko_transcript_new())

    def clone(self) -> "Transcript":
        return Transcript(lib().zko_transcript_clone(self.h))

    def append(self, data: bytes) -> None:
        lib().zko_transcript_append(self.h, data, len(data))

    def sample_random_challenge(self) -> bytes:
        out = C.create_string_buffer(32)
        lib().zko_transcript_sample(self.h, out)
        return out.raw

    def random_challenge_as_field_element(self, fid: int) -> np.ndarray:
        out = np.zeros(4, dtype=np.uint64)
        lib().zko_transcript_challenge(self.h, fid, _p(out))
        return out

    def __del__(self):
        try:
            lib().zko_transcript_free(self.h)
        except Exception:
            pass


# ------------------------------------------------------------------ MLE
def mle_partial_evaluate(fid: int, table: np.ndarray, var: int, r: np.ndarray) -> np.ndarray:
    table = _arr(table).reshape(-1, 4)
    n = table.shape[0]
    out = np.zeros((max(n // 2, 1), 4), dtype=np.uint64)
    rc = lib().zko_mle_partial_evaluate(fid, _p(table), C.c_uint64(n), C.c_uint32(var), _p(_arr(r)), _p(out))
    if rc:
        raise AssertionError("Evaluated values must be a power of 2")
    return out[: n // 2]


def mle_evaluate(fid: int, table: np.ndarray, rs: np.ndarray) -> np.ndarray:
    table = _arr(table).reshape(-1, 4)
    rs = _arr(rs).reshape(-1, 4)
    out = np.zeros(4, dtype=np.uint64)
    rc = lib().zko_mle_evaluate(fid, _p(table), C.c_uint64(table.shape[0]), _p(rs), C.c_uint32(rs.shape[0]), _p(out))
    if rc:
        raise AssertionError("Evaluated values must be a power of 2")
    return out


def mle_to_bytes(fid: int, table: np.ndarray) -> bytes:
    table = _arr(table).reshape(-1, 4)
    out = C.create_string_buffer(32 * table.shape[0])
    lib().zko_mle_to_bytes(fid, _p(table), C.c_uint64(table.shape[0]), C.cast(out, u8p))
    return out.raw


def split_and_sum(fid: int, table: np.ndarray) -> np.ndarray:
    table = _arr(table).reshape(-1, 4)
    out = np.zeros((2, 4), dtype=np.uint64)
    lib().zko_split_and_sum(fid, _p(table), C.c_uint64(table.shape[0]), _p(out))
    return out


def tensor(fid: int, op: str, wb: np.ndarray, wc: np.ndarray) -> np.ndarray:
    wb, wc = _arr(wb).reshape(-1, 4), _arr(wc).reshape(-1, 4)
    if wb.shape != wc.shape:
        raise AssertionError("Different polynomial length")
    n = wb.shape[0]
    out = np.zeros((n * n, 4), dtype=np.uint64)
    getattr(lib(), "zko_mle_tensor_" + op)(fid, _p(wb), _p(wc), C.c_uint64(n), _p(out))
    return out


def sumpoly_reduce(fid: int, tables: np.ndarray) -> np.ndarray:
    """tables: (P, D, len, 4)"""
    tables = _arr(tables)
    P, D, n, _ = tables.shape
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().zko_sumpoly_reduce(fid, _p(tables), C.c_uint32(P), C.c_uint32(D), C.c_uint64(n), _p(out))
    return out


def univariate_evaluate(fid: int, coeffs: np.ndarray, x: np.ndarray) -> np.ndarray:
    coeffs = _arr(coeffs).reshape(-1, 4)
    out = np.zeros(4, dtype=np.uint64)
    lib().zko_univariate_evaluate(fid, _p(coeffs), C.c_uint32(coeffs.shape[0]), _p(_arr(x)), _p(out))
    return out


def lagrange_interpolate(fid: int, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    xs, ys = _arr(xs).reshape(-1, 4), _arr(ys).reshape(-1, 4)
    out = np.zeros_like(xs)
    lib().zko_lagrange_interpolate(fid, _p(xs), _p(ys), C.c_uint32(xs.shape[0]), _p(out))
    return out


# ------------------------------------------------------------------ sumchecks
def basic_prove(fid: int, table: np.ndarray):
    """-> (claimed_sum (4,), round_polys (n,2,4), challenges (n,4), final_eval (4,))"""
    table = _arr(table).reshape(-1, 4)
    N = table.shape[0]
    n = N.bit_length() - 1
    claimed = np.zeros(4, dtype=np.uint64)
    rp = np.zeros((max(n, 1), 2, 4), dtype=np.uint64)
    ch = np.zeros((max(n, 1), 4), dtype=np.uint64)
    fin = np.zeros(4, dtype=np.uint64)
    rc = lib().zko_basic_prove(fid, _p(table), C.c_uint64(N), _p(claimed), _p(rp), _p(ch), _p(fin))
    if rc:
        raise AssertionError("Evaluated values must be a power of 2")
    return claimed, rp[:n], ch[:n], fin


def basic_verify(fid: int, table: np.ndarray, claimed: np.ndarray, round_polys: np.ndarray) -> bool:
    table = _arr(table).reshape(-1, 4)
    rp = _arr(round_polys).reshape(-1, 2, 4)
    return bool(lib().zko_basic_verify(fid, _p(table), C.c_uint64(table.shape[0]), _p(_arr(claimed)),
                                       _p(rp) if rp.size else None, C.c_uint32(rp.shape[0])))


def generate_round_univariate(fid: int, tables: np.ndarray) -> np.ndarray:
    tables = _arr(tables)
    P, D, n, _ = tables.shape
    out = np.zeros((D + 1, 4), dtype=np.uint64)
    lib().zko_generate_round_univariate(fid, _p(tables), C.c_uint32(P), C.c_uint32(D), C.c_uint64(n), _p(out))
    return out


def product_prove(fid: int, tables: np.ndarray, claimed: np.ndarray, transcript: Transcript):
    """tables (P, D, len, 4) -> (coeffs (n, D+1, 4), challenges (n, 4), final_tables (P, D, 4))"""
    tables = _arr(tables)
    P, D, N, _ = tables.shape
    n = N.bit_length() - 1
    coeffs = np.zeros((max(n, 1), D + 1, 4), dtype=np.uint64)
    ch = np.zeros((max(n, 1), 4), dtype=np.uint64)
    fin = np.zeros((P, D, 4), dtype=np.uint64)
    rc = lib().zko_product_prove(fid, _p(tables), C.c_uint32(P), C.c_uint32(D), C.c_uint64(N), _p(_arr(claimed)),
                                 transcript.h, _p(coeffs), _p(ch), _p(fin))
    if rc:
        raise AssertionError("oracle product_prove rejected the input (P,D >= 2 and power-of-two tables required)")
    return coeffs[:n], ch[:n], fin


def product_verify(fid: int, claimed: np.ndarray, coeffs: np.ndarray, transcript: Transcript):
    coeffs = _arr(coeffs)
    n, Dp1, _ = coeffs.shape
    ch = np.zeros((max(n, 1), 4), dtype=np.uint64)
    last = np.zeros(4, dtype=np.uint64)
    ok = lib().zko_product_verify(fid, _p(_arr(claimed)), _p(coeffs), C.c_uint32(n), C.c_uint32(Dp1 - 1),
                                  transcript.h, _p(ch), _p(last))
    return bool(ok), ch[:n], last


# ------------------------------------------------------------------ circuit + GKR
class Circuit:
    """layers: list (output layer first) of lists of (left, right, out, op) with op 0 = Add, 1 = Mul."""

    def __init__(self, layers: Sequence[Sequence[Tuple[int, int, int, int]]]):
        self.layers = [list(l) for l in layers]
        off = [0]
        for l in self.layers:
            off.append(off[-1] + len(l))
        flat = [g for l in self.layers for g in l]
        self.off = np.array(off, dtype=np.uint64)
        self.left = np.array([g[0] for g in flat], dtype=np.uint32)
        self.right = np.array([g[1] for g in flat], dtype=np.uint32)
        self.out = np.array([g[2] for g in flat], dtype=np.uint32)
        self.op = np.array([g[3] for g in flat], dtype=np.uint8)
        self.c = _Circuit(len(self.layers), self.off.ctypes.data_as(u64p), self.left.ctypes.data_as(u32p),
                          self.right.ctypes.data_as(u32p), self.out.ctypes.data_as(u32p), self.op.ctypes.data_as(u8p))

    def evaluate(self, fid: int, inputs: np.ndarray) -> List[np.ndarray]:
        inputs = _arr(inputs).reshape(-1, 4)
        L = len(self.layers)
        sizes = np.zeros(L + 1, dtype=np.uint64)
        cap = inputs.shape[0] + sum((max((g[2] for g in l), default=0) + 1) for l in self.layers)
        vals = np.zeros((cap, 4), dtype=np.uint64)
        rc = lib().zko_circuit_evaluate(fid, C.byref(self.c), _p(inputs), C.c_uint64(inputs.shape[0]), _p(sizes),
                                        _p(vals), C.c_uint64(cap))
        if rc:
            raise AssertionError("circuit evaluation failed (%d)" % rc)
        out, o = [], 0
        for s in sizes:
            out.append(vals[o:o + int(s)].copy())
            o += int(s)
        return out

    def add_i_mul_i(self, fid: int, layer: int):
        n = 1 << lib().zko_num_of_layer_variables(C.c_uint32(layer))
        a = np.zeros((n, 4), dtype=np.uint64)
        m = np.zeros((n, 4), dtype=np.uint64)
        lib().zko_add_i_mul_i(fid, C.byref(self.c), C.c_uint32(layer), _p(a), _p(m))
        return a, m


class GkrProof:
    def __init__(self, n_layers: int, n_out_cap: int):
        L = n_layers
        R = int(lib().zko_gkr_total_rounds(C.c_uint32(L)))
        self.L, self.R = L, R
        self.output = np.zeros((max(n_out_cap, 1), 4), dtype=np.uint64)
        self.layer_claims = np.zeros((L, 4), dtype=np.uint64)
        self.coeffs = np.zeros((R, 3, 4), dtype=np.uint64)
        self.challenges = np.zeros((R, 4), dtype=np.uint64)
        self.wb = np.zeros((max(L - 1, 1), 4), dtype=np.uint64)
        self.wc = np.zeros((max(L - 1, 1), 4), dtype=np.uint64)
        self.c = _GkrProof(_p(self.output), 0, (C.c_uint64 * 4)(), _p(self.layer_claims), _p(self.coeffs),
                           _p(self.challenges), _p(self.wb), _p(self.wc))

    @property
    def claimed_sum(self) -> np.ndarray:
        return np.array(list(self.c.claimed_sum), dtype=np.uint64)

    @property
    def circuit_output(self) -> np.ndarray:
        return self.output[: int(self.c.n_output)]


def gkr_prove(fid: int, circuit: Circuit, inputs: np.ndarray) -> GkrProof:
    inputs = _arr(inputs).reshape(-1, 4)
    cap = max((g[2] for g in circuit.layers[0]), default=0) + 1
    pf = GkrProof(len(circuit.layers), cap)
    rc = lib().zko_gkr_prove(fid, C.byref(circuit.c), _p(inputs), C.c_uint64(inputs.shape[0]), C.byref(pf.c))
    if rc:
        raise AssertionError("oracle gkr_prove failed (%d)" % rc)
    return pf


def gkr_verify(fid: int, circuit: Circuit, proof: GkrProof, inputs: np.ndarray) -> bool:
    inputs = _arr(inputs).reshape(-1, 4)
    return bool(lib().zko_gkr_verify(fid, C.byref(circuit.c), C.byref(proof.c), _p(inputs), C.c_uint64(inputs.shape[0])))


# ------------------------------------------------------------------ GKR over explicit layer widths (gate-list form)
class SparseCircuit:
    """layer_bits[li] = log2(#values of layer li), li = 0..L (last = inputs); layers: per layer a list of
    (left, right, out, op) tuples or an (n, 4) integer array with those columns."""

    def __init__(self, layer_bits: Sequence[int], layers):
        self.bits = np.array([int(b) for b in layer_bits], dtype=np.uint32)
        self.L = len(layers)
        assert len(self.bits) == self.L + 1
        arrs = [np.asarray(l, dtype=np.int64).reshape(-1, 4) for l in layers]
        off = [0]
        for a in arrs:
            off.append(off[-1] + a.shape[0])
        flat = np.concatenate(arrs) if arrs else np.zeros((0, 4), dtype=np.int64)
        self.off = np.array(off, dtype=np.uint64)
        self.left = np.ascontiguousarray(flat[:, 0].astype(np.uint32))
        self.right = np.ascontiguousarray(flat[:, 1].astype(np.uint32))
        self.out = np.ascontiguousarray(flat[:, 2].astype(np.uint32))
        self.op = np.ascontiguousarray(flat[:, 3].astype(np.uint8))

    def total_rounds(self) -> int:
        return int(sum(2 * int(b) for b in self.bits[1:]))

    def _args(self):
        return (C.c_uint32(self.L), self.bits.ctypes.data_as(u32p), self.off.ctypes.data_as(u64p), self.left.ctypes.data_as(u32p),
                self.right.ctypes.data_as(u32p), self.out.ctypes.data_as(u32p), self.op.ctypes.data_as(u8p))


class SparseGkrProof(GkrProof):
    def __init__(self, circuit: SparseCircuit):
        L, R = circuit.L, circuit.total_rounds()
        self.L, self.R = L, R
        self.output = np.zeros((1 << int(max(circuit.bits[0], 1)), 4), dtype=np.uint64)
        self.layer_claims = np.zeros((L, 4), dtype=np.uint64)
        self.coeffs = np.zeros((max(R, 1), 3, 4), dtype=np.uint64)
        self.challenges = np.zeros((max(R, 1), 4), dtype=np.uint64)
        self.wb = np.zeros((max(L - 1, 1), 4), dtype=np.uint64)
        self.wc = np.zeros((max(L - 1, 1), 4), dtype=np.uint64)
        self.c = _GkrProof(_p(self.output), 0, (C.c_uint64 * 4)(), _p(self.layer_claims), _p(self.coeffs),
                           _p(self.challenges), _p(self.wb), _p(self.wc))


def gkr_prove_sparse(fid: int, circuit: SparseCircuit, inputs: np.ndarray) -> SparseGkrProof:
    """gkr_protocol::prove (gkr_protocol.rs:26-143) with the wiring predicates evaluated from the gate list"""
    inputs = _arr(inputs).reshape(-1, 4)
    pf = SparseGkrProof(circuit)
    rc = lib().zko_gkr_prove_sparse(fid, *circuit._args(), _p(inputs), C.c_uint64(inputs.shape[0]), C.byref(pf.c))
    if rc:
        raise AssertionError("oracle gkr_prove_sparse failed (%d)" % rc)
    return pf


def gkr_verify_sparse(fid: int, circuit: SparseCircuit, proof, inputs: np.ndarray) -> bool:
    """gkr_protocol::verify (gkr_protocol.rs:146-236); `proof` is anything with a `.c` _GkrProof (see make_gkr_proof)"""
    inputs = _arr(inputs).reshape(-1, 4)
    return bool(lib().zko_gkr_verify_sparse(fid, *circuit._args(), C.byref(proof.c), _p(inputs), C.c_uint64(inputs.shape[0])))


def make_gkr_proof(circuit: SparseCircuit, output, claimed_sum, layer_claims, coeffs, wb, wc) -> SparseGkrProof:
    """wrap proof arrays (e.g. what the CUDA prover returned) for gkr_verify_sparse"""
    pf = SparseGkrProof(circuit)
    out = _arr(output).reshape(-1, 4)
    pf.output[: out.shape[0]] = out
    pf.c.n_output = out.shape[0]
    pf.layer_claims[:] = _arr(layer_claims).reshape(-1, 4)
    c = _arr(coeffs).reshape(-1, 3, 4)
    pf.coeffs[: c.shape[0]] = c
    for name, src in (("wb", wb), ("wc", wc)):
        a = _arr(src).reshape(-1, 4)
        getattr(pf, name)[: a.shape[0]] = a
    cs = _arr(claimed_sum).reshape(4)
    for k in range(4):
        pf.c.claimed_sum[k] = int(cs[k])
    return pf


def table_generate(fid: int, seed: int, table_id: int, n: int, first: int = 0, step: int = 1) -> np.ndarray:
    """the bench workload's seeded table, identical to the CUDA generator's (zk_table_generate)"""
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().zko_table_generate(fid, C.c_uint64(seed), C.c_uint64(table_id), C.c_uint64(n), C.c_uint64(first), C.c_uint64(step), _p(out))
    return out


# ------------------------------------------------------------------ multilinear KZG over BLS12-381 G1 (zkoracle_kzg.c)
FR = 2   # ZKO_BLS12_381_FR: the scalar field of every KZG call


def g1_from_ints(points) -> np.ndarray:
    """[(x, y) | None, ...] canonical Python ints -> (len, 12) Montgomery limb array (None = infinity = all zero)"""
    out = np.zeros((len(points), 12), dtype=np.uint64)
    tmp = np.zeros(6, dtype=np.uint64)
    for i, pt in enumerate(points):
        if pt is None:
            continue
        for c in range(2):
            for k in range(6):
                tmp[k] = (pt[c] >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
            lib().zko_fq_from_canonical(_p(tmp), _p(out[i, 6 * c:6 * c + 6]))
    return out


def g1_to_ints(a: np.ndarray):
    a = _arr(a).reshape(-1, 12)
    tmp = np.zeros(6, dtype=np.uint64)
    res = []
    for i in range(a.shape[0]):
        if not a[i].any():
            res.append(None)
            continue
        xy = []
        for c in range(2):
            lib().zko_fq_to_canonical(_p(np.ascontiguousarray(a[i, 6 * c:6 * c + 6])), _p(tmp))
            xy.append(sum(int(tmp[k]) << (64 * k) for k in range(6)))
        res.append(tuple(xy))
    return res


def g1_generator() -> np.ndarray:
    out = np.zeros(12, dtype=np.uint64)
    lib().zko_g1_generator(_p(out))
    return out


def g1_is_on_curve(p: np.ndarray) -> bool:
    return bool(lib().zko_g1_is_on_curve(_p(_arr(p).reshape(12))))


def g1_add(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    out = np.zeros(12, dtype=np.uint64)
    lib().zko_g1_add(_p(_arr(a).reshape(12)), _p(_arr(b).reshape(12)), _p(out))
    return out


def g1_mul(p: np.ndarray, k: int) -> np.ndarray:
    """mul_bigint by the canonical integer k < 2^256"""
    out = np.zeros(12, dtype=np.uint64)
    kk = np.array([(k >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
    lib().zko_g1_mul(_p(_arr(p).reshape(12)), _p(kk), _p(out))
    return out


def kzg_setup_g1(taus: np.ndarray) -> np.ndarray:
    """trusted_setup.rs:26-63: g1_powers_of_tau for taus (n, 4) -> (2^n, 12)"""
    taus = _arr(taus).reshape(-1, 4)
    n = taus.shape[0]
    out = np.zeros((1 << n, 12), dtype=np.uint64)
    rc = lib().zko_kzg_setup_g1(_p(taus), C.c_uint32(n), _p(out))
    if rc:
        raise AssertionError("requires at least one variable")
    return out


def kzg_commit(vals: np.ndarray, g1: np.ndarray) -> np.ndarray:
    """multilinear_kzg.rs:25-46"""
    vals = _arr(vals).reshape(-1, 4)
    g1 = _arr(g1).reshape(-1, 12)
    out = np.zeros(12, dtype=np.uint64)
    rc = lib().zko_kzg_commit(_p(vals), C.c_uint64(vals.shape[0]), _p(g1), C.c_uint64(g1.shape[0]), _p(out))
    if rc:
        raise AssertionError("Polynomial evaluation must match g1 length")
    return out


def kzg_open(vals: np.ndarray, g1: np.ndarray, opening: np.ndarray):
    """multilinear_kzg.rs:51-127 -> (evaluation (4,), proofs (n, 12))"""
    vals = _arr(vals).reshape(-1, 4)
    g1 = _arr(g1).reshape(-1, 12)
    opening = _arr(opening).reshape(-1, 4)
    nvars = vals.shape[0].bit_length() - 1
    ev = np.zeros(4, dtype=np.uint64)
    proofs = np.zeros((max(nvars, 1), 12), dtype=np.uint64)
    rc = lib().zko_kzg_open(_p(vals), C.c_uint32(nvars), _p(g1), C.c_uint64(g1.shape[0]), _p(opening),
                            C.c_uint32(opening.shape[0]), _p(ev), _p(proofs))
    if rc == -1:
        raise AssertionError("number of polynomial variables must match length of opening values")
    if rc:
        raise AssertionError("Opening values must match number of variables from trusted setup")
    return ev, proofs[:nvars]


def kzg_verify_trapdoor(taus: np.ndarray, commitment: np.ndarray, opening: np.ndarray, evaluation: np.ndarray,
                        proofs: np.ndarray) -> bool:
    """multilinear_kzg.rs:132-159 checked in G1 with the toxic waste known (no pairing)"""
    taus = _arr(taus).reshape(-1, 4)
    return bool(lib().zko_kzg_verify_trapdoor(_p(taus), C.c_uint32(taus.shape[0]), _p(_arr(commitment).reshape(12)),
                                              _p(_arr(opening).reshape(-1, 4)), _p(_arr(evaluation).reshape(4)),
                                              _p(_arr(proofs).reshape(-1, 12))))
