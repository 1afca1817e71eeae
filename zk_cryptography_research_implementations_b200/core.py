"""Context, field-element helpers and device tables over the C-ABI (include/zk_sumcheck.h).

Field elements cross every interface as numpy uint64 arrays whose last axis is 4: little-endian
limbs of the Montgomery form -- arkworks' in-memory layout, so a reference user's `Vec<F>` is the
same bytes.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import u64p, u8p, vp

BN254_FQ, BN254_FR, BLS12_381_FR = 0, 1, 2
FIELD_NAMES = {BN254_FQ: "BN254_FQ", BN254_FR: "BN254_FR", BLS12_381_FR: "BLS12_381_FR"}
MODULUS = {
    BN254_FQ: 21888242871839275222246405745257275088696311157297823662689037894645226208583,
    BN254_FR: 21888242871839275222246405745257275088548364400416034343698204186575808495617,
    BLS12_381_FR: 52435875175126190479447740508185965837690552500527637822603658699938581184513,
}


class ZkError(RuntimeError):
    pass


class ReferencePanic(AssertionError):
    """Raised where the reference would `panic!`; the message is the reference's panic text."""


def _ptr(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def as_elems(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.shape[-1] != 4:
        raise ValueError("field elements are (..., 4) uint64 limb arrays")
    return a


# ------------------------------------------------------------------ host-side element helpers
def fe_from_ints(field: int, vals: Iterable[int]) -> np.ndarray:
    """canonical Python ints (any sign / size, reduced mod p) -> Montgomery limbs (n, 4)"""
    lib = _lib.load()
    p = MODULUS[field]
    vals = [int(v) % p for v in vals]
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    tmp = np.zeros(4, dtype=np.uint64)
    for i, v in enumerate(vals):
        for k in range(4):
            tmp[k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
        lib.zk_fe_from_canonical(field, _ptr(tmp), _ptr(out[i]))
    return out


def fe_from_int(field: int, v: int) -> np.ndarray:
    return fe_from_ints(field, [v])[0]


def fe_to_ints(field: int, a) -> List[int]:
    lib = _lib.load()
    a = as_elems(a).reshape(-1, 4)
    tmp = np.zeros(4, dtype=np.uint64)
    res = []
    for i in range(a.shape[0]):
        lib.zk_fe_to_canonical(field, _ptr(a[i]), _ptr(tmp))
        res.append(sum(int(tmp[k]) << (64 * k) for k in range(4)))
    return res


def fe_binop(name: str, field: int, a, b) -> np.ndarray:
    out = np.zeros(4, dtype=np.uint64)
    getattr(_lib.load(), "zk_fe_" + name)(field, _ptr(as_elems(a)), _ptr(as_elems(b)), _ptr(out))
    return out


# ------------------------------------------------------------------ context
class Context:
    """One GPU, one stream, one field (zk_ctx).  Not thread-safe."""

    def __init__(self, field: int, device: int = 0, stream: Optional[int] = None):
        self.lib = _lib.load()
        self.field = field
        self.device = device
        h = vp()
        if stream is None:
            rc = self.lib.zk_ctx_create(C.byref(h), field, device)
        else:
            rc = self.lib.zk_ctx_create_on_stream(C.byref(h), field, device, vp(stream))
        if rc != 0 or not h:
            raise ZkError("zk_ctx_create failed (%d): no usable CUDA device? This library has no CPU fallback." % rc)
        self.h = h

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.zk_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int) -> None:
        if rc == 0:
            return
        msg = self.lib.zk_last_error(self.h).decode()
        if rc == _lib.ZK_ERR_ASSERT:
            raise ReferencePanic(msg)
        raise ZkError("%s (status %d)" % (msg, rc))

    def synchronize(self) -> None:
        self.check(self.lib.zk_ctx_synchronize(self.h))

    def set_profiling(self, on: bool) -> None:
        self.check(self.lib.zk_ctx_set_profiling(self.h, int(on)))

    def set_tail_log(self, tail_log: int) -> None:
        """Tables of at most 2**tail_log entries finish in one launch with the transcript on the device (0: never)."""
        self.check(self.lib.zk_ctx_set_tail_log(self.h, int(tail_log)))

    def tail_log(self) -> int:
        return int(self.lib.zk_ctx_get_tail_log(self.h))

    def reset_stats(self) -> None:
        self.check(self.lib.zk_ctx_reset_stats(self.h))

    def stats(self) -> dict:
        a, b = C.c_uint64(), C.c_uint64()
        ms, by = C.c_double(), C.c_double()
        self.check(self.lib.zk_ctx_get_stats(self.h, C.byref(a), C.byref(b), C.byref(ms), C.byref(by)))
        return {"launches": a.value, "round_launches": b.value, "round_ms": ms.value, "round_bytes": by.value}

    # ---- tables
    def upload(self, elems) -> "DeviceTable":
        elems = as_elems(elems).reshape(-1, 4)
        h = vp()
        self.check(self.lib.zk_table_upload(self.h, _ptr(elems), elems.shape[0], C.byref(h)))
        return DeviceTable(self, h)

    def generate(self, seed: int, table_id: int, n: int, first: int = 0, step: int = 1) -> "DeviceTable":
        h = vp()
        self.check(self.lib.zk_table_generate(self.h, seed, table_id, n, first, step, C.byref(h)))
        return DeviceTable(self, h)

    def wrap(self, device_ptr: int, n: int) -> "DeviceTable":
        h = vp()
        self.check(self.lib.zk_table_wrap(self.h, vp(device_ptr), n, C.byref(h)))
        return DeviceTable(self, h)


class DeviceTable:
    """A table of field elements resident in HBM (zk_table).  `h` raises once the table has been moved into a
    sumpoly / consumed by a prover / freed, so a stale Python object can never hand a dangling handle to the library."""

    def __init__(self, ctx: Context, handle, owned: bool = True):
        self.ctx = ctx
        self._h = handle
        self.owned = owned

    @property
    def h(self):
        if self._h is None:
            raise ReferencePanic("use of a moved table (consumed by a prover, given to a sumpoly, or freed)")
        return self._h

    @property
    def alive(self) -> bool:
        return self._h is not None

    def __len__(self) -> int:
        return int(self.ctx.lib.zk_table_len(self.h))

    @property
    def device_ptr(self) -> int:
        return int(self.ctx.lib.zk_table_device_ptr(self.h))

    def clone(self) -> "DeviceTable":
        h = vp()
        self.ctx.check(self.ctx.lib.zk_table_clone(self.ctx.h, self.h, C.byref(h)))
        return DeviceTable(self.ctx, h)

    def download(self) -> np.ndarray:
        out = np.zeros((len(self), 4), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.zk_table_download(self.ctx.h, self.h, _ptr(out)))
        return out

    def regenerate(self, seed: int, table_id: int, n: int, first: int = 0, step: int = 1) -> None:
        self.ctx.check(self.ctx.lib.zk_table_regenerate(self.ctx.h, self.h, seed, table_id, n, first, step))

    def release(self):
        """give up ownership: returns the raw handle (it now belongs to the caller, e.g. to a sumpoly) and kills this object"""
        h = self.h
        self.owned = False
        self._h = None
        return h

    def free(self) -> None:
        if self.owned and self._h and self.ctx.h:
            self.ctx.lib.zk_table_free(self.ctx.h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def synthetic_table_ints(field: int, seed: int, table_id: int, n: int, first: int = 0, step: int = 1) -> List[int]:
    """Host restatement of the device generator, as canonical ints -- used by tests to
    check `zk_table_generate` and to feed the oracle the same inputs."""
    M = (1 << 64) - 1
    G = 0x9E3779B97F4A7C15
    base = (seed ^ ((table_id * G) & M)) & M
    p = MODULUS[field]

    def sm(ctr: int) -> int:
        z = (base + (ctr + 1) * G) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)

    out = []
    for j in range(n):
        g = first + j * step
        v = 0
        for l in range(4):
            v |= sm(4 * g + l) << (64 * l)
        out.append(v % p)
    return out
