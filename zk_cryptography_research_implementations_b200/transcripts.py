"""`transcripts` crate mirror: transcripts/src/fiat_shamir/{interface,fiat_shamir_transcript}.rs.

Keccak-256 Fiat-Shamir transcript; stays on the host (tens of bytes per round)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .core import _ptr


class Transcript:
    """`Transcript` (fiat_shamir_transcript.rs:5-43)."""

    def __init__(self):                                   # Transcript::new :12-16
        self.lib = _lib.load()
        self.h = C.c_void_p(self.lib.zk_transcript_new())

    def append(self, incoming_data: bytes) -> None:       # :22-24
        self.lib.zk_transcript_append(self.h, bytes(incoming_data), len(incoming_data))

    def sample_random_challenge(self) -> bytes:           # :29-36
        out = C.create_string_buffer(32)
        self.lib.zk_transcript_sample(self.h, out)
        return out.raw

    def random_challenge_as_field_element(self, field: int) -> np.ndarray:   # :38-43
        out = np.zeros(4, dtype=np.uint64)
        self.lib.zk_transcript_challenge(self.h, field, _ptr(out))
        return out

    def __del__(self):
        try:
            if self.h:
                self.lib.zk_transcript_free(self.h)
                self.h = None
        except Exception:
            pass
