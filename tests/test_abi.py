"""The C-ABI library: loads without a GPU, exports every symbol include/zk_sumcheck.h declares, its
host-side pieces (field helpers, Keccak transcript) match the oracle, and it refuses to run without
a GPU instead of falling back.  CPU only -- no kernel is launched here."""
import ctypes
import os
import random
import re

import numpy as np
import pytest

import pyoracle as po
from conftest import FIELDS, ROOT


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "zk_sumcheck.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(zk_[a-z0-9_]+)\s*\(", hdr)))


def test_library_loads_and_exports_every_declared_symbol(zk):
    from zk_cryptography_research_implementations_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    # and the Python binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == syms


def test_host_field_helpers_match_oracle(zk, co):
    rng = random.Random(3)
    for name, fid in FIELDS.items():
        p = po.P[name]
        vals = [0, 1, p - 1, (1 << 200) + 5] + [rng.randrange(p) for _ in range(40)]
        A = zk.fe_from_ints(fid, vals)
        assert np.array_equal(A, co.from_ints(fid, vals))
        assert zk.fe_to_ints(fid, A) == vals
        for i in range(0, len(vals) - 1):
            for op in ("add", "sub", "mul"):
                assert np.array_equal(zk.fe_binop(op, fid, A[i], A[i + 1]), co.fe_op(op, fid, A[i], A[i + 1]))
        assert zk.fe_to_ints(fid, zk.fe_from_ints(fid, [-1, p, p + 5])) == [p - 1, 0, 5]


def test_host_transcript_matches_oracle_and_appendix_b(zk, co, golden):
    from zk_cryptography_research_implementations_b200.transcripts import Transcript
    b = golden["appendix_b"]["transcript"]
    t = Transcript()
    t.append(b["append"].encode())
    assert t.sample_random_challenge().hex() == b["sample"]
    assert zk.fe_to_ints(0, t.random_challenge_as_field_element(0)) == [b["challenge"]]
    rng = random.Random(4)
    for fid in (0, 1, 2):
        a, o = Transcript(), co.Transcript()
        for step in range(40):
            n = rng.choice([0, 1, 31, 32, 96, 135, 136, 137, 500])
            data = bytes(rng.randrange(256) for _ in range(n))
            a.append(data)
            o.append(data)
            if step % 3 == 0:
                assert a.sample_random_challenge() == o.sample_random_challenge()
            else:
                assert np.array_equal(a.random_challenge_as_field_element(fid), o.random_challenge_as_field_element(fid))


def test_synthetic_table_generator_host_model(zk):
    # the host restatement of the device generator is deterministic, sharding-consistent and canonical
    full = zk.synthetic_table_ints(0, 0xB200, 1, 16)
    assert all(0 <= v < po.P["BN254_FQ"] for v in full)
    assert zk.synthetic_table_ints(0, 0xB200, 1, 8, first=1, step=2) == full[1::2]
    assert zk.synthetic_table_ints(0, 0xB200, 2, 16) != full


def test_no_cpu_fallback(zk):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(zk.ZkError):
        zk.Context(0, 0)


def test_product_path_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "zk_cryptography_research_implementations_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "zkoracle" not in src and "coracle" not in src and "pyoracle" not in src, f
