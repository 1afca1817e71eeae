import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

FIELDS = {"BN254_FQ": 0, "BN254_FR": 1, "BLS12_381_FR": 2}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def kzg_golden():
    """tests/golden/kzg_golden.json (made by tests/golden/make_golden_kzg.py from the Python big-int model)"""
    with open(os.path.join(ROOT, "tests", "golden", "kzg_golden.json")) as f:
        return json.load(f)


def golden_point(p):
    """golden affine point -> (x, y) ints or None"""
    return None if p is None else (int(p[0], 16), int(p[1], 16))


@pytest.fixture(scope="session")
def co():
    """the C oracle (test infrastructure)"""
    import coracle
    coracle.build()
    return coracle


@pytest.fixture(scope="session")
def zk():
    """the product package (libzkb200.so must be built; build it if the .so is missing)"""
    import zk_cryptography_research_implementations_b200 as pkg
    from zk_cryptography_research_implementations_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    _lib.load()
    return pkg


_ctx_cache = {}


@pytest.fixture
def ctx_for(zk):
    """ctx_for(field_id) -> Context on cuda:0 (cached per field)"""
    def get(fid):
        if fid not in _ctx_cache:
            _ctx_cache[fid] = zk.Context(fid, 0)
        return _ctx_cache[fid]
    return get
