"""In-tree build of libzkb200.so (hand-written sm_100a CUDA + the C-ABI host layer).

    python -m zk_cryptography_research_implementations_b200.build [--force]

nvcc cross-compiles without a GPU.  The translation units are compiled in parallel (the round
kernels are heavy: seven (P, D) shapes x three fields), then linked into one shared library next
to this file.  The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# ZKB200_VARIANT=name builds an experiment next to the product (objects in build_name/, library libzkb200_name.so);
# select it at run time with ZKB200_LIB=<path> (see _lib.py).  Used for A/B runs of kernel variants on the GPU box.
_VARIANT = os.environ.get("ZKB200_VARIANT", "")
OBJ = os.path.join(HERE, "build" + ("_" + _VARIANT if _VARIANT else ""))
LIB = os.path.join(HERE, "libzkb200" + ("_" + _VARIANT if _VARIANT else "") + ".so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
    # host side (transcript Keccak, per-round field arithmetic): BMI2 andn/rorx; every B200 host CPU is x86-64-v3 or newer
    "-Xcompiler", "-march=x86-64-v3",
]
# kernel-variant experiments: ZKB200_DEFINES="-DZK_PIPE_VARIANT=1" python -m ...build --force
NVCC_FLAGS += os.environ.get("ZKB200_DEFINES", "").split()


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_header_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    paths.append(os.path.join(HERE, "..", "include", "zk_sumcheck.h"))
    return max(os.path.getmtime(p) for p in paths)


def _compile(src: str, force: bool, hdr_mtime: float) -> str:
    obj = os.path.join(OBJ, src[:-3] + ".o")
    srcp = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(srcp), hdr_mtime):
        return obj
    log = obj[:-2] + ".ptxas.log"
    cmd = ["nvcc", *NVCC_FLAGS, "-c", srcp, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, (res.stdout + res.stderr)[-4000:]))
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    hdr = _newest_header_mtime()
    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, hdr), srcs))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = ["nvcc", "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs,
               "-ldl", "-lrt"]
        subprocess.check_call(cmd)
        if verbose:
            print("linked", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
