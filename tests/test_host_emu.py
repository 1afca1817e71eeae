"""Unit test of the device field arithmetic (csrc/fp.cuh) compiled for the host with the PTX carry
instructions emulated (-DZK_HOST_EMU), checked limb for limb against the C oracle: Montgomery
product, unreduced multiply-accumulate + wide REDC, fold-by-scalar (table + Barrett), 9-limb sums,
including 0 / 1 / p-1 / all-ones edge inputs.  CPU only."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def test_fp_cuh_host_emulation(tmp_path):
    exe = str(tmp_path / "emu")
    cmd = ["g++", "-O1", "-std=c++17", "-DZK_HOST_EMU", "-x", "c++",
           "-I", os.path.join(ROOT, "zk_cryptography_research_implementations_b200", "csrc"),
           os.path.join(ROOT, "tests", "host_emu", "emu_main.cpp"), os.path.join(ROOT, "oracle", "zkoracle.c"),
           "-o", exe]
    subprocess.check_call(cmd)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    assert out.stdout.count("ok") == 10, out.stdout[-2000:]


def test_fq381_and_g1_host_emulation(tmp_path):
    """csrc/fq381.cuh (12-limb Montgomery product of the BLS12-381 base field) and csrc/g1.cuh (XYZZ group law, inversion,
    affine conversion) compiled for the host, against the C oracle's 6x64-limb field and Jacobian arithmetic"""
    exe = str(tmp_path / "emu_g1")
    cmd = ["g++", "-O1", "-std=c++17", "-DZK_HOST_EMU", "-x", "c++",
           "-I", os.path.join(ROOT, "zk_cryptography_research_implementations_b200", "csrc"),
           os.path.join(ROOT, "tests", "host_emu", "emu_g1.cpp"), os.path.join(ROOT, "oracle", "zkoracle.c"),
           os.path.join(ROOT, "oracle", "zkoracle_kzg.c"), "-o", exe]
    subprocess.check_call(cmd)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    assert out.stdout.count("ok") == 3, out.stdout[-2000:]


def test_msm_reduction_replayed_on_the_cpu(tmp_path):
    """the multi-scalar multiplication's own plan, signed-digit decomposition (csrc/msm_plan.cuh) and host combine
    (csrc/msm_host.h) driven through buckets, chunk running sums and bit planes on the CPU, against sums of the oracle's scalar
    multiplications: window widths 2..16, the library's plans, the grouped (halving) layout of the batched opening rounds"""
    exe = str(tmp_path / "emu_msm")
    cmd = ["g++", "-O2", "-std=c++17", "-DZK_HOST_EMU", "-x", "c++", "-w",
           "-I", os.path.join(ROOT, "zk_cryptography_research_implementations_b200", "csrc"),
           os.path.join(ROOT, "tests", "host_emu", "emu_msm.cpp"), os.path.join(ROOT, "oracle", "zkoracle.c"),
           os.path.join(ROOT, "oracle", "zkoracle_kzg.c"), "-o", exe]
    subprocess.check_call(cmd)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    assert out.stdout.count("ok") == 3, out.stdout[-2000:]


def test_warp_keccak_lane_tables_match_their_generator():
    """the packed lane-routing words in csrc/dev_transcript.cuh are exactly what tools/gen_keccak_lanes.py derives (and
    checks against a textbook Keccak-f[1600] and hashlib's SHA3-256)"""
    import re
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_keccak_lanes.py")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    gen = {m.group(1): re.findall(r"0x[0-9a-f]{8}u", m.group(2)) for m in re.finditer(r"#define ZK_WK_([AB])_INIT \{(.*)\}", out.stdout)}
    src = open(os.path.join(ROOT, "zk_cryptography_research_implementations_b200", "csrc", "dev_transcript.cuh")).read()
    for name in "AB":
        body = src[src.index("#define ZK_WK_%s_INIT" % name):]
        body = body[:body.index("}")]
        assert re.findall(r"0x[0-9a-f]{8}u", body) == gen[name] and len(gen[name]) == 32
