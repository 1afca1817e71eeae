"""Unit test of the device field arithmetic (csrc/fp.cuh) compiled for the host with the PTX carry
instructions emulated (-DZK_HOST_EMU), checked limb for limb against the C oracle: Montgomery
product, unreduced multiply-accumulate + wide REDC, fold-by-scalar (table + Barrett), 9-limb sums,
including 0 / 1 / p-1 / all-ones edge inputs.  CPU only."""
import os
import subprocess

import pytest

from conftest import ROOT


@pytest.mark.parametrize("defines", [[], ["-DZK_FOLD_COLS=1"]], ids=["product", "column-form fold (experiment)"])
def test_fp_cuh_host_emulation(tmp_path, defines):
    exe = str(tmp_path / "emu")
    cmd = ["g++", "-O1", "-std=c++17", "-DZK_HOST_EMU", *defines, "-x", "c++",
           "-I", os.path.join(ROOT, "zk_cryptography_research_implementations_b200", "csrc"),
           os.path.join(ROOT, "tests", "host_emu", "emu_main.cpp"), os.path.join(ROOT, "oracle", "zkoracle.c"),
           "-o", exe]
    subprocess.check_call(cmd)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    assert out.stdout.count("ok") == 10, out.stdout[-2000:]
