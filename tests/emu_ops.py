"""Test-only: bind the kernel-logic emulator (tests/emu/libpbnn_emu.so, host pointers) through the same
marshalling layer as the CUDA library.  Never imported by the pbnn_b200 package."""
import os
import subprocess

import torch

from pbnn_b200 import _lib
from pbnn_b200.ops import Ops

_HERE = os.path.dirname(os.path.abspath(__file__))
EMU_SRC = os.path.join(_HERE, "emu", "pbnn_emu.cpp")
EMU_LIB = os.path.join(_HERE, "emu", "libpbnn_emu.so")


def build_emu(force=False):
    deps = [EMU_SRC] + [os.path.join(_HERE, "..", "pbnn_b200", "csrc", f)
                        for f in ("pbnn_engine.cuh", "pbnn_core.h", "pbnn_rng.cuh", "pbnn_api.inc")]
    if force or not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-fopenmp", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                               "-o", EMU_LIB, EMU_SRC])
    return EMU_LIB


class EmuOps(Ops):
    def __init__(self):
        super().__init__(lib=_lib.load_library(build_emu()))

    def _device(self):
        return torch.device("cpu")

    def _stream(self):
        return None
