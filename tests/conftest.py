import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def emu_ops():
    """Kernel-logic emulator (same kernel bodies compiled by g++; test infrastructure only)."""
    from emu_ops import EmuOps
    return EmuOps()


@pytest.fixture(scope="session")
def cuda_ops():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pbnn_b200.ops import default_ops
    return default_ops()


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu),
                        pytest.param("cuda_notc", marks=pytest.mark.gpu),
                        pytest.param("cuda_generic", marks=pytest.mark.gpu)])
def ops(request, monkeypatch):
    """Parity tests run against the emulator here (CPU) and against the CUDA library on the GPU box -- there three
    times: with the default kernel selection (tcgen05 engine for HMC on one-hidden-layer relu nets, fused FP32 path
    for the other one-hidden-layer cases), with the tensor-core engine disabled (PBNN_NO_TC=1) and with the generic
    tiled-GEMM engine forced (PBNN_NO_H1=1; the plan is rebuilt on every call, so the environment switch is enough)."""
    if request.param == "cuda_notc":
        monkeypatch.setenv("PBNN_NO_TC", "1")
        return request.getfixturevalue("cuda_ops")
    if request.param == "cuda_generic":
        monkeypatch.setenv("PBNN_NO_H1", "1")
        return request.getfixturevalue("cuda_ops")
    return request.getfixturevalue(request.param + "_ops")
