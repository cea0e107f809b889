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
                        pytest.param("cuda_generic", marks=pytest.mark.gpu),
                        pytest.param("cuda_warp", marks=pytest.mark.gpu)])
def ops(request):
    """Parity tests run against the emulator here (CPU) and against the CUDA library on the GPU box -- there three
    times: with the default kernel selection (tcgen05 engines where they apply, fused FP32 path for the other
    one-hidden-layer cases), with the tensor-core engines disabled (option "no_tc") and with the generic tiled-GEMM
    engine forced ("no_tc" + "no_h1"); a fourth pass ("w1_min_chains" = 1) sends every eligible SG / Adam call through the
    warp-per-chain engine, which by default is only chosen for many chains.  Engines are selected through the library's explicit option table
    (pbnn_set_option; the plan is rebuilt on every call); nothing is read from the environment."""
    if request.param in ("cuda_notc", "cuda_generic"):
        o = request.getfixturevalue("cuda_ops")
        o.reset_options()
        o.set_option("no_tc", 1)
        o.set_option("no_w1c", 1)  # (few-chain SGLD on relu nets: the fused path, not the four-warp variant of the warp engine)
        if request.param == "cuda_generic":
            o.set_option("no_h1", 1)
        yield o
        o.reset_options()
        return
    if request.param == "cuda_warp":
        o = request.getfixturevalue("cuda_ops")
        o.reset_options()
        o.set_option("w1_min_chains", 1)
        yield o
        o.reset_options()
        return
    o = request.getfixturevalue(request.param + "_ops")
    o.reset_options()
    yield o
    o.reset_options()


@pytest.fixture
def options(ops):
    """set_option with automatic reset: ``options(no_sched=1, seg_len=3)``."""
    def set_(**kw):
        for k, v in kw.items():
            ops.set_option(k, v)
    yield set_
    ops.reset_options()


@pytest.fixture
def tc_options(cuda_ops):
    """As ``options`` for tests that take the CUDA library directly."""
    cuda_ops.reset_options()

    def set_(**kw):
        for k, v in kw.items():
            cuda_ops.set_option(k, v)
    yield set_
    cuda_ops.reset_options()
