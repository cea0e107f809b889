"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/pbnn_b200.h declares (no compute
calls: there is no GPU here).  Also checks that the product has no CPU fallback."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cuda_lib_path():
    import __graft_entry__ as g
    return g.build_cuda()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pbnn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pbnn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(cuda_lib_path):
    from pbnn_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    lib = ctypes.CDLL(cuda_lib_path)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pbnn_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototypes out of sync with the header"
    bound = _lib.load_library(cuda_lib_path)
    assert bound.pbnn_version() == 200


def test_host_side_calls_without_gpu(cuda_lib_path):
    """Argument validation and planning are host code: they must work (and fail loudly) without a device."""
    from pbnn_b200 import _lib
    lib = _lib.load_library(cuda_lib_path)
    spec = _lib.make_spec((13, 50, 1), 0, 0.1)
    assert lib.pbnn_num_params(spec) == 751
    out = (ctypes.c_int32 * 4)()
    assert lib.pbnn_plan_query(spec, 506, 6, out) == 0
    assert out[0] % 32 == 0 and out[1] % 32 == 0 and out[2] <= 227 * 1024
    big = _lib.make_spec((90, 256, 256, 1), 0, 0.1)  # state goes to a global workspace, activations stay on chip
    assert lib.pbnn_plan_query(big, 128, 2, out) == 0
    # (relu, B <= 128: the two-hidden-layer tcgen05 engine adds one operand-image block per CTA; tanh: state only)
    assert lib.pbnn_workspace_bytes(big, _lib.OP_SGLD, 128, 10) > 10 * 2 * out[3] * 4 > 0
    big_tanh = _lib.make_spec((90, 256, 256, 1), 1, 0.1)
    assert lib.pbnn_workspace_bytes(big_tanh, _lib.OP_SGLD, 128, 10) == 10 * 2 * out[3] * 4
    # resident plans need no workspace; the tcgen05 HMC plan asks for hand-over buffers once chains outnumber the SMs
    assert lib.pbnn_workspace_bytes(spec, _lib.OP_HMC, 506, 100) == 0
    assert lib.pbnn_workspace_bytes(spec, _lib.OP_HMC, 506, 256) == 256 * (751 + 4) * 4 + 16
    assert lib.pbnn_workspace_bytes(spec, _lib.OP_SGLD, 32, 256) == 0
    assert lib.pbnn_plan_engine(spec, _lib.OP_HMC, 506, 256) == _lib.ENGINE_TCGEN05
    assert lib.pbnn_plan_engine(spec, _lib.OP_SGLD, 32, 256) == _lib.ENGINE_CTA4_H1   # few chains, relu: four warps per chain
    assert lib.pbnn_plan_engine(spec, _lib.OP_PSGLD, 32, 256) == _lib.ENGINE_FUSED_H1
    assert lib.pbnn_plan_engine(_lib.make_spec((13, 50, 1), 1, 0.1), _lib.OP_SGLD, 32, 256) == _lib.ENGINE_FUSED_H1  # tanh
    # many chains: one warp per chain (SG samplers and Adam; the gradient hook and HMC keep their engines)
    for op in (_lib.OP_SGLD, _lib.OP_PSGLD, _lib.OP_SGHMC, _lib.OP_ADAM):
        assert lib.pbnn_plan_engine(spec, op, 32, 2048) == _lib.ENGINE_WARP_H1
        assert lib.pbnn_workspace_bytes(spec, op, 32, 2048) == 0
    assert lib.pbnn_plan_engine(spec, _lib.OP_GRAD, 32, 2048) == _lib.ENGINE_FUSED_H1
    assert lib.pbnn_plan_engine(spec, _lib.OP_HMC, 506, 2048) == _lib.ENGINE_TCGEN05
    assert lib.pbnn_plan_engine(spec, _lib.OP_SGLD, 257, 2048) != _lib.ENGINE_WARP_H1  # rows no longer fit a warp's block
    assert lib.pbnn_plan_engine(_lib.make_spec((13, 65, 1), 0, 0.1), _lib.OP_SGLD, 32, 2048) != _lib.ENGINE_WARP_H1
    # Adam on small two-hidden-layer relu nets (BASELINE configs[3]): four warps per member; everything else as before
    ens = _lib.make_spec((8, 50, 50, 1), 0, 0.1)
    assert lib.pbnn_plan_engine(ens, _lib.OP_ADAM, 32, 512) == _lib.ENGINE_WARP_MLP2
    assert lib.pbnn_workspace_bytes(ens, _lib.OP_ADAM, 32, 512) == 0
    assert lib.pbnn_plan_engine(ens, _lib.OP_ADAM, 33, 512) == _lib.ENGINE_TILED           # more than 32 rows
    assert lib.pbnn_plan_engine(ens, _lib.OP_SGLD, 32, 512) == _lib.ENGINE_TILED           # samplers keep their engines
    assert lib.pbnn_plan_engine(_lib.make_spec((8, 50, 50, 1), 1, 0.1), _lib.OP_ADAM, 32, 512) == _lib.ENGINE_TILED  # tanh
    assert lib.pbnn_plan_engine(_lib.make_spec((8, 50, 40, 1), 0, 0.1), _lib.OP_ADAM, 32, 512) == _lib.ENGINE_TILED
    assert lib.pbnn_plan_engine(_lib.make_spec((8, 65, 65, 1), 0, 0.1), _lib.OP_ADAM, 32, 512) == _lib.ENGINE_TILED
    assert lib.pbnn_plan_engine(spec, _lib.OP_HMC, 4096, 4) == _lib.ENGINE_TCGEN05     # X streams through TMEM, 4 tiles at a time
    assert lib.pbnn_plan_engine(spec, _lib.OP_HMC, 4097, 4) != _lib.ENGINE_TCGEN05     # larger data sets: FP32 engines
    assert lib.pbnn_plan_engine(_lib.make_spec((13, 65, 1), 0, 0.1), _lib.OP_HMC, 506, 4) != _lib.ENGINE_TCGEN05
    assert lib.pbnn_plan_engine(_lib.make_spec((16, 50, 1), 0, 0.1), _lib.OP_HMC, 506, 4) != _lib.ENGINE_TCGEN05
    assert lib.pbnn_plan_engine(big, _lib.OP_SGLD, 128, 10) == _lib.ENGINE_TCGEN05_MLP2
    assert lib.pbnn_plan_engine(big_tanh, _lib.OP_SGLD, 128, 10) == _lib.ENGINE_TILED_GLOBAL_STATE
    assert lib.pbnn_plan_engine(big, _lib.OP_SGLD, 129, 10) == _lib.ENGINE_TILED_GLOBAL_STATE  # more than one 128-row tile
    assert lib.pbnn_plan_engine(_lib.make_spec((9, 100, 100, 1), 0, 0.1), _lib.OP_SGLD, 128, 1024) == _lib.ENGINE_TCGEN05_MLP2
    assert lib.pbnn_plan_engine(_lib.make_spec((9, 100, 100, 1), 1, 0.1), _lib.OP_SGLD, 128, 1024) == _lib.ENGINE_TILED
    assert lib.pbnn_plan_engine(_lib.make_spec((9, 100, 90, 1), 0, 0.1), _lib.OP_SGLD, 128, 1024) == _lib.ENGINE_TILED
    assert lib.pbnn_plan_engine(_lib.make_spec((13, 50, 1), 1, 0.1), _lib.OP_HMC, 506, 256) == _lib.ENGINE_FUSED_H1  # tanh
    huge = _lib.make_spec((3000, 4000, 1), 0, 0.1)
    assert lib.pbnn_plan_query(huge, 128, 2, out) == -2
    assert b"does not fit" in lib.pbnn_last_error()
    assert lib.pbnn_sgld_run(spec, None, None, 10, None, 0, 0, 4, None, None, None, 1, 0, 1, 1, 1, 0, None, None, None,
                             None, 1.0, None, 0, None) == -1
    bad = _lib.make_spec((13, 50, 1), 7, 0.1)
    assert lib.pbnn_num_params(bad) == 751 and lib.pbnn_plan_query(bad, 8, 2, out) == -1


def test_sass_is_sm100a(cuda_lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", cuda_lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_in_product():
    """The package must not import the oracle, and must refuse to run without CUDA."""
    import torch
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pbnn_b200")):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("for the oracle", "").replace("the oracle", ""), (dirpath, f)
    from pbnn_b200.ops import Ops
    if not torch.cuda.is_available():
        import numpy as np
        from pbnn_b200.ops import FlatSpec
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            Ops().normal(np.array([0, 1], np.uint32), 4)
