"""Smallest invocations of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py
The tcgen05 HMC kernel (13-50-1), the two-hidden-layer tcgen05 SG kernel with its bulk-copy ring, noise warps and fused
drains (9-100-100-1, B = 128), the fused one-hidden-layer FP32 kernel (8-50-1) and the tiled FP32 kernel (8-12-12-1),
the warp-per-chain SG kernel (8-50-1), the four-warp Adam kernel (8-50-50-1), plus the predictive kernels.  Where the
sanitizer is closed (this pool): `cuda-gdb -batch -ex "set cuda memcheck on" -ex run --args python scripts/sanitize_smoke.py`.  Results are checked for finiteness only -- parity is the test suite's job."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pbnn_b200.ops import FlatSpec, default_ops

ops = default_ops()
rng = np.random.default_rng(0)
which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["hmc", "tc2", "h1", "warp", "adam2", "tiled", "predict"]


def problem(layers, N, C):
    X = rng.standard_normal((N, layers[0])).astype(np.float32)
    y = rng.standard_normal((N, 1)).astype(np.float32)
    fs = FlatSpec(tuple(layers), 0, 0.1)
    keys = ops.split(np.array([[0, 7]], np.uint32), C)[0]
    th0 = 0.01 * ops.normal(keys, fs.num_params)
    return fs, X, y, keys, th0


if "hmc" in which:
    fs, X, y, keys, th0 = problem((13, 50, 1), 300, 3)
    r = ops.hmc(fs, X, y, th0, keys, 2, 1e-4, np.ones(fs.num_params, np.float32), 3)
    assert torch.isfinite(r["trace"]).all()
    print("hmc tcgen05 ok", flush=True)
if "tc2" in which:
    fs, X, y, keys, th0 = problem((9, 100, 100, 1), 400, 2)
    for fn in (lambda: ops.sgld(fs, X, y, th0, keys, 128, 1e-6, 2),
               lambda: ops.psgld(fs, X, y, th0, keys, 128, 1e-7, 2, 0.95),
               lambda: ops.sghmc(fs, X, y, th0, keys, 128, 1e-7, 1, 2, momentum_resampling="sigma_sqrt_step")):
        assert torch.isfinite(fn()["trace"]).all()
    print("sg tcgen05 (two hidden layers) ok", flush=True)
if "h1" in which:
    fs, X, y, keys, th0 = problem((8, 50, 1), 200, 3)
    assert torch.isfinite(ops.sgld(fs, X, y, th0, keys, 32, 1e-5, 3)["trace"]).all()
    print("sg fused fp32 ok", flush=True)
if "warp" in which:  # warp-per-chain engine (pbnn_warp.cuh), forced at a handful of chains
    ops.set_option("w1_min_chains", 1)
    fs, X, y, keys, th0 = problem((8, 50, 1), 200, 3)
    for fn in (lambda: ops.sgld(fs, X, y, th0, keys, 32, 1e-5, 3), lambda: ops.psgld(fs, X, y, th0, keys, 40, 1e-6, 2, 0.95),
               lambda: ops.sghmc(fs, X, y, th0, keys, 32, 1e-7, 1, 2, momentum_resampling="sigma_sqrt_step")):
        assert torch.isfinite(fn()["trace"]).all()
    ops.reset_options()
    print("sg warp-per-chain ok", flush=True)
if "adam2" in which:  # four-warp Adam kernel (pbnn_warp2.cuh)
    fs, X, y, keys, th0 = problem((8, 50, 50, 1), 200, 3)
    idx = torch.randint(0, 200, (3, 3, 32), dtype=torch.int32, device="cuda")
    assert torch.isfinite(ops.adam_ensemble(fs, X, y, th0, idx, 1e-3)["theta"]).all()
    print("adam four-warp ok", flush=True)
if "tiled" in which:
    fs, X, y, keys, th0 = problem((8, 12, 12, 1), 200, 3)
    assert torch.isfinite(ops.sgld(fs, X, y, th0, keys, 32, 1e-5, 3)["trace"]).all()
    idx = torch.randint(0, 200, (3, 2, 32), dtype=torch.int32, device="cuda")
    assert torch.isfinite(ops.adam_ensemble(fs, X, y, th0, idx, 1e-3)["theta"]).all()
    print("sg tiled fp32 + adam ok", flush=True)
if "predict" in which:
    for layers in ((13, 50, 1), (8, 12, 12, 1), (8, 50, 50, 1)):
        fs, X, y, keys, th0 = problem(layers, 200, 5)
        m = ops.predict(fs, th0, X, want_preds=True, want_moments=True)
        assert torch.isfinite(m["mean"]).all()
    print("predict ok", flush=True)
torch.cuda.synchronize()
print("sanitize_smoke: all done")
