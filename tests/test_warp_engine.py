"""Warp-per-chain engine (pbnn_b200/csrc/pbnn_warp.cuh): the default for the SG samplers / Adam on one-hidden-layer
nets once there are enough chains to fill the SMs with independent warps.  `tests/conftest.py` already sends every
parity test through it (fixture parameter `cuda_warp`: option w1_min_chains = 1); here: the default selection at a real
chain count, independence of a chain from the launch it runs in, shapes at the edges of the engine."""
import numpy as np
import pytest

from oracle import pbnn_oracle as o
from pbnn_b200 import _lib
from pbnn_b200.ops import FlatSpec

pytestmark = pytest.mark.gpu


def npy(t):
    return t.detach().cpu().numpy()


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def problem(layers, act, C, N=300, seed=5, scale=0.2):
    X, y = o.synthetic_regression(N, layers[0], seed)
    sp, fs = o.MlpSpec(tuple(layers), act, 0.1), FlatSpec(tuple(layers), act, 0.1)
    keys = o.split(o.PRNGKey(seed + 7), C)
    th0 = (scale * np.random.default_rng(seed).standard_normal((C, sp.num_params))).astype(np.float32)
    return X, y, sp, fs, keys, th0


def test_default_selection_and_chain_independence(cuda_ops, tc_options):
    """640 chains of the cfg1 net (8-50-1, B = 32, N = 1030): the plan picks the warp engine by itself; a sample of the
    chains matches the oracle; and a chain's trace is bit-identical when it runs alone (results do not depend on the
    warp / CTA / launch a chain lands in)."""
    C, T, B, eps = 640, 5, 32, 1e-5
    X, y, sp, fs, keys, th0 = problem((8, 50, 1), o.RELU, C, N=1030)
    assert cuda_ops.lib.pbnn_plan_engine(fs.c_struct(), _lib.OP_SGLD, B, C) == _lib.ENGINE_WARP_H1
    out = cuda_ops.sgld(fs, X, y, th0, keys, B, eps, T)
    assert not npy(out["flags"]).any()
    pick = np.array([0, 1, 31, 32, 333, C - 1])
    ref = o.sgld(sp, X, y, th0[pick], B, eps, T, keys[pick])
    assert rel(npy(out["trace"])[pick], ref) <= 1e-5
    tc_options(w1_min_chains=1)
    solo = cuda_ops.sgld(fs, X, y, th0[pick], keys[pick], B, eps, T)
    assert np.array_equal(npy(solo["trace"]), npy(out["trace"])[pick])
    # and the multi-warp fused path (what few chains get by default) agrees to rounding, not bit for bit
    tc_options(no_w1=1)
    fused = cuda_ops.sgld(fs, X, y, th0[pick], keys[pick], B, eps, T)
    assert rel(npy(fused["trace"]), npy(solo["trace"])) <= 1e-5
    assert not np.array_equal(npy(fused["trace"]), npy(solo["trace"]))


@pytest.mark.parametrize("layers,act,B", [((15, 64, 1), o.RELU, 96), ((1, 1, 1), o.TANH, 1), ((3, 7, 1), o.RELU, 33),
                                          ((12, 63, 1), o.TANH, 64), ((4, 50, 1), o.RELU, 256)])
def test_shapes_at_the_edges(cuda_ops, tc_options, layers, act, B):
    """Widest / narrowest nets of the engine, one float4 chunk to four per input row, ragged and multi-block batches."""
    tc_options(w1_min_chains=1)
    X, y, sp, fs, keys, th0 = problem(layers, act, 3, N=400)
    assert cuda_ops.lib.pbnn_plan_engine(fs.c_struct(), _lib.OP_SGLD, B, 3) == _lib.ENGINE_WARP_H1
    T, eps = 4, 1e-5
    ref = o.sgld(sp, X, y, th0, B, eps, T, keys)
    out = cuda_ops.sgld(fs, X, y, th0, keys, B, eps, T)
    assert rel(npy(out["trace"]), ref) <= 1e-5
    assert np.array_equal(npy(out["theta"]), npy(out["trace"])[:, -1])
    refp, (gref, Vref) = o.psgld(sp, X, y, th0, B, 1e-6, T, 0.95, keys, return_state=True)
    outp = cuda_ops.psgld(fs, X, y, th0, keys, B, 1e-6, T, 0.95)
    assert rel(npy(outp["trace"]), refp) <= 1e-5
    assert rel(npy(outp["square_avg"]), Vref) <= 1e-4


def test_thinning_burnin_and_repeat(cuda_ops, tc_options):
    tc_options(w1_min_chains=1)
    X, y, sp, fs, keys, th0 = problem((8, 50, 1), o.RELU, 4, N=1030, scale=0.05)
    full = cuda_ops.sgld(fs, X, y, th0, keys, 32, 1e-6, 12)
    assert np.isfinite(npy(full["trace"])).all() and not npy(full["flags"]).any()
    thin = cuda_ops.sgld(fs, X, y, th0, keys, 32, 1e-6, 12, thin=3, burnin=2)
    assert np.array_equal(npy(thin["trace"]), npy(full["trace"])[:, 2::3])
    again = cuda_ops.sgld(fs, X, y, th0, keys, 32, 1e-6, 12)
    assert np.array_equal(npy(again["trace"]), npy(full["trace"]))
    none = cuda_ops.sgld(fs, X, y, th0, keys, 32, 1e-6, 12, keep_trace=False)
    assert np.array_equal(npy(none["theta"]), npy(full["trace"])[:, -1])


def test_diverged_chain_raises_the_flag(cuda_ops, tc_options):
    """A chain that blows up (step size far past the stability limit) ends non-finite with bit 0 of its flag word set."""
    tc_options(w1_min_chains=1)
    X, y, sp, fs, keys, th0 = problem((8, 50, 1), o.RELU, 3, N=1030, scale=0.5)
    out = cuda_ops.sgld(fs, X, y, th0, keys, 32, 1e-2, 30)
    bad = ~np.isfinite(npy(out["theta"])).all(axis=1)
    assert bad.any()
    assert np.array_equal((npy(out["flags"]) & 1).astype(bool), bad)
