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
    # and the multi-warp fused path agrees to rounding, not bit for bit
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


# ------------------------------------------------------------------------------------------------------------------
# Adam for small two-hidden-layer relu nets, four warps per member (pbnn_b200/csrc/pbnn_warp2.cuh)
# ------------------------------------------------------------------------------------------------------------------
def _adam_reference(sp, X, y, th0, idx, lr, steps):
    th, m, v, cnt = th0.copy(), np.zeros_like(th0), np.zeros_like(th0), 0
    for s_ in range(steps):
        g = -o.grad_estimator(sp, th, X[idx[:, s_]], y[idx[:, s_]], X.shape[0])
        th, m, v, cnt = o.adam_step(th, g.astype(np.float32), m, v, cnt, lr)
    return th, m, v


@pytest.mark.parametrize("layers,B", [((8, 50, 50, 1), 32), ((15, 64, 64, 1), 32), ((3, 33, 33, 1), 17), ((1, 1, 1, 1), 1),
                                      ((12, 7, 7, 1), 5), ((4, 16, 16, 1), 32), ((9, 57, 57, 1), 31)])
def test_adam_two_hidden_layers_edge_shapes(cuda_ops, tc_options, layers, B):
    """Widest / narrowest / odd widths (the last lane's second unit does not exist), ragged batches, one to four float4
    chunks per input row: final parameters and both Adam moments against the numpy oracle; the tiled FP32 engine (option
    no_w2) agrees to rounding and is a different kernel."""
    C, N, steps, lr = 3, 200, 6, 1e-2
    X, y = o.synthetic_regression(N, layers[0], 4)
    sp, fs = o.MlpSpec(tuple(layers), o.RELU, 0.1), FlatSpec(tuple(layers), o.RELU, 0.1)
    th0 = (0.1 * np.random.default_rng(4).standard_normal((C, sp.num_params))).astype(np.float32)
    idx = np.random.default_rng(5).integers(0, N, (C, steps, B)).astype(np.int32)
    assert cuda_ops.lib.pbnn_plan_engine(fs.c_struct(), _lib.OP_ADAM, B, C) == _lib.ENGINE_WARP_MLP2
    th, m, v = _adam_reference(sp, X, y, th0, idx, lr, steps)
    out = cuda_ops.adam_ensemble(fs, X, y, th0, idx, lr)
    assert not npy(out["flags"]).any()
    assert rel(npy(out["theta"]), th) <= 5e-5  # six free-running Adam steps (the first moves every weight by ~lr)
    assert rel(npy(out["m"]), m) <= 5e-5 and rel(npy(out["v"]), v) <= 1e-4
    one = cuda_ops.adam_ensemble(fs, X, y, th0, idx[:, :1], lr)
    th1, _, _ = _adam_reference(sp, X, y, th0, idx, lr, 1)
    assert rel(npy(one["theta"]), th1) <= 1e-5
    tc_options(no_w2=1)
    assert cuda_ops.lib.pbnn_plan_engine(fs.c_struct(), _lib.OP_ADAM, B, C) != _lib.ENGINE_WARP_MLP2
    tiled = cuda_ops.adam_ensemble(fs, X, y, th0, idx, lr)
    assert rel(npy(tiled["theta"]), npy(out["theta"])) <= 5e-5


def test_adam_two_hidden_layers_members_are_independent_and_resumable(cuda_ops):
    """A member's result is bit-identical alone and among 300; two halves with the carried moments equal one run."""
    layers, B, N, C, steps = (8, 50, 50, 1), 32, 9568, 300, 8
    X, y = o.synthetic_regression(N, layers[0], 6)
    sp, fs = o.MlpSpec(layers, o.RELU, 0.1), FlatSpec(layers, o.RELU, 0.1)
    th0 = (0.05 * np.random.default_rng(6).standard_normal((C, sp.num_params))).astype(np.float32)
    idx = np.random.default_rng(7).integers(0, N, (C, steps, B)).astype(np.int32)
    out = cuda_ops.adam_ensemble(fs, X, y, th0, idx, 1e-2)
    pick = np.array([0, 1, 149, C - 1])
    solo = cuda_ops.adam_ensemble(fs, X, y, th0[pick], idx[pick], 1e-2)
    for k in ("theta", "m", "v"):
        assert np.array_equal(npy(solo[k]), npy(out[k])[pick])
    th, m, v = _adam_reference(sp, X, y, th0[pick], idx[pick], 1e-2, steps)
    assert rel(npy(solo["theta"]), th) <= 5e-5
    a = cuda_ops.adam_ensemble(fs, X, y, th0[pick], idx[pick][:, :3], 1e-2)
    b = cuda_ops.adam_ensemble(fs, X, y, a["theta"], idx[pick][:, 3:], 1e-2, m=a["m"], v=a["v"], count0=a["count"])
    assert np.array_equal(npy(b["theta"]), npy(solo["theta"]))
    bad = idx[pick].copy()
    bad[0, 2, 5], bad[1, 0, 0] = N + 3, -1
    c = cuda_ops.adam_ensemble(fs, X, y, th0[pick], bad, 1e-2)
    d = cuda_ops.adam_ensemble(fs, X, y, th0[pick], bad.clip(0, N - 1), 1e-2)
    assert np.array_equal(npy(c["theta"]), npy(d["theta"]))
    assert (npy(c["flags"])[:2] & 4).all() and not npy(d["flags"]).any()


@pytest.mark.parametrize("layers,B,C", [((8, 50, 1), 32, 1), ((13, 50, 1), 40, 5), ((15, 64, 1), 96, 3), ((3, 7, 1), 33, 2),
                                        ((1, 1, 1), 1, 1), ((4, 50, 1), 256, 2)])
def test_four_warps_per_chain_variant(cuda_ops, tc_options, layers, B, C):
    """Few chains, SGLD, relu: the warp engine with a CTA of four warps per chain (replicated weights, split forward /
    backward / noise).  Default selection, parity with the oracle (trace, thinning, per-step minibatches, control
    variates through the API tests), bit-identical repeats, and the fused path as the second opinion."""
    X, y, sp, fs, keys, th0 = problem(layers, o.RELU, C, N=400)
    assert cuda_ops.lib.pbnn_plan_engine(fs.c_struct(), _lib.OP_SGLD, B, C) == _lib.ENGINE_CTA4_H1
    T, eps = 6, 1e-6
    ref = o.sgld(sp, X, y, th0, B, eps, T, keys)
    out = cuda_ops.sgld(fs, X, y, th0, keys, B, eps, T)
    assert rel(npy(out["trace"]), ref) <= 1e-5 and not npy(out["flags"]).any()
    assert np.array_equal(npy(out["theta"]), npy(out["trace"])[:, -1])
    again = cuda_ops.sgld(fs, X, y, th0, keys, B, eps, T)
    assert np.array_equal(npy(again["trace"]), npy(out["trace"]))
    thin = cuda_ops.sgld(fs, X, y, th0, keys, B, eps, T, thin=2, burnin=1)
    assert np.array_equal(npy(thin["trace"]), npy(out["trace"])[:, 1::2])
    refp = o.sgld(sp, X, y, th0, B, eps, 4, keys, minibatch_mode="per_step")
    outp = cuda_ops.sgld(fs, X, y, th0, keys, B, eps, 4, minibatch_mode="per_step")
    assert rel(npy(outp["trace"]), refp) <= 1e-5
    tc_options(no_w1c=1)
    assert cuda_ops.lib.pbnn_plan_engine(fs.c_struct(), _lib.OP_SGLD, B, C) == _lib.ENGINE_FUSED_H1
    fused = cuda_ops.sgld(fs, X, y, th0, keys, B, eps, T)
    assert rel(npy(fused["trace"]), npy(out["trace"])) <= 1e-5
