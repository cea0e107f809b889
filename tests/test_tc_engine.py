"""tcgen05 HMC engine (pbnn_b200/csrc/pbnn_tc.cuh): parity with the oracle through the C ABI, and plan selection."""
import numpy as np
import pytest

from oracle import pbnn_oracle as o
from pbnn_b200.ops import FlatSpec

pytestmark = pytest.mark.gpu


def npy(t):
    return t.detach().cpu().numpy()


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def problem(layers, N, C, seed=3, scale=0.2):
    X, y = o.synthetic_regression(N, layers[0], seed)
    sp, fs = o.MlpSpec(tuple(layers), o.RELU, 0.1), FlatSpec(tuple(layers), o.RELU, 0.1)
    th = (scale * np.random.default_rng(seed).standard_normal((C, sp.num_params))).astype(np.float32)
    return X, y, sp, fs, th


# one tile, an odd number of tiles, a full last tile, the widest / narrowest shapes of the engine, N = 2
SHAPES = [((13, 50, 1), 506), ((8, 50, 1), 300), ((5, 64, 1), 100), ((15, 33, 1), 640), ((3, 7, 1), 129),
          ((1, 1, 1), 2), ((2, 64, 1), 256), ((15, 64, 1), 768), ((8, 50, 1), 1030), ((6, 20, 1), 2500)]


@pytest.mark.parametrize("layers,N", SHAPES)
def test_tc_gradient_matches_float64(cuda_ops, tc_options, layers, N):
    """3xTF32 split operands: the gradient agrees with the float64 oracle to fp32-summation accuracy (the scalar
    FP32 engine reaches ~1e-7 on the same inputs, the tolerance here is 5e-6)."""
    X, y, sp, fs, th = problem(layers, N, 5)
    lp64, g64 = o.logposterior_full(sp, th.astype(np.float64), X.astype(np.float64), y.astype(np.float64))
    lp, g = cuda_ops.logposterior_grad(fs, X, y, th)
    assert rel(npy(g), g64) <= 5e-6
    assert rel(npy(lp), lp64) <= 5e-6
    tc_options(no_tc=1)
    lp1, g1 = cuda_ops.logposterior_grad(fs, X, y, th)
    assert rel(npy(g), npy(g1)) <= 5e-6


def test_tc_plan_is_selected_and_can_be_disabled(cuda_ops, tc_options):
    """The engines must actually differ: same inputs, not bit-identical results (else the switch does nothing)."""
    X, y, sp, fs, th = problem((13, 50, 1), 506, 3)
    _, g_tc = cuda_ops.logposterior_grad(fs, X, y, th)
    tc_options(no_tc=1)
    _, g_h1 = cuda_ops.logposterior_grad(fs, X, y, th)
    assert not np.array_equal(npy(g_tc), npy(g_h1))
    assert rel(npy(g_tc), npy(g_h1)) <= 5e-6


@pytest.mark.parametrize("layers,N", [((13, 50, 1), 506), ((8, 50, 1), 300), ((4, 20, 1), 64), ((8, 50, 1), 1030)])
def test_tc_hmc_trajectories(cuda_ops, layers, N):
    """Full HMC through the tensor-core engine: fused half-kicks / drift, L = 1 and L > 1, several iterations."""
    X, y, sp, fs, th0 = problem(layers, N, 4, scale=0.05)
    keys = o.split(o.PRNGKey(11), 4)
    minv = np.linspace(0.8, 1.25, sp.num_params).astype(np.float32)
    for L, S, eps in ((1, 3, 2e-4), (5, 4, 2e-4), (12, 2, 1e-4)):
        ref, info = o.hmc(sp, X, y, th0, S, eps, minv, L, keys, return_info=True)
        out = cuda_ops.hmc(fs, X, y, th0, keys, S, eps, minv, L, info=True)
        acc = npy(out["accept"]).astype(bool)
        agree = np.ones(4, bool)
        for s in range(S):
            agree &= acc[:, s] == info["accept"][:, s]
            assert rel(npy(out["trace"])[agree, s], ref[agree, s]) <= 2e-5, (L, s)
        assert agree.sum() >= 3
        assert np.allclose(npy(out["delta"])[agree], info["delta"][agree], atol=0.05)


def test_tc_hmc_rejections_and_resume(cuda_ops):
    """Large step: rejected proposals restore (theta, grad); a run split in two calls equals one run (t0 offsets)."""
    X, y, sp, fs, th0 = problem((6, 16, 1), 200, 6, scale=0.3)
    keys = o.split(o.PRNGKey(5), 6)
    minv = np.ones(sp.num_params, np.float32)
    S, L, eps = 8, 6, 6e-3
    ref, info = o.hmc(sp, X, y, th0, S, eps, minv, L, keys, return_info=True)
    assert (~info["accept"]).any() and info["accept"].any()
    out = cuda_ops.hmc(fs, X, y, th0, keys, S, eps, minv, L, info=True)
    acc = npy(out["accept"]).astype(bool)
    near = np.abs(info["u"] - np.minimum(1, np.exp(np.minimum(info["delta"], 0)))) < 1e-3
    agree = np.ones(6, bool)
    for s in range(S):
        same = acc[:, s] == info["accept"][:, s]
        assert np.all(same[agree] | near[agree, s])
        agree &= same
        assert rel(npy(out["trace"])[agree, s], ref[agree, s]) <= 1e-4
    assert agree.sum() >= 4
    a = cuda_ops.hmc(fs, X, y, th0, keys, 5, eps, minv, L)
    b = cuda_ops.hmc(fs, X, y, npy(a["theta"]), keys, 3, eps, minv, L, t0=5)
    full = npy(out["trace"])
    got = np.concatenate([npy(a["trace"]), npy(b["trace"])], 1)
    assert np.array_equal(got, full)


def test_tc_many_chains_two_waves(cuda_ops):
    """More chains than SMs (one CTA per SM: the second wave re-allocates TMEM) -- every chain must be finite and
    chains with the same key and start must be bit-identical wherever they are scheduled."""
    X, y, sp, fs, th = problem((13, 50, 1), 506, 1, scale=0.05)
    C = 333
    keys = np.repeat(o.split(o.PRNGKey(2), 1), C, 0)
    th0 = np.repeat(th, C, 0)
    out = cuda_ops.hmc(fs, X, y, th0, keys, 3, 2e-4, np.ones(sp.num_params, np.float32), 4)
    tr = npy(out["trace"])
    assert np.isfinite(tr).all()
    assert np.array_equal(tr, np.repeat(tr[:1], C, 0))


@pytest.mark.parametrize("seg_len", [1, 2, 3])
def test_tc_balanced_schedule_is_bit_identical(cuda_ops, tc_options, seg_len):
    """More chains than SMs: the persistent grid hands chains between CTAs through global memory; the arithmetic is
    the same, so traces, accept decisions and final states equal the one-CTA-per-chain run bit for bit."""
    X, y, sp, fs, th = problem((13, 50, 1), 506, 1, scale=0.05)
    C, S, L = 200, 7, 3
    keys = o.split(o.PRNGKey(21), C)
    th0 = (0.05 * np.random.default_rng(4).standard_normal((C, sp.num_params))).astype(np.float32)
    minv = np.ones(sp.num_params, np.float32)
    assert cuda_ops.lib.pbnn_workspace_bytes(fs.c_struct(), 5, 506, C) > 0
    tc_options(seg_len=seg_len)
    a = cuda_ops.hmc(fs, X, y, th0, keys, S, 3e-4, minv, L, info=True, thin=2, burnin=1)
    tc_options(no_sched=1)
    b = cuda_ops.hmc(fs, X, y, th0, keys, S, 3e-4, minv, L, info=True, thin=2, burnin=1)
    for k in ("trace", "theta", "logp", "accept_count", "delta", "accept", "flags"):
        assert np.array_equal(npy(a[k]), npy(b[k])), k
    assert not npy(a["flags"]).any()


def test_tc_repeated_runs_are_bit_identical(cuda_ops):
    """Dynamic item scheduling, two CTAs per SM, mbarrier hand-shakes: none of it may change a single bit between
    runs (a missed fence or barrier shows up here as run-to-run differences)."""
    X, y, sp, fs, th = problem((13, 50, 1), 506, 1, scale=0.05)
    C, S, L = 300, 6, 7
    keys = o.split(o.PRNGKey(33), C)
    th0 = (0.05 * np.random.default_rng(8).standard_normal((C, sp.num_params))).astype(np.float32)
    minv = np.ones(sp.num_params, np.float32)
    ref = None
    for rep in range(4):
        out = cuda_ops.hmc(fs, X, y, th0, keys, S, 3e-4, minv, L, info=True)
        got = {k: npy(out[k]).copy() for k in ("trace", "theta", "logp", "accept", "delta")}
        assert np.isfinite(got["trace"]).all()
        if ref is None:
            ref = got
        else:
            for k in ref:
                assert np.array_equal(got[k], ref[k]), (rep, k)


@pytest.mark.parametrize("layers,S,M", [((13, 50, 1), 37, 150), ((8, 50, 1), 70, 1024), ((3, 7, 1), 5, 129),
                                         ((15, 64, 1), 33, 700), ((1, 1, 1), 3, 1), ((6, 20, 1), 1, 2500)])
def test_tc_predict(cuda_ops, tc_options, layers, S, M):
    """Posterior predictive on the tcgen05 engine (pb_predict_tc_kernel): predictions and moments vs the float64
    oracle, and vs the FP32 kernel (option "no_tc") -- ragged tile counts, more / fewer samples than chunks."""
    sp, fs = o.MlpSpec(tuple(layers), o.RELU, 0.1), FlatSpec(tuple(layers), o.RELU, 0.1)
    rng = np.random.default_rng(5)
    th = (0.3 * rng.standard_normal((S, sp.num_params))).astype(np.float32)
    Xt = rng.standard_normal((M, layers[0])).astype(np.float32)
    F = o.predict(sp, th.astype(np.float64), Xt.astype(np.float64), dtype=np.float64)
    res = cuda_ops.predict(fs, th, Xt, want_preds=True, want_moments=True)
    assert rel(npy(res["preds"]), F) <= 5e-6
    assert rel(npy(res["mean"]), F.mean(0)) <= 5e-6
    assert np.abs(npy(res["var"]) - F.var(0)).max() <= 1e-5 * max(1.0, float((F ** 2).mean(0).max()))
    only_m = cuda_ops.predict(fs, th, Xt, want_preds=False, want_moments=True)
    assert np.array_equal(npy(only_m["mean"]), npy(res["mean"]))
    tc_options(no_tc=1)
    ref = cuda_ops.predict(fs, th, Xt, want_preds=True, want_moments=True)
    assert rel(npy(res["preds"]), npy(ref["preds"])) <= 5e-6
    assert not np.array_equal(npy(res["preds"]), npy(ref["preds"])) or S * M < 8  # the two kernels really differ
