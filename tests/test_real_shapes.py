"""Parity at the EXACT shapes of BASELINE.json configs 3-5 (the plans the benchmarked workloads really take: the
two-hidden-layer tcgen05 engine with its global workspace / operand images for 9-100-100-1 and 90-256-256-1 at B = 128,
the tiled FP32 engine for 8-50-50-1 at B = 32), plus the engine's own edge cases.

Tolerances: gradient norm-wise rel <= 1e-5 vs the float64 oracle (the bf16x3 tensor-core products land at ~2e-6, the
FP32 engines at ~2e-7); one teacher-forced sampler step rel <= 1e-5 on theta (north_star); pSGLD state 1e-4 (f32
cancellation in g at the final point, as in test_parity.py)."""
import numpy as np
import pytest

from oracle import pbnn_oracle as o
from pbnn_b200 import _lib
from pbnn_b200.ops import FlatSpec

pytestmark = pytest.mark.gpu

RTOL_STEP = 1e-5


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def npy(t):
    return t.detach().cpu().numpy()


def problem(layers, C, N, seed=1, scale=None):
    X, y = o.synthetic_regression(N, layers[0], seed)
    sp, fs = o.MlpSpec(tuple(layers), o.RELU, 0.1), FlatSpec(tuple(layers), 0, 0.1)
    keys = o.split(o.PRNGKey(seed + 100), C)
    if scale is None:
        th0 = o.init_positions(sp, keys)  # the Flax init law N(0, 0.01^2) of models.py:21-22
    else:
        th0 = (scale * np.random.default_rng(seed).standard_normal((C, sp.num_params))).astype(np.float32)
    return X, y, sp, fs, keys, th0


@pytest.fixture(params=["default", "no_tc2"])
def eng(request, cuda_ops):
    cuda_ops.reset_options()
    if request.param == "no_tc2":
        cuda_ops.set_option("no_tc2", 1)
    yield cuda_ops
    cuda_ops.reset_options()


CFG3 = ((9, 100, 100, 1), 128, 2000)
CFG4 = ((8, 50, 50, 1), 32, 9568)
CFG5 = ((90, 256, 256, 1), 128, 3000)


def test_engines_selected_for_the_baseline_shapes(cuda_ops):
    cuda_ops.reset_options()
    e = lambda layers, op, B, C: cuda_ops.lib.pbnn_plan_engine(FlatSpec(layers, 0, 0.1).c_struct(), op, B, C)
    for op in (_lib.OP_SGLD, _lib.OP_PSGLD, _lib.OP_SGHMC, _lib.OP_GRAD):
        assert e(CFG3[0], op, 128, 1024) == _lib.ENGINE_TCGEN05_MLP2
        assert e(CFG5[0], op, 128, 1024) == _lib.ENGINE_TCGEN05_MLP2
    assert e(CFG4[0], _lib.OP_ADAM, 32, 512) == _lib.ENGINE_WARP_MLP2  # too small for a 128-row MMA tile: one warp per member
    assert e(CFG4[0], _lib.OP_SGLD, 32, 512) == _lib.ENGINE_TILED
    cuda_ops.set_option("no_tc2", 1)
    assert e(CFG3[0], _lib.OP_SGLD, 128, 1024) == _lib.ENGINE_TILED
    assert e(CFG5[0], _lib.OP_SGLD, 128, 1024) == _lib.ENGINE_TILED_GLOBAL_STATE
    cuda_ops.reset_options()


@pytest.mark.parametrize("layers,B,N", [CFG3, CFG4, CFG5, ((70, 160, 160, 1), 77, 400), ((20, 64, 64, 1), 64, 300)])
@pytest.mark.parametrize("scale", [None, 0.1])
def test_gradient_at_real_shapes(eng, layers, B, N, scale):
    if scale is not None and layers[1] >= 256:
        scale = 0.03  # keep the float64 yard-stick itself well conditioned
    X, y, sp, fs, keys, th0 = problem(layers, 3, N, scale=scale)
    idx = np.random.default_rng(2).integers(0, N, (3, B)).astype(np.int32)
    g, ll = eng.grad_estimator(fs, X, y, th0, idx)
    g64 = o.grad_estimator(sp, th0.astype(np.float64), X[idx].astype(np.float64), y[idx].astype(np.float64), N)
    ll64, _ = o.loglik_and_grad(sp, th0.astype(np.float64), X[idx].astype(np.float64), y[idx].astype(np.float64), 1.0)
    assert rel(npy(g), g64) <= 1e-5
    assert rel(npy(ll), ll64) <= 1e-5
    # per Dense layer as well (a wrong small layer must not hide behind the big one)
    off = 0
    for i, oo in zip(layers[:-1], layers[1:]):
        n = oo * (i + 1)
        assert rel(npy(g)[:, off:off + n], g64[:, off:off + n]) <= 2e-5, (i, oo)
        off += n


@pytest.mark.parametrize("layers,B,N", [CFG3, CFG5])
def test_sgld_teacher_forced_at_real_shapes(eng, layers, B, N):
    X, y, sp, fs, keys, th0 = problem(layers, 2, N)
    T, eps = 4, 1e-6
    ref = o.sgld(sp, X, y, th0, B, eps, T, keys)
    out = eng.sgld(fs, X, y, th0, keys, B, eps, T)
    assert rel(npy(out["trace"]), ref) <= RTOL_STEP
    assert np.array_equal(npy(out["theta"]), npy(out["trace"])[:, -1])
    assert not npy(out["flags"]).any()
    for t in range(1, T):
        one = eng.sgld(fs, X, y, ref[:, t - 1], keys, B, eps, 1, t0=t)
        assert rel(npy(one["trace"])[:, 0], ref[:, t]) <= RTOL_STEP
    # thinning / burn-in / per-step minibatches / scheduled step sizes go through the same fused drains
    thin = eng.sgld(fs, X, y, th0, keys, B, eps, T, thin=2, burnin=1)
    assert np.array_equal(npy(thin["trace"]), npy(out["trace"])[:, 1::2])
    sched = [eps * (1 + t) for t in range(T)]  # scheduled_sgld (langevin.py:143-204): one oracle step per step size
    cur, rows = th0, []
    for t in range(T):
        cur = o.sgld(sp, X, y, cur, B, sched[t], 1, keys, t0=t)[:, 0]
        rows.append(cur)
    assert rel(npy(eng.sgld(fs, X, y, th0, keys, B, sched, T)["trace"]), np.stack(rows, 1)) <= RTOL_STEP
    ref_p = o.sgld(sp, X, y, th0, B, eps, T, keys, minibatch_mode="per_step")
    assert rel(npy(eng.sgld(fs, X, y, th0, keys, B, eps, T, minibatch_mode="per_step")["trace"]), ref_p) <= RTOL_STEP


def test_psgld_at_cfg3(eng):
    """pSGLD divides by sqrt(V) + 1e-8 with V0 = g0^2 (psgld.py:25-29, diffusions.py:42-59): elements whose gradient is
    near zero get a huge preconditioner, the injected noise sqrt(eps G) z is then 65 % of |theta| here and carries HALF the
    element-wise relative error of g.  The float32 oracle itself sits 3.9e-6 from the float64 oracle on this very problem
    (measured, same seeds); the bf16x3 tensor-core gradient (2e-6 norm-wise instead of 2e-7) lands at ~2e-5.  Stated
    tolerance for the preconditioned sampler: 5e-5 per trace (the FP32 engines stay below 1e-5)."""
    layers, B, N = CFG3
    X, y, sp, fs, keys, th0 = problem(layers, 2, N)
    T, eps, alpha = 4, 1e-7, 0.95
    tol = 5e-5 if eng.get_option("no_tc2") is None else RTOL_STEP
    ref, (gref, Vref) = o.psgld(sp, X, y, th0, B, eps, T, alpha, keys, return_state=True)
    out = eng.psgld(fs, X, y, th0, keys, B, eps, T, alpha)
    assert rel(npy(out["trace"]), ref) <= tol
    assert rel(npy(out["grad"]), gref) <= 1e-4 and rel(npy(out["square_avg"]), Vref) <= 1e-4
    ref2, st2 = o.psgld(sp, X, y, th0, B, eps, 2, alpha, keys, return_state=True)
    cont = eng.psgld(fs, X, y, ref2[:, -1], keys, B, eps, 2, alpha, t0=2, state=st2)
    assert rel(npy(cont["trace"]), ref[:, 2:]) <= tol
    # against the float64 oracle the engine must be no worse than a few times the float32 oracle's own distance
    ref64 = o.psgld(sp, X, y, th0, B, eps, T, alpha, keys, dtype=np.float64)
    assert rel(npy(out["trace"]), ref64) <= max(tol, 5 * rel(ref, ref64))


def test_sghmc_at_cfg3(eng):
    """SGHMC, L = 10 (pipelines/sgmcmc.py:128), momentum ~ N(0, step) at eps = 1e-6 (stable: |theta| stays < 0.1 over the
    20 inner steps; 1e-5 diverges): p ~ 1e-3 per element and eps * g ~ 1e-5 .. 1e-3, so the drift term is resolved."""
    layers, B, N = CFG3
    X, y, sp, fs, keys, th0 = problem(layers, 2, N)
    eps_h = 1e-6
    refh = o.sghmc(sp, X, y, th0, B, eps_h, 2, 10, keys, momentum=o.MOMENTUM_SIGMA_SQRT_STEP)
    outh = eng.sghmc(fs, X, y, th0, keys, B, eps_h, 2, 10, momentum_resampling="sigma_sqrt_step")
    assert rel(npy(outh["trace"]), refh) <= RTOL_STEP
    # ... and the increments themselves (theta_T - theta_0 is dominated by the momentum, so compare the drift part too)
    d_ref, d_out = refh[:, -1] - th0, npy(outh["trace"])[:, -1] - th0
    assert rel(d_out, d_ref) <= 1e-4
    # the default ("unit") momentum law of SURVEY.md 2b as well: theta moves by N(0, 1) per inner step, so only two inner
    # steps stay finite (ten give NaN in the oracle too -- the reason the law is a parameter)
    refu = o.sghmc(sp, X, y, th0, B, 1e-7, 1, 2, keys)
    outu = eng.sghmc(fs, X, y, th0, keys, B, 1e-7, 1, 2)
    assert np.isfinite(refu).all()
    assert rel(npy(outu["trace"]), refu) <= RTOL_STEP


def test_adam_epoch_at_cfg4(eng):
    """One epoch of permutation-derived minibatches (deep_ensembles.py:59-75) at N = 9568, B = 32."""
    layers, B, N = CFG4
    X, y, sp, fs, keys, th0 = problem(layers, 3, N)
    perm = npy(eng.permutation(keys, N))
    assert np.array_equal(perm, np.stack([o.permutation(k, N) for k in keys]))
    steps = 12  # the first 12 of the epoch's 299 steps against the numpy oracle
    idx = perm[:, : (N // B) * B].reshape(3, N // B, B)[:, :steps].astype(np.int32)
    th, m, v, cnt = th0.copy(), np.zeros_like(th0), np.zeros_like(th0), 0
    for s_ in range(steps):
        g = -o.grad_estimator(sp, th, X[idx[:, s_]], y[idx[:, s_]], N)
        th, m, v, cnt = o.adam_step(th, g.astype(np.float32), m, v, cnt, 1e-2)
    out = eng.adam_ensemble(fs, X, y, th0, idx, 1e-2)
    assert rel(npy(out["theta"]), th) <= 1e-4  # 12 free-running Adam steps
    one = eng.adam_ensemble(fs, X, y, th0, idx[:, :1], 1e-2)
    th1, _, _, _ = o.adam_step(th0, (-o.grad_estimator(sp, th0, X[idx[:, 0]], y[idx[:, 0]], N)).astype(np.float32),
                               np.zeros_like(th0), np.zeros_like(th0), 0, 1e-2)
    assert rel(npy(one["theta"]), th1) <= RTOL_STEP


def test_tc2_control_variates_and_determinism(cuda_ops):
    cuda_ops.reset_options()
    layers, B, N = CFG3
    X, y, sp, fs, keys, th0 = problem(layers, 4, N)
    centre = th0 + 0.001
    ref = o.sgld_cv(sp, X, y, th0, B, 1e-6, 3, centre, keys)
    idx = cuda_ops.minibatch_indices(keys, 2, B, N)
    _, g_full = cuda_ops.logposterior_grad(fs, X, y, centre)
    g_est, _ = cuda_ops.grad_estimator(fs, X, y, centre, idx)
    out = cuda_ops.sgld(fs, X, y, th0, keys, B, 1e-6, 3, idx=idx, cv=(g_full, g_est))
    assert rel(npy(out["trace"]), ref) <= RTOL_STEP
    # mbarrier hand-shakes, bulk copies, fused drains: not a bit may change between runs or with the chain count
    a = npy(cuda_ops.sgld(fs, X, y, th0, keys, B, 1e-6, 5)["trace"])
    for _ in range(3):
        assert np.array_equal(npy(cuda_ops.sgld(fs, X, y, th0, keys, B, 1e-6, 5)["trace"]), a)
    C = 300  # more chains than SMs: persistent CTAs loop over their chains
    keys2 = o.split(o.PRNGKey(7), C)
    th2 = np.tile(th0[:1], (C, 1))
    big = npy(cuda_ops.sgld(fs, X, y, th2, keys2, B, 1e-6, 3)["trace"])
    for c in (0, 147, 148, 299):
        solo = npy(cuda_ops.sgld(fs, X, y, th2[c:c + 1], keys2[c:c + 1], B, 1e-6, 3)["trace"])
        assert np.array_equal(big[c:c + 1], solo), c


def test_out_of_range_indices_are_clamped_and_flagged(eng):
    layers, B, N = (9, 100, 100, 1), 64, 200
    X, y, sp, fs, keys, th0 = problem(layers, 2, N)
    idx = np.random.default_rng(3).integers(0, N, (2, B)).astype(np.int32)
    bad = idx.copy()
    bad[0, 3], bad[1, 7] = N + 5, -2
    good = bad.clip(0, N - 1)
    a = eng.sgld(fs, X, y, th0, keys, B, 1e-6, 2, idx=bad)
    b = eng.sgld(fs, X, y, th0, keys, B, 1e-6, 2, idx=good)
    assert np.array_equal(npy(a["trace"]), npy(b["trace"]))
    assert (npy(a["flags"]) & 4).all() and not npy(b["flags"]).any()


@pytest.mark.parametrize("layers,S,M", [(CFG3[0], 37, 1000), (CFG5[0], 5, 300), ((20, 64, 64, 1), 300, 130),
                                        ((70, 160, 160, 1), 3, 64), (CFG3[0], 1, 513), (CFG4[0], 64, 700),
                                        ((5, 32, 32, 1), 9, 129), ((8, 33, 33, 1), 40, 200)])
def test_predict_two_hidden_layers_on_tensor_cores(cuda_ops, layers, S, M):
    """predict_fn + moments (langevin.py:75-76, metrics_prediction_intervals.py:188) for the nets pbnn's MLP produces:
    tcgen05 kernel (pbnn_tc2_pred.cuh) vs the float64 oracle and vs the FP32 kernel; ragged tiles, one sample, more
    sample chunks than samples.  Tolerance 1e-5 (predictions and their first two moments)."""
    cuda_ops.reset_options()
    rng = np.random.default_rng(5)
    sp, fs = o.MlpSpec(tuple(layers), o.RELU, 0.1), FlatSpec(tuple(layers), 0, 0.1)
    scale = 0.05 if layers[1] >= 160 else 0.1
    smp = (scale * rng.standard_normal((S, sp.num_params))).astype(np.float32)
    Xt = rng.standard_normal((M, layers[0])).astype(np.float32)
    ref = o.predict(sp, smp, Xt, dtype=np.float64).reshape(S, M)
    out = cuda_ops.predict(fs, smp, Xt, want_preds=True, want_moments=True)
    assert rel(npy(out["preds"]).reshape(S, M), ref) <= 1e-5
    assert rel(npy(out["sum"]).reshape(M), ref.sum(0)) <= 1e-5
    assert rel(npy(out["sumsq"]).reshape(M), (ref * ref).sum(0)) <= 1e-5
    only = cuda_ops.predict(fs, smp, Xt, want_preds=False, want_moments=True)  # moments alone: same partial sums
    assert np.array_equal(npy(only["sum"]), npy(out["sum"])) and np.array_equal(npy(only["sumsq"]), npy(out["sumsq"]))
    cuda_ops.set_option("no_tc2", 1)
    fp32 = cuda_ops.predict(fs, smp, Xt, want_preds=True, want_moments=True)
    cuda_ops.reset_options()
    assert rel(npy(fp32["preds"]).reshape(S, M), ref) <= 1e-5
    assert not np.array_equal(npy(fp32["preds"]), npy(out["preds"]))  # (two different engines did run)
