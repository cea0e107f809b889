"""Parity of the kernels with the oracle.

Every test takes the ``ops`` fixture: on the GPU box (``-m gpu``) it is the CUDA library called through the C ABI;
on the CPU-only box (``-m "not gpu"``) it is the kernel-logic emulator, i.e. the same kernel bodies compiled by g++
(checks indexing / key flow, not GPU arithmetic).

Tolerances (float32):
  * RNG integers (bits, split, fold_in, randint, permutation): exact
  * normals: abs <= 1e-6 in the bulk (|z| < 4), rel <= 5e-6 in the tails (log1p / polynomial evaluation order)
  * gradient: norm-wise rel <= 1e-5 vs the float64 oracle
  * one sampler step (teacher-forced from the oracle state): norm-wise rel <= 1e-5 on theta   [north_star]
  * free-running T=50 trajectory: rel <= 1e-3 (documented drift bound)
"""
import numpy as np
import pytest

from oracle import pbnn_oracle as o
from pbnn_b200.ops import FlatSpec

RTOL_STEP = 1e-5


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def npy(t):
    return t.detach().cpu().numpy()


def u32(t):
    return npy(t).view(np.uint32)


NETS = [
    ((8, 50, 1), o.RELU, 32),       # cfg1 shape
    ((13, 50, 1), o.TANH, 40),      # cfg2 shape, ragged batch
    ((9, 20, 20, 1), o.RELU, 200),  # two hidden layers, batch > one tile
    ((5, 7, 3), o.TANH, 16),        # out > 1 (generic last layer)
    ((4, 2), o.RELU, 8),            # single Dense layer
    ((6, 40, 36, 1), o.TANH, 100),  # wide enough for the 8x8 / 8x4 register-tile GEMMs (batch tile 128)
]


def problem(layers, act, C=3, N=300, seed=1, scale=0.2):
    X, y = o.synthetic_regression(N, layers[0], seed)
    if layers[-1] > 1:
        y = (np.repeat(y, layers[-1], 1) * np.arange(1, layers[-1] + 1, dtype=np.float32)).astype(np.float32)
    sp = o.MlpSpec(tuple(layers), act, 0.1)
    fs = FlatSpec(tuple(layers), act, 0.1)
    keys = o.split(o.PRNGKey(seed + 100), C)
    th0 = (scale * np.random.default_rng(seed).standard_normal((C, sp.num_params))).astype(np.float32)
    return X, y, sp, fs, keys, th0


# ------------------------------------------------------------------------------------------------ RNG
def test_rng_integers_exact(ops):
    keys = np.concatenate([o.split(o.PRNGKey(3), 4), np.array([[0, 0], [0xFFFFFFFF, 0xFFFFFFFF]], np.uint32)])
    n = 1000
    assert np.array_equal(u32(ops.bits(keys, n)), np.stack([o.random_bits(k, n) for k in keys]))
    assert np.array_equal(u32(ops.split(keys, 7)), np.stack([o.split(k, 7) for k in keys]))
    assert np.array_equal(u32(ops.fold_in(keys, 0xDEADBEEF)), np.stack([o.fold_in(k, 0xDEADBEEF) for k in keys]))
    # documented JAX values through the device path
    assert u32(ops.split(o.PRNGKey(0), 2))[0].tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]


def test_rng_floats(ops):
    keys = o.split(o.PRNGKey(11), 3)
    n = 20000
    z = npy(ops.normal(keys, n))
    zr = np.stack([o.normal(k, n) for k in keys])
    bulk = np.abs(zr) < 4
    assert np.abs(z - zr)[bulk].max() <= 1e-6
    assert (np.abs(z - zr) / np.maximum(np.abs(zr), 1))[~bulk].max(initial=0) <= 5e-6
    u = npy(ops.uniform(keys, n))
    assert np.array_equal(u, np.stack([o.uniform(k, n) for k in keys]))
    assert abs(float(npy(ops.normal(o.PRNGKey(0), 1))[0, 0]) - 1.6226422) < 1e-6
    assert abs(float(npy(ops.uniform(o.PRNGKey(0), 1))[0, 0]) - 0.947667) < 1e-6


@pytest.mark.parametrize("B,N", [(32, 1030), (128, 45730), (7, 3), (1, 1), (64, 515345)])
def test_minibatch_indices_exact(ops, B, N):
    keys = o.split(o.PRNGKey(5), 3)
    for draw in (1, 2, 5):
        got = npy(ops.minibatch_indices(keys, draw, B, N))
        want = np.stack([o.reference_draw(k, B, N, draw) for k in keys])
        assert np.array_equal(got, want)
        assert got.min() >= 0 and got.max() < N


@pytest.mark.parametrize("N", [1, 2, 10, 1030, 1626, 9568])
def test_permutation_exact(ops, N):
    keys = o.split(o.PRNGKey(9), 3)
    got = npy(ops.permutation(keys, N))
    want = np.stack([o.permutation(k, N) for k in keys])
    assert np.array_equal(got, want)
    assert np.array_equal(np.sort(got, axis=1), np.tile(np.arange(N), (3, 1)))


# ------------------------------------------------------------------------------------------------ gradient
@pytest.mark.parametrize("layers,act,B", NETS)
def test_grad_estimator(ops, layers, act, B):
    X, y, sp, fs, keys, th0 = problem(layers, act)
    N, C = X.shape[0], th0.shape[0]
    idx = np.random.default_rng(2).integers(0, N, (C, B)).astype(np.int32)
    g, ll = ops.grad_estimator(fs, X, y, th0, idx)
    g64 = o.grad_estimator(sp, th0.astype(np.float64), X[idx].astype(np.float64), y[idx].astype(np.float64), N)
    ll64, _ = o.loglik_and_grad(sp, th0.astype(np.float64), X[idx].astype(np.float64), y[idx].astype(np.float64), 1.0)
    assert rel(npy(g), g64) <= 1e-5
    assert rel(npy(ll), ll64) <= 1e-5


# ------------------------------------------------------------------------------------------------ SG samplers
@pytest.mark.parametrize("layers,act,B", NETS)
def test_sgld_teacher_forced_and_free(ops, layers, act, B):
    X, y, sp, fs, keys, th0 = problem(layers, act)
    T, eps = 6, 1e-5
    ref = o.sgld(sp, X, y, th0, B, eps, T, keys)
    # free-running (identical start): short horizon stays within the per-step tolerance times T
    out = ops.sgld(fs, X, y, th0, keys, B, eps, T)
    assert rel(npy(out["trace"]), ref) <= RTOL_STEP
    assert np.array_equal(npy(out["theta"]), npy(out["trace"])[:, -1])
    assert not npy(out["flags"]).any()
    # teacher-forced: restart every step from the oracle state (t0 selects keys[t])
    for t in range(1, T):
        one = ops.sgld(fs, X, y, ref[:, t - 1], keys, B, eps, 1, t0=t)
        assert rel(npy(one["trace"])[:, 0], ref[:, t]) <= RTOL_STEP


@pytest.mark.parametrize("layers,act,B", NETS[:3])
def test_psgld(ops, layers, act, B):
    X, y, sp, fs, keys, th0 = problem(layers, act)
    T, eps, alpha = 6, 1e-6, 0.95
    ref, (gref, Vref) = o.psgld(sp, X, y, th0, B, eps, T, alpha, keys, return_state=True)
    out = ops.psgld(fs, X, y, th0, keys, B, eps, T, alpha)
    assert rel(npy(out["trace"]), ref) <= RTOL_STEP
    assert rel(npy(out["grad"]), gref) <= 1e-4   # gradient at the final point (f32 cancellation)
    assert rel(npy(out["square_avg"]), Vref) <= 1e-4
    # resume from the oracle state (exact restart: t0 + state)
    ref2, st2 = o.psgld(sp, X, y, th0, B, eps, 3, alpha, keys, return_state=True)
    cont = ops.psgld(fs, X, y, ref2[:, -1], keys, B, eps, 3, alpha, t0=3, state=st2)
    assert rel(npy(cont["trace"]), ref[:, 3:]) <= RTOL_STEP


@pytest.mark.parametrize("layers,act,B", NETS[:3])
def test_sghmc(ops, layers, act, B):
    X, y, sp, fs, keys, th0 = problem(layers, act, scale=0.05)
    T, eps, L = 3, 1e-9, 3  # unit-variance momentum is added to the position: keep the horizon short
    ref = o.sghmc(sp, X, y, th0, B, eps, T, L, keys)
    out = ops.sghmc(fs, X, y, th0, keys, B, eps, T, L)
    assert rel(npy(out["trace"]), ref) <= RTOL_STEP
    for t in range(1, T):
        one = ops.sghmc(fs, X, y, ref[:, t - 1], keys, B, eps, 1, L, t0=t)
        assert rel(npy(one["trace"])[:, 0], ref[:, t]) <= RTOL_STEP


@pytest.mark.parametrize("momentum", ["sigma_sqrt_step", "mu_sqrt_step", (0.25, 0.5, 0.125, 2.0)])
@pytest.mark.parametrize("layers,act,B", NETS[:3])
def test_sghmc_momentum_variants_resolve_the_drift(ops, layers, act, B, momentum):
    """The momentum resampling of blackjax.sghmc is a parameter (the upstream body is not vendored with the reference:
    src/pbnn/mcmc/hamiltonian.py:118-128 is only the call site).  With p ~ N(0, step) the position moves by O(sqrt(eps))
    per inner step, so the test can run at eps = 1e-5 -- where eps * g is ~1e-2 of the increment and a wrong sign or scale
    of the drift shows up far above the 1e-5 tolerance (the unit-variance test above needs eps = 1e-9 and is nearly blind
    to it)."""
    X, y, sp, fs, keys, th0 = problem(layers, act, scale=0.05)
    coef = {"sigma_sqrt_step": o.MOMENTUM_SIGMA_SQRT_STEP, "mu_sqrt_step": o.MOMENTUM_MU_SQRT_STEP}.get(momentum, momentum)
    eps = 1e-5 if momentum == "sigma_sqrt_step" else 1e-9
    T, L = 3, 4
    ref = o.sghmc(sp, X, y, th0, B, eps, T, L, keys, momentum=coef)
    out = ops.sghmc(fs, X, y, th0, keys, B, eps, T, L, momentum_resampling=momentum)
    assert rel(npy(out["trace"]), ref) <= RTOL_STEP
    if momentum == "sigma_sqrt_step":
        # the increments (drift + momentum + noise), and the drift is really in there: dropping eps * g moves them > 1e-3
        inc_ref, inc_out = ref[:, -1] - th0, npy(out["trace"])[:, -1] - th0
        assert rel(inc_out, inc_ref) <= 1e-4
        nodrift = o.sghmc(sp, X, y, th0, B, eps, T, L, keys, momentum=coef, _drop_drift=True)
        assert rel(nodrift[:, -1] - th0, inc_ref) >= 1e-3
    for t in range(1, T):
        one = ops.sghmc(fs, X, y, ref[:, t - 1], keys, B, eps, 1, L, t0=t, momentum_resampling=momentum)
        assert rel(npy(one["trace"])[:, 0], ref[:, t]) <= RTOL_STEP


def test_options_api(ops):
    """The engine switches are an explicit option table (pbnn_set_option), not environment variables."""
    ops.reset_options()  # (the engine-forcing fixture variants arrive with options set)
    assert ops.get_option("no_tc") is None
    ops.set_option("no_tc", 1)
    assert ops.get_option("no_tc") == 1
    ops.reset_options()
    assert ops.get_option("no_tc") is None
    with pytest.raises(ValueError):
        ops.set_option("no_such_option", 1)
    import os
    os.environ["PBNN_NO_TC"] = "1"  # has no effect any more
    try:
        assert ops.get_option("no_tc") is None
    finally:
        os.environ.pop("PBNN_NO_TC")


def test_out_of_range_minibatch_indices_are_clamped(ops):
    """A JAX gather clamps; an out-of-range index must not become an out-of-bounds read (flags bit 2 reports it)."""
    layers, act, B = NETS[0]
    X, y, sp, fs, keys, th0 = problem(layers, act)
    N = X.shape[0]
    idx = np.random.default_rng(3).integers(0, N, (3, B)).astype(np.int32)
    bad = idx.copy()
    bad[0, 3], bad[2, 7] = N + 5, -2
    a = ops.sgld(fs, X, y, th0, keys, B, 1e-5, 2, idx=bad)
    b = ops.sgld(fs, X, y, th0, keys, B, 1e-5, 2, idx=bad.clip(0, N - 1))
    assert np.array_equal(npy(a["trace"]), npy(b["trace"]))
    assert (npy(a["flags"]) & 4).tolist() == [4, 0, 4] and not npy(b["flags"]).any()


def test_sgld_options(ops):
    """thin / burnin / scheduled step sizes / per-step minibatches / explicit idx."""
    layers, act, B = NETS[0]
    X, y, sp, fs, keys, th0 = problem(layers, act)
    N, C, T = X.shape[0], th0.shape[0], 12
    sched = lambda t: 1e-5 / t
    steps = np.array([sched(t) for t in range(1, T + 1)], np.float32)
    # oracle with a per-step step size: run step by step
    th, ref = th0.copy(), []
    for t in range(T):
        th = o.sgld(sp, X, y, th, B, steps[t], 1, keys, t0=t)[:, 0]
        ref.append(th)
    ref = np.stack(ref, 1)
    out = ops.sgld(fs, X, y, th0, keys, B, sched, T, thin=3, burnin=2)
    assert npy(out["trace"]).shape == (C, 4, sp.num_params)
    assert rel(npy(out["trace"]), ref[:, 2::3]) <= RTOL_STEP
    # per-step mode == oracle per_step mode; and == explicit per-step idx
    refp = o.sgld(sp, X, y, th0, B, 1e-5, 5, keys, minibatch_mode="per_step")
    outp = ops.sgld(fs, X, y, th0, keys, B, 1e-5, 5, minibatch_mode="per_step")
    assert rel(npy(outp["trace"]), refp) <= RTOL_STEP
    idx = np.stack([[o.reference_draw(k, B, N, t + 2) for t in range(5)] for k in keys]).astype(np.int32)
    oute = ops.sgld(fs, X, y, th0, keys, B, 1e-5, 5, idx=idx)
    assert rel(npy(oute["trace"]), refp) <= RTOL_STEP
    # keep_trace=False keeps only the final state
    outn = ops.sgld(fs, X, y, th0, keys, B, 1e-5, 5, keep_trace=False)
    assert outn["trace"] is None
    ref5 = o.sgld(sp, X, y, th0, B, 1e-5, 5, keys)
    assert rel(npy(outn["theta"]), ref5[:, -1]) <= RTOL_STEP


def test_free_running_drift(ops):
    layers, act, B = NETS[0]
    X, y, sp, fs, keys, th0 = problem(layers, act, scale=0.01)
    T = 50
    ref = o.sgld(sp, X, y, th0, B, 1e-5, T, keys)
    out = npy(ops.sgld(fs, X, y, th0, keys, B, 1e-5, T)["trace"])
    drift = [rel(out[:, t], ref[:, t]) for t in range(T)]
    assert max(drift) <= 1e-3


# ------------------------------------------------------------------------------------------------ HMC
@pytest.mark.parametrize("layers,act,N", [((13, 50, 1), o.RELU, 506), ((4, 6, 6, 1), o.TANH, 70)])
def test_hmc(ops, layers, act, N):
    X, y, sp, fs, keys, th0 = problem(layers, act, N=N, scale=0.05)
    S, L, eps = 4, 5, 2e-4
    minv = np.ones(sp.num_params, np.float32)
    ref, info = o.hmc(sp, X, y, th0, S, eps, minv, L, keys, return_info=True)
    out = ops.hmc(fs, X, y, th0, keys, S, eps, minv, L, info=True)
    # energies are O(|logp|): float32 spacing at that magnitude bounds the agreement of delta
    lp64, _ = o.logposterior_full(sp, th0.astype(np.float64), X.astype(np.float64), y.astype(np.float64))
    tol = 64 * np.spacing(np.float32(np.abs(lp64).max()))
    d, dr = npy(out["delta"]), info["delta"]
    pacc = np.minimum(1.0, np.exp(np.minimum(dr.astype(np.float64), 0)))
    decisive = np.abs(info["u"] - pacc) > 10 * tol * np.maximum(pacc, 1e-6) + 1e-5
    # compare step by step while the accept decisions agree (a flipped near-tie forks the trajectory)
    agree = np.ones(th0.shape[0], bool)
    for s in range(S):
        a = npy(out["accept"])[:, s].astype(bool)
        assert np.all(np.abs(d[agree, s] - dr[agree, s]) <= tol), (s, d[:, s], dr[:, s], tol)
        same = (a == info["accept"][:, s])
        assert np.all(same[agree] | ~decisive[agree, s])
        agree &= same
        assert rel(npy(out["trace"])[agree, s], ref[agree, s]) <= RTOL_STEP
    assert agree.any()
    assert np.array_equal(npy(out["accept_count"]), npy(out["accept"]).sum(1))


def test_hmc_nonunit_mass_and_reject(ops):
    layers, act = (3, 4, 1), o.TANH
    X, y, sp, fs, keys, th0 = problem(layers, act, N=40, C=6, scale=0.3)
    minv = np.linspace(0.5, 2.0, sp.num_params).astype(np.float32)
    S, L, eps = 6, 8, 2e-2  # large step: some rejections
    ref, info = o.hmc(sp, X, y, th0, S, eps, minv, L, keys, return_info=True)
    out = ops.hmc(fs, X, y, th0, keys, S, eps, minv, L, info=True)
    acc = npy(out["accept"]).astype(bool)
    assert (~info["accept"]).any(), "test should exercise a rejection"
    near_tie = np.abs(info["u"] - np.minimum(1, np.exp(np.minimum(info["delta"], 0)))) < 1e-3
    agree = np.ones(6, bool)
    for s in range(S):
        same = acc[:, s] == info["accept"][:, s]
        assert np.all(same[agree] | near_tie[agree, s])
        agree &= same
        assert rel(npy(out["trace"])[agree, s], ref[agree, s]) <= 5e-5
    assert agree.sum() >= 4


# ------------------------------------------------------------------------------------------------ predictive
@pytest.mark.parametrize("layers,act,B", NETS)
def test_predict_and_moments(ops, layers, act, B):
    X, y, sp, fs, keys, th0 = problem(layers, act, C=37)
    Xt = o.synthetic_regression(150, layers[0], 7)[0]
    res = ops.predict(fs, th0, Xt, want_preds=True, want_moments=True)
    F = o.predict(sp, th0.astype(np.float64), Xt.astype(np.float64), dtype=np.float64)
    assert rel(npy(res["preds"]), F) <= 1e-5
    assert rel(npy(res["mean"]), F.mean(0)) <= 1e-5
    assert np.abs(npy(res["var"]) - F.var(0)).max() <= 1e-5 * max(1.0, float((F ** 2).mean(0).max()))


# ------------------------------------------------------------------------------------------------ ensembles
def test_adam_ensemble(ops):
    layers, act = (8, 12, 12, 1), o.RELU
    X, y, sp, fs, keys, th0 = problem(layers, act, C=4, N=120, scale=0.01)
    N, B, E = X.shape[0], 16, 1
    mk = [k for k, _ in o.ensemble_member_keys(o.PRNGKey(3), 4)]
    ref = o.deep_ensembles(sp, X, y, th0, B, E, 1e-3, mk)
    idx = np.stack([np.concatenate([o.ensemble_epoch_indices(e, N, B) for e in o.split(k, E + 1)]) for k in mk])
    out = ops.adam_ensemble(fs, X, y, th0, idx.astype(np.int32), 1e-3)
    assert rel(npy(out["theta"]), ref) <= 2e-5
    # resume: two halves == one run
    h = idx.shape[1] // 2
    a = ops.adam_ensemble(fs, X, y, th0, idx[:, :h].astype(np.int32), 1e-3)
    b = ops.adam_ensemble(fs, X, y, a["theta"], idx[:, h:].astype(np.int32), 1e-3, m=a["m"], v=a["v"], count0=a["count"])
    assert rel(npy(b["theta"]), npy(out["theta"])) <= 1e-6


@pytest.mark.parametrize("n,D,size", [(200, 37, 20), (1000, 501, 50), (65, 3, 65)])
def test_thinning_matches_reference_order(ops, n, D, size):
    """thinning_fn (utils/misc.py:58-125): same index sequence as the float64 restatement on generic data."""
    rng = np.random.default_rng(n)
    pos = (rng.standard_normal((n, D)) * (1 + rng.random((n, 1)))).astype(np.float32)
    got = npy(ops.thinning(pos, size))
    want = o.thinning(pos, size, dtype=np.float64)
    assert got.shape == (size,) and len(set(got.tolist())) <= size
    assert np.array_equal(got, want)


def test_heteroscedastic_likelihood(ops):
    """heteroscedastic_loglike_fn (benchmark/pipelines/logprobs.py:37-42): outputs (mu, log sigma^2), one target."""
    from pbnn_b200._lib import LIK_HETEROSCEDASTIC
    layers = (3, 10, 2)
    X, y = o.synthetic_regression(80, 3, 4)
    sp = o.MlpSpec(layers, o.TANH, 1.0, heteroscedastic=True)
    fs = FlatSpec(layers, o.TANH, 1.0, 1.0, LIK_HETEROSCEDASTIC)
    keys = o.split(o.PRNGKey(8), 3)
    th0 = (0.3 * np.random.default_rng(3).standard_normal((3, sp.num_params))).astype(np.float32)
    idx = np.random.default_rng(2).integers(0, 80, (3, 24)).astype(np.int32)
    g, ll = ops.grad_estimator(fs, X, y, th0, idx)
    g64 = o.grad_estimator(sp, th0.astype(np.float64), X[idx].astype(np.float64), y[idx].astype(np.float64), 80)
    ll64, _ = o.loglik_and_grad(sp, th0.astype(np.float64), X[idx].astype(np.float64), y[idx].astype(np.float64), 1.0)
    assert rel(npy(g), g64) <= 1e-5 and rel(npy(ll), ll64) <= 1e-5
    # finite-difference check of the oracle's heteroscedastic gradient
    t = th0[:1].astype(np.float64)
    f = lambda v: (o.logprior(sp, v) + 80 * o.loglik_and_grad(sp, v, X[idx[0]].astype(np.float64), y[idx[0]].astype(np.float64), 1.0)[0] / 24)[0]
    fd = np.array([(f(t + 1e-6 * np.eye(sp.num_params)[j]) - f(t - 1e-6 * np.eye(sp.num_params)[j])) / 2e-6
                   for j in range(sp.num_params)])
    assert rel(g64[0], fd) <= 1e-6
    ref = o.sgld(sp, X, y, th0, 24, 1e-4, 5, keys)
    assert rel(npy(ops.sgld(fs, X, y, th0, keys, 24, 1e-4, 5)["trace"]), ref) <= RTOL_STEP
    minv = np.ones(sp.num_params, np.float32)
    ref, info = o.hmc(sp, X, y, th0, 3, 1e-3, minv, 4, keys, return_info=True)
    out = ops.hmc(fs, X, y, th0, keys, 3, 1e-3, minv, 4, info=True)
    assert np.array_equal(npy(out["accept"]).astype(bool), info["accept"])
    assert rel(npy(out["trace"]), ref) <= RTOL_STEP
    with pytest.raises(ValueError):
        ops.sgld(FlatSpec((3, 10, 1), o.TANH, 1.0, 1.0, LIK_HETEROSCEDASTIC), X, y, th0[:, :51], keys, 24, 1e-4, 2)


# ------------------------------------------------------------------------------------------------ edges / errors
def test_edge_cases_and_errors(ops):
    layers, act, B = NETS[0]
    X, y, sp, fs, keys, th0 = problem(layers, act)
    # zero steps: state unchanged, empty trace
    out = ops.sgld(fs, X, y, th0, keys, B, 1e-5, 0)
    assert npy(out["trace"]).shape[1] == 0 and np.array_equal(npy(out["theta"]), th0)
    # a single datum / batch of one
    out = ops.sgld(fs, X[:1], y[:1], th0, keys, 1, 1e-6, 2)
    ref = o.sgld(sp, X[:1], y[:1], th0, 1, 1e-6, 2, keys)
    assert rel(npy(out["trace"]), ref) <= RTOL_STEP
    # divergence is flagged, not hidden (pipelines/sgmcmc.py:207-209 checks NaNs post hoc)
    out = ops.sgld(fs, X, y, th0 * 50, keys, B, 10.0, 30)
    assert npy(out["flags"]).all()
    with pytest.raises(ValueError):
        ops.sgld(fs, X, y, th0[:, :-1], keys, B, 1e-5, 2)
    with pytest.raises(ValueError):
        ops.sgld(fs, X, y, th0, keys, B, 1e-5, 2, thin=0)
    with pytest.raises(ValueError):
        ops.sghmc(fs, X, y, th0, keys, B, 1e-5, 2, 0)
    with pytest.raises(ValueError):
        ops.sgld(FlatSpec((8, 50, 1), 0, -1.0), X, y, th0, keys, B, 1e-5, 2)
    huge = FlatSpec((3000, 4000, 1), 0, 0.1)
    with pytest.raises(MemoryError):
        ops.plan(huge, 128, 2)


def test_streaming_tiles_path(ops, options):
    """Row sets that do not fit shared memory are streamed tile by tile from global memory (large-N HMC, large
    minibatches); force that path and check it like the resident one."""
    options(no_resident=1, no_h1=1, no_tc=1)
    layers, act, B = NETS[2]
    X, y, sp, fs, keys, th0 = problem(layers, act)
    ref = o.sgld(sp, X, y, th0, B, 1e-5, 4, keys)
    assert rel(npy(ops.sgld(fs, X, y, th0, keys, B, 1e-5, 4)["trace"]), ref) <= RTOL_STEP
    layers, act = (13, 50, 1), o.RELU
    X, y, sp, fs, keys, th0 = problem(layers, act, N=300, scale=0.05)
    minv = np.ones(sp.num_params, np.float32)
    ref, info = o.hmc(sp, X, y, th0, 3, 2e-4, minv, 4, keys, return_info=True)
    out = ops.hmc(fs, X, y, th0, keys, 3, 2e-4, minv, 4, info=True)
    assert np.array_equal(npy(out["accept"]).astype(bool), info["accept"])
    assert rel(npy(out["trace"]), ref) <= RTOL_STEP


def test_global_workspace_state_path(ops, options):
    """Models whose state does not fit shared memory keep it in a global workspace (the 90-256-256-1 case); force that
    path on small nets and check it against the oracle like the resident path."""
    options(force_gstate=1, no_tc2=1)
    for layers, act, B in (NETS[0], NETS[2], NETS[3]):
        X, y, sp, fs, keys, th0 = problem(layers, act)
        ref = o.sgld(sp, X, y, th0, B, 1e-5, 4, keys)
        assert rel(npy(ops.sgld(fs, X, y, th0, keys, B, 1e-5, 4)["trace"]), ref) <= RTOL_STEP
        ref = o.psgld(sp, X, y, th0, B, 1e-6, 3, 0.95, keys)
        assert rel(npy(ops.psgld(fs, X, y, th0, keys, B, 1e-6, 3, 0.95)["trace"]), ref) <= RTOL_STEP
        Xt = X[:50]
        res = ops.predict(fs, th0, Xt, want_preds=True, want_moments=True)
        F = o.predict(sp, th0.astype(np.float64), Xt.astype(np.float64), dtype=np.float64)
        assert rel(npy(res["preds"]), F) <= 1e-5 and rel(npy(res["mean"]), F.mean(0)) <= 1e-5
    layers, act = (4, 6, 6, 1), o.TANH
    X, y, sp, fs, keys, th0 = problem(layers, act, N=70, scale=0.05)
    minv = np.ones(sp.num_params, np.float32)
    ref, info = o.hmc(sp, X, y, th0, 3, 2e-4, minv, 4, keys, return_info=True)
    out = ops.hmc(fs, X, y, th0, keys, 3, 2e-4, minv, 4, info=True)
    assert np.array_equal(npy(out["accept"]).astype(bool), info["accept"])
    assert rel(npy(out["trace"]), ref) <= RTOL_STEP
