"""The pbnn-compatible Python surface (same names / argument order as src/pbnn/mcmc/*.py and deep_ensembles.py).
Runs against the emulator on the CPU box and against the CUDA library under -m gpu."""
import numpy as np
import pytest

import pbnn_b200
from oracle import pbnn_oracle as o
from pbnn_b200 import logprobs, pytree
from pbnn_b200.mcmc import (cyclical_sgld, hmc, kernels, pSGLD, predictive_moments, scheduled_sgld, sghmc, sghmc_cv,
                            sghmc_svrg, sgld, sgld_cv, sgld_svrg)
from pbnn_b200.models import MLP
from pbnn_b200.ops import set_default_ops


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def npy(t):
    return t.detach().cpu().numpy()


@pytest.fixture
def api(ops):
    set_default_ops(ops)
    yield ops
    set_default_ops(None)


def setup(C=None, layers=(8, 50, 1), act="relu", N=200):
    X, y = o.synthetic_regression(N, layers[0], 3)
    net = MLP.from_widths(layers, act)
    sp = o.MlpSpec(layers, o.RELU if act == "relu" else o.TANH, 0.1)
    key = o.PRNGKey(7) if C is None else o.split(o.PRNGKey(7), C)
    th0 = o.init_positions(sp, key)
    tree = {k: {kk: vv if C is not None else vv[0] for kk, vv in v.items()} for k, v in o.to_pytree(sp, th0).items()}
    lik = logprobs.GaussianMLPLikelihood(net, 0.1)
    return X, y, net, sp, key, th0, tree, lik, logprobs.StdNormalPrior()


def test_sgld_single_chain_signature(api):
    X, y, net, sp, key, th0, tree, lik, prior = setup()
    positions, ravel_fn, predict_fn = sgld(X, y, lik, prior, tree, 32, 1e-5, 8, key)
    assert set(positions) == {"Dense_0", "Dense_1"}
    assert tuple(positions["Dense_0"]["kernel"].shape) == (8, 8, 50)      # leading T, Flax layout (in, out)
    assert tuple(positions["Dense_1"]["bias"].shape) == (8, 1)
    flat = ravel_fn(positions)
    assert tuple(flat.shape) == (8, sp.num_params)
    ref = o.sgld(sp, X, y, th0, 32, 1e-5, 8, key)
    assert rel(npy(flat), ref[0]) <= 1e-5
    preds = predict_fn(net, positions, X[:20])
    assert tuple(preds.shape) == (8, 20, 1)
    assert rel(npy(preds), o.predict(sp, ref[0], X[:20])) <= 1e-5
    mean, var = predictive_moments(net, positions, X[:20])
    assert rel(npy(mean), o.predict(sp, ref[0], X[:20]).mean(0)) <= 1e-5


def test_samplers_batched_chains(api):
    C = 3
    X, y, net, sp, keys, th0, tree, lik, prior = setup(C, layers=(9, 20, 20, 1), act="tanh")
    pos, ravel_fn, _ = pSGLD(X, y, lik, prior, tree, 40, 1e-6, 5, 0.95, keys)
    assert tuple(pos["Dense_1"]["kernel"].shape) == (C, 5, 20, 20)
    assert rel(npy(ravel_fn(pos)), o.psgld(sp, X, y, th0, 40, 1e-6, 5, 0.95, keys)) <= 1e-5
    pos, ravel_fn, _ = sghmc(X, y, lik, prior, tree, 40, 1e-9, 3, 2, keys)
    assert rel(npy(ravel_fn(pos)), o.sghmc(sp, X, y, th0, 40, 1e-9, 3, 2, keys)) <= 1e-5
    sched = lambda t: 1e-5 / (1 + t)
    pos, ravel_fn, _ = scheduled_sgld(X, y, lik, prior, tree, 40, sched, 4, keys)
    th, ref = th0.copy(), []
    for t in range(4):
        th = o.sgld(sp, X, y, th, 40, np.float32(sched(t + 1)), 1, keys, t0=t)[:, 0]
        ref.append(th)
    assert rel(npy(ravel_fn(pos)), np.stack(ref, 1)) <= 1e-5
    logprob = logprobs.FullBatchLogPosterior(lik, prior, X, y)
    minv = np.ones(sp.num_params, np.float32)
    pos, ravel_fn, predict_fn, info = hmc(logprob, tree, 3, 1e-4, minv, 4, keys, return_info=True)
    ref, rinfo = o.hmc(sp, X, y, th0, 3, 1e-4, minv, 4, keys, return_info=True)
    assert np.array_equal(npy(info["is_accepted"]).astype(bool), rinfo["accept"])
    assert rel(npy(ravel_fn(pos)), ref) <= 1e-5


def test_density_objects_are_callable_and_checked(api):
    X, y, net, sp, key, th0, tree, lik, prior = setup()
    lp = float(prior(tree))
    assert abs(lp - float(o.logprior(sp, th0)[0])) < 1e-3
    ll = float(lik(tree, (X[:16], y[:16])))
    ll_ref, _ = o.loglik_and_grad(sp, th0, X[:16], y[:16], 1.0)
    assert abs(ll - float(ll_ref[0])) <= 1e-4 * abs(float(ll_ref[0]))
    full = logprobs.FullBatchLogPosterior(lik, prior, X, y)
    ref = o.logposterior_full(sp, th0, X, y)[0][0]
    assert abs(float(full(tree)) - ref) <= 1e-4 * abs(ref)
    with pytest.raises(TypeError):
        sgld(X, y, lambda p, d: 0.0, prior, tree, 32, 1e-5, 2, key)
    with pytest.raises(TypeError):
        sgld(X, y, lik, lambda p: 0.0, tree, 32, 1e-5, 2, key)
    with pytest.raises(TypeError):
        hmc(lambda p: 0.0, tree, 2, 1e-3, np.ones(sp.num_params), 2, key)


def test_blackjax_style_psgld_kernel(api):
    """kernels.psgld(grad_estimator_fn, alpha) -> SamplingAlgorithm(init, step) (kernels.py:80-94)."""
    X, y, net, sp, key, th0, tree, lik, prior = setup(layers=(8, 12, 1))
    N, B = len(X), 16
    grad_fn = kernels.grad_estimator(prior, lik, N)
    kern = kernels.psgld(grad_fn, 0.9)
    batch = (X[:B], y[:B])
    state = kern.init(tree, batch)
    g_ref = o.grad_estimator(sp, th0, X[:B], y[:B], N)
    assert state.step == 0
    assert rel(npy(pytree.ravel(state.logprob_grad)), g_ref[0]) <= 1e-5
    assert rel(npy(pytree.ravel(state.square_avg)), g_ref[0] ** 2) <= 2e-5
    step_key = o.split(key, 3)[2]
    new = kern.step(step_key, state, (X[B:2 * B], y[B:2 * B]), 1e-6)
    G = 1.0 / (np.sqrt(g_ref ** 2) + 1e-6)
    th1 = th0 + np.float32(1e-6) * G * g_ref + np.sqrt(2 * G * np.float32(1e-6)) * o.normal(step_key, sp.num_params)
    assert new.step == 1
    assert rel(npy(pytree.ravel(new.position)), th1[0]) <= 1e-5
    g1 = o.grad_estimator(sp, th1.astype(np.float32), X[B:2 * B], y[B:2 * B], N)
    assert rel(npy(pytree.ravel(new.logprob_grad)), g1[0]) <= 1e-4
    assert rel(npy(pytree.ravel(new.square_avg)), 0.9 * g_ref[0] ** 2 + 0.1 * g1[0] ** 2) <= 1e-4


def test_deep_ensembles(api):
    layers = (8, 10, 10, 1)
    X, y = o.synthetic_regression(96, 8, 3)
    net = MLP(10, 1)  # reference MLP: two equal hidden layers (models.py:19-35)
    assert net.widths(8) == layers
    sp = o.MlpSpec(layers, o.RELU, 0.1)
    lik, prior = logprobs.GaussianMLPLikelihood(net, 0.1), logprobs.StdNormalPrior()
    K, B, E = 3, 16, 1
    key = o.PRNGKey(11)
    init = o.init_positions(sp, o.split(o.PRNGKey(5), K))
    params, ravel_fn, predict_fn = pbnn_b200.deep_ensembles_fn(X, y, lik, prior, net, B, E, 1e-3, K, key,
                                                              init_params=o.to_pytree(sp, init))
    assert tuple(params["Dense_1"]["kernel"].shape) == (K, 10, 10)
    mk = [k for k, _ in o.ensemble_member_keys(key, K)]
    ref = o.deep_ensembles(sp, X, y, init, B, E, 1e-3, mk)
    assert rel(npy(ravel_fn(params)), ref) <= 2e-5
    assert tuple(predict_fn(net, params, X[:5]).shape) == (K, 5, 1)
    # default initialisation path (Flax-style N(0, 0.01^2) init on device): runs and is finite
    params2, _, _ = pbnn_b200.deep_ensembles_fn(X, y, lik, prior, net, B, 0, 1e-3, 2, key)
    flat2 = npy(pytree.ravel(params2))
    assert np.isfinite(flat2).all() and flat2.shape == (2, sp.num_params)


def test_flax_style_init_and_apply(api):
    net = MLP(6, 1)
    x = np.zeros((4,), np.float32)
    p = net.init(o.PRNGKey(1), x)["params"]
    flat = npy(pytree.ravel(p))
    assert flat.shape == (4 * 6 + 6 + 6 * 6 + 6 + 6 + 1,)
    assert 0.003 < flat.std() < 0.03  # normal(stddev=0.01) for kernels and biases (models.py:19-22)
    X = o.synthetic_regression(9, 4, 0)[0]
    out = net.apply({"params": p}, X)
    sp = o.MlpSpec((4, 6, 6, 1), o.RELU, 0.1)
    assert rel(npy(out), o.predict(sp, flat[None], X)[0]) <= 1e-5


def test_control_variates_and_svrg(api):
    """sgld_cv / sghmc_cv / sgld_svrg / sghmc_svrg (langevin.py:207-354, hamiltonian.py:207-359)."""
    C = 2
    X, y, net, sp, keys, th0, tree, lik, prior = setup(C, layers=(8, 12, 1), act="tanh", N=150)
    centre = (th0 * 0.5).astype(np.float32)
    ctree = o.to_pytree(sp, centre)
    # the two constant terms against the float64 oracle
    idx = np.stack([o.reference_draw(k, 16, len(X), 2) for k in keys])
    gf64, ge64 = o.control_variate_terms(sp, X, y, centre, X[idx], y[idx], dtype=np.float64)
    lp, gf = api.logposterior_grad(lik.flat_spec(8, prior), X, y, centre)
    assert rel(npy(gf), gf64) <= 1e-5
    assert rel(npy(lp), o.logposterior_full(sp, centre.astype(np.float64), X.astype(np.float64), y.astype(np.float64))[0]) <= 1e-5
    pos, ravel_fn, _ = sgld_cv(X, y, lik, prior, tree, 16, 1e-6, 5, ctree, keys)
    assert rel(npy(ravel_fn(pos)), o.sgld_cv(sp, X, y, th0, 16, 1e-6, 5, centre, keys)) <= 1e-5
    pos, ravel_fn, _ = sghmc_cv(X, y, lik, prior, tree, 16, 1e-9, 3, ctree, 2, keys)
    assert rel(npy(ravel_fn(pos)), o.sghmc_cv(sp, X, y, th0, 16, 1e-9, 3, centre, 2, keys)) <= 1e-5
    pos, ravel_fn, _ = sgld_svrg(X, y, lik, prior, tree, 16, 1e-6, ctree, 3, 2, keys)
    flat = npy(ravel_fn(pos))
    assert flat.shape == (C, 6, sp.num_params)
    assert rel(flat, o.svrg("sgld", sp, X, y, th0, 16, 1e-6, centre, 3, 2, keys)) <= 1e-5
    pos, ravel_fn, _ = sghmc_svrg(X, y, lik, prior, tree, 16, 1e-9, ctree, 2, 2, 2, keys)
    assert rel(npy(ravel_fn(pos)), o.svrg("sghmc", sp, X, y, th0, 16, 1e-9, centre, 2, 2, keys, L=2)) <= 1e-5


def test_cyclical_sgld_structure(api):
    """cyclical_sgld (langevin.py:364-470): shapes, burn-in per cycle, temperature-0 SGD phase == plain gradient steps."""
    X, y, net, sp, key, th0, tree, lik, prior = setup(layers=(8, 12, 1), act="tanh", N=150)
    pos, ravel_fn, _ = cyclical_sgld(X, y, lik, prior, tree, 16, 1e-6, 2, key, num_sgd_steps=3, num_sgld_steps=5,
                                     burnin_sgld=2)
    flat = npy(ravel_fn(pos))
    assert flat.shape == (2 * 3, sp.num_params) and np.isfinite(flat).all()
    # SGD phase: theta <- theta + schedule(it) * grad, no noise
    fs = lik.flat_spec(8, prior)
    k1 = o.split(key, 2)[1]
    idx = o.reference_draw(k1, 16, len(X), 1)[None]
    steps = [0.5 * (np.cos(np.pi * it / 8) + 1) * 1e-6 for it in range(3)]
    res = api.sgld(fs, X, y, th0, k1, 16, steps, 3, idx=idx, temperature=0.0)
    th = th0.astype(np.float64)
    for it in range(3):
        th = th + np.float32(steps[it]) * o.grad_estimator(sp, th, X[idx].astype(np.float64), y[idx].astype(np.float64), len(X))
    assert rel(npy(res["theta"]), th) <= 1e-5


def test_thinning_fn_and_npz_wire_format(api, tmp_path):
    """Post-processing right after the chain (benchmark/pipelines/sgmcmc.py:172,213-232): burn-in, thinning, predict,
    savez(positions (S,P), predictions (S,M), time)."""
    from pbnn_b200.utils import save_results, thinning_fn
    X, y, net, sp, key, th0, tree, lik, prior = setup(layers=(8, 12, 1), act="tanh", N=150)
    positions, ravel_fn, predict_fn = sgld(X, y, lik, prior, tree, 16, 1e-5, 60, key, burnin=20)
    flat = ravel_fn(positions)
    idx = thinning_fn(flat, 10)
    assert np.array_equal(npy(idx), o.thinning(npy(flat), 10, dtype=np.float64))
    thinned = flat[idx.long()]
    preds = predict_fn(net, pytree.unravel(thinned, (8, 12, 1)), X[:7]).squeeze(-1)
    save_results(tmp_path / "out.npz", thinned, preds, 1.5)
    z = np.load(tmp_path / "out.npz")
    assert z["positions"].shape == (10, sp.num_params) and z["predictions"].shape == (10, 7) and float(z["time"]) == 1.5


def test_map_estimation_train_fn(api):
    """map_estimation.train_fn (src/pbnn/map_estimation.py:42-108): Adam / SGD on -logposterior, one fixed batch order
    replayed every epoch."""
    from pbnn_b200.map_estimation import build_logposterior_estimator_fn, train_fn
    X, y, net, sp, key, th0, tree, lik, prior = setup(layers=(8, 12, 1), act="tanh", N=96)
    logpost = build_logposterior_estimator_fn(prior, lik, len(X))
    batches = np.random.default_rng(0).permutation(96)[:80].reshape(5, 16).astype(np.int32)
    params = train_fn(logpost, net, {"x": X, "y": y}, 16, 3, 1e-3, key, "adam", init_params=tree, batch_indices=batches)
    idx = np.tile(batches, (3, 1))[None]
    ref = o.deep_ensembles(sp, X, y, th0, 16, 0, 1e-3, None, idx=idx)
    assert rel(npy(pytree.ravel(params)), ref[0]) <= 2e-5
    params = train_fn(logpost, net, {"x": X, "y": y}, 16, 2, 1e-6, key, "sgd", init_params=tree, batch_indices=batches)
    th = th0.astype(np.float64)
    for row in np.tile(batches, (2, 1)):
        th = th + np.float32(1e-6) * o.grad_estimator(sp, th, X[row][None].astype(np.float64), y[row][None].astype(np.float64), 96)
    assert rel(npy(pytree.ravel(params)), th[0]) <= 1e-5
    import torch
    torch.manual_seed(0)
    p1 = train_fn(logpost, net, {"x": X, "y": y}, 16, 1, 1e-3, key)  # default init + torch-sampler batch order
    assert np.isfinite(npy(pytree.ravel(p1))).all()
    with pytest.raises(TypeError):
        train_fn(lambda p, b: 0.0, net, {"x": X, "y": y}, 16, 1, 1e-3, key)
