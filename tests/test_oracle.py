"""Pins the oracle (CPU, no GPU): RNG against external known answers, math against first principles, the C
restatement against the numpy restatement, and both against the committed golden fixtures."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle as co
from oracle import pbnn_oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))
TINY = np.load(os.path.join(HERE, "golden", "tiny.npz"))


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))


def test_threefry_kats():
    for v in KAT["threefry2x32_20"]:
        x0, x1 = o.threefry2x32(v["key"][0], v["key"][1], v["ctr"][0], v["ctr"][1])
        assert [int(x0), int(x1)] == v["out"]


def test_documented_jax_values():
    d = KAT["jax_docs_partitionable"]
    assert o.split(o.PRNGKey(0), 2).tolist() == d["split(PRNGKey(0),2)"]
    assert o.split(o.PRNGKey(42), 2).tolist() == d["split(PRNGKey(42),2)"]
    assert abs(float(o.normal(o.PRNGKey(0), 1)[0]) - d["normal(PRNGKey(0),())"]) < 1e-6
    assert abs(float(o.normal(o.PRNGKey(42), 1)[0]) - d["normal(PRNGKey(42),())"]) < 1e-7
    assert abs(float(o.uniform(o.PRNGKey(0), 1)[0]) - d["uniform(PRNGKey(0),())"]) < 1e-6
    # the legacy (jax < 0.5) bit layout gives different, also documented, values: guards against a wrong mode
    leg = KAT["jax_docs_legacy"]
    for seed in (0, 42):
        z = float(o.bits_to_normal(o.random_bits_legacy(o.PRNGKey(seed), 1))[0])
        assert abs(z - leg[f"normal(PRNGKey({seed}),())"]) < 1e-6
        assert abs(float(o.normal(o.PRNGKey(seed), 1)[0]) - z) > 1e-3


def test_rng_statistics_and_structure():
    z = o.normal(o.PRNGKey(7), 200000)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    u = o.uniform(o.PRNGKey(7), 100000)
    assert u.min() >= 0 and u.max() < 1 and abs(u.mean() - 0.5) < 0.01
    assert np.array_equal(o.fold_in(o.PRNGKey(3), 5), o.split(o.PRNGKey(3), 6)[5])
    for N in (1, 2, 17, 1030, 9568):
        p = o.permutation(o.PRNGKey(N), N)
        assert np.array_equal(np.sort(p), np.arange(N))
    r = o.randint(o.PRNGKey(1), 50000, 0, 7)
    assert r.min() == 0 and r.max() == 6 and np.abs(np.bincount(r) / 50000 - 1 / 7).max() < 0.01
    # span > 2**16: lax.mul wraps, the multiplier is 0 and only the low word is used
    big = o.randint(o.PRNGKey(1), 100, 0, 515345)
    k2 = o.split(o.PRNGKey(1), 2)[1]
    assert np.array_equal(big, (o.random_bits(k2, 100) % 515345).astype(np.int32))


def test_erfinv_against_scipy():
    from scipy.special import erfinv
    x = np.linspace(-0.999999, 0.999999, 20001).astype(np.float32)
    assert np.abs(o.erf_inv_f32(x) - erfinv(x.astype(np.float64))).max() < 5e-6 * 4


@pytest.mark.parametrize("layers,act", [((3, 5, 1), o.TANH), ((4, 6, 6, 2), o.TANH), ((2, 3, 1), o.RELU)])
def test_gradient_finite_differences(layers, act):
    rng = np.random.default_rng(0)
    sp = o.MlpSpec(layers, act, 0.3)
    N, B = 40, 8
    X = rng.standard_normal((N, layers[0]))
    y = rng.standard_normal((N, layers[-1]))
    th = 0.5 * rng.standard_normal((1, sp.num_params))
    idx = rng.integers(0, N, B)

    def f(t):
        ll, _ = o.loglik_and_grad(sp, t, X[idx], y[idx], 1.0)
        return (o.logprior(sp, t) + N * ll / B)[0]

    g = o.grad_estimator(sp, th, X[idx], y[idx], N)[0]
    fd = np.zeros_like(g)
    for j in range(sp.num_params):
        e = np.zeros_like(th)
        e[0, j] = 1e-6
        fd[j] = (f(th + e) - f(th - e)) / 2e-6
    assert rel(g, fd) < 1e-3 if act == o.RELU else rel(g, fd) < 1e-6


def test_flat_layout_matches_ravel_pytree_order():
    sp = o.MlpSpec((8, 50, 1), o.RELU, 0.1)
    assert sp.num_params == 501
    assert sp.offsets() == [(0, 50, 8, 50), (450, 451, 50, 1)]
    sp = o.MlpSpec((9, 100, 100, 1), o.RELU, 0.1)
    assert sp.num_params == 11201 and [t[:2] for t in sp.offsets()] == [(0, 100), (1000, 1100), (11100, 11101)]
    assert o.MlpSpec((90, 256, 256, 1)).num_params == 89345


def test_adam_matches_torch():
    rng = np.random.default_rng(0)
    th = rng.standard_normal(50).astype(np.float32)
    p = torch.nn.Parameter(torch.tensor(th.copy()))
    opt = torch.optim.Adam([p], lr=1e-2)
    m = v = np.zeros_like(th)
    cnt, t = 0, th.copy()
    for k in range(20):
        g = np.sin(t + k).astype(np.float32)
        t, m, v, cnt = o.adam_step(t, g, m, v, cnt, 1e-2)
        p.grad = torch.sin(p.detach() + k)
        opt.step()
    assert np.abs(t - p.detach().numpy()).max() < 2e-6


def test_hmc_energy_conservation_and_acceptance():
    X, y = o.synthetic_regression(60, 3, 0)
    sp = o.MlpSpec((3, 4, 1), o.TANH, 0.5)
    keys = o.split(o.PRNGKey(0), 8)
    th0 = o.init_positions(sp, keys)
    _, info = o.hmc(sp, X, y, th0, 20, 1e-4, np.ones(sp.num_params), 5, keys, dtype=np.float64, return_info=True)
    assert np.abs(info["delta"]).max() < 1e-3 and info["accept"].mean() > 0.99  # eps -> 0 => accept ~ 1
    _, info = o.hmc(sp, X, y, th0, 30, 0.3, np.ones(sp.num_params), 5, keys, dtype=np.float64, return_info=True)
    assert info["accept"].mean() < 0.9  # a coarse step does reject


def test_reference_quirks():
    # Q1: draw #2 is the fixed minibatch; draws are a deterministic function of the key chain
    k = o.PRNGKey(5)
    d1, d2 = o.reference_draw(k, 8, 100, 1), o.reference_draw(k, 8, 100, 2)
    it = o.batch_index_stream(k, 8, 100)
    assert np.array_equal(next(it), d1) and np.array_equal(next(it), d2)
    # Q3/Q4: num_epochs + 1 epochs; members k >= 1 are seeded from PRNGKey(k)
    mk = o.ensemble_member_keys(o.PRNGKey(9), 3)
    assert np.array_equal(mk[1][0], o.split(o.PRNGKey(1), 2)[0]) and np.array_equal(mk[2][1], o.split(o.PRNGKey(2), 2)[1])


def test_oracle_reproduces_golden_fixtures():
    sp = o.MlpSpec(tuple(int(v) for v in TINY["layers"]), int(TINY["act"]), float(TINY["noise_level"]))
    X, y, keys, th0, B, T = TINY["X"], TINY["y"], TINY["keys"], TINY["theta0"], int(TINY["B"]), int(TINY["T"])
    assert np.array_equal(np.stack([o.reference_draw(k, B, len(X), 2) for k in keys]), TINY["idx_draw2"])
    assert np.array_equal(o.sgld(sp, X, y, th0, B, 1e-3, T, keys), TINY["sgld_f32"])
    assert np.array_equal(o.psgld(sp, X, y, th0, B, 1e-3, T, 0.95, keys), TINY["psgld_f32"])
    assert np.array_equal(o.sghmc(sp, X, y, th0, B, 1e-3, T, 2, keys), TINY["sghmc_f32"])
    assert np.array_equal(o.hmc(sp, X, y, th0, T, 5e-2, np.ones(sp.num_params), 4, keys), TINY["hmc_f32"])
    for a in ("sgld", "psgld", "sghmc", "hmc", "adam"):
        assert rel(TINY[f"{a}_f32"], TINY[f"{a}_f64"]) < 1e-5


def test_c_restatement_matches_numpy_restatement():
    k = o.split(o.PRNGKey(1), 4)
    assert np.abs(co.normal(k, 5000) - np.stack([o.normal(x, 5000) for x in k])).max() < 5e-7
    for B, N in ((32, 1030), (128, 515345)):
        assert np.array_equal(co.reference_draw(k, 2, B, N), np.stack([o.reference_draw(x, B, N, 2) for x in k]))
    layers = (9, 20, 20, 1)
    X, y = o.synthetic_regression(300, 9, 1)
    sp = o.MlpSpec(layers, o.RELU, 0.1)
    th0 = (0.2 * np.random.default_rng(0).standard_normal((4, sp.num_params))).astype(np.float32)
    _, tr = co.sg_run(layers, 0, 0.1, "sgld", X, y, th0, 40, 1e-6, 5, k)
    assert rel(tr, o.sgld(sp, X, y, th0, 40, 1e-6, 5, k)) < 1e-6
    _, tr = co.sg_run(layers, 0, 0.1, "psgld", X, y, th0, 40, 1e-6, 5, k, alpha=0.95)
    assert rel(tr, o.psgld(sp, X, y, th0, 40, 1e-6, 5, 0.95, k)) < 1e-6
    _, tr = co.sg_run(layers, 0, 0.1, "sghmc", X, y, th0, 40, 1e-9, 3, k, L=3)
    assert rel(tr, o.sghmc(sp, X, y, th0, 40, 1e-9, 3, 3, k)) < 1e-5
    _, tr, acc = co.hmc_run(layers, 0, 0.1, X, y, th0 * 0.2, k, 4, 2e-4, np.ones(sp.num_params), 5)
    ref, info = o.hmc(sp, X, y, th0 * 0.2, 4, 2e-4, np.ones(sp.num_params), 5, k, return_info=True)
    assert rel(tr, ref) < 1e-6 and np.array_equal(acc.astype(bool), info["accept"])
    idx = np.random.default_rng(3).integers(0, 300, (4, 12, 16)).astype(np.int32)
    assert rel(co.adam_run(layers, 0, 0.1, X, y, th0, idx, 1e-3), o.deep_ensembles(sp, X, y, th0, 16, 0, 1e-3, None, idx=idx)) < 1e-5
    sm, sq = co.predict_moments(layers, 0, 0.1, th0, X[:50])
    F = o.predict(sp, th0, X[:50])
    assert rel(sm, F.sum(0)) < 1e-6 and rel(sq, (F ** 2).sum(0)) < 1e-6
