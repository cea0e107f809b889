"""Generates the golden fixtures under tests/golden/ (committed together with this script).

The reference pins no numerical results (tests/test_dummpy.py:1-2 is `assert True`) and cannot be executed here
(no JAX), so the fixtures are:
  * kat.json          -- threefry2x32 known-answer vectors (Random123 distribution, kat_vectors) and values printed
                         in the public JAX documentation for jax >= 0.5 (partitionable) and jax < 0.5 (legacy);
                         these are EXTERNAL facts, typed in by hand, not produced by the oracle;
  * tiny_<algo>.npz   -- hand-checkable 3-step trajectories of every sampler on a [2,3,1] tanh MLP, computed by the
                         oracle in float64 (the yard-stick) and float32.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import pbnn_oracle as o  # noqa: E402

KAT = {
    "threefry2x32_20": [  # key, counter -> output (Random123 kat_vectors, threefry2x32 20 rounds)
        {"key": [0, 0], "ctr": [0, 0], "out": [0x6B200159, 0x99BA4EFE]},
        {"key": [0xFFFFFFFF, 0xFFFFFFFF], "ctr": [0xFFFFFFFF, 0xFFFFFFFF], "out": [0x1CB996FC, 0xBB002BE7]},
        {"key": [0x13198A2E, 0x03707344], "ctr": [0x243F6A88, 0x85A308D3], "out": [0xC4923A9C, 0x483DF7A0]},
    ],
    "jax_docs_partitionable": {  # jax >= 0.5 defaults
        "split(PRNGKey(0),2)": [[1797259609, 2579123966], [928981903, 3453687069]],
        "split(PRNGKey(42),2)": [[1832780943, 270669613], [64467757, 2916123636]],
        "normal(PRNGKey(0),())": 1.6226422,
        "normal(PRNGKey(42),())": -0.028304616,
        "uniform(PRNGKey(0),())": 0.947667,
    },
    "jax_docs_legacy": {  # jax < 0.5 defaults: a mis-configured oracle would reproduce these instead
        "normal(PRNGKey(42),())": -0.18471177,
        "normal(PRNGKey(0),())": -0.20584226,
    },
}


def tiny_problem():
    rng = np.random.default_rng(123)
    X = rng.standard_normal((12, 2)).astype(np.float32)
    y = (np.sin(X[:, :1]) + 0.1 * rng.standard_normal((12, 1))).astype(np.float32)
    sp = o.MlpSpec((2, 3, 1), o.TANH, 0.5)
    keys = o.split(o.PRNGKey(2024), 2)
    th0 = (0.5 * rng.standard_normal((2, sp.num_params))).astype(np.float32)
    return X, y, sp, keys, th0


def main():
    json.dump(KAT, open(os.path.join(HERE, "kat.json"), "w"), indent=1)
    X, y, sp, keys, th0 = tiny_problem()
    B, T = 4, 3
    out = {"X": X, "y": y, "keys": keys, "theta0": th0, "layers": np.array(sp.layers), "act": sp.act,
           "noise_level": sp.noise_level, "B": B, "T": T}
    for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
        out[f"idx_draw1"] = np.stack([o.reference_draw(k, B, len(X), 1) for k in keys])
        out[f"idx_draw2"] = np.stack([o.reference_draw(k, B, len(X), 2) for k in keys])
        out[f"sgld_{tag}"] = o.sgld(sp, X, y, th0, B, 1e-3, T, keys, dtype=dt)
        out[f"psgld_{tag}"] = o.psgld(sp, X, y, th0, B, 1e-3, T, 0.95, keys, dtype=dt)
        out[f"sghmc_{tag}"] = o.sghmc(sp, X, y, th0, B, 1e-3, T, 2, keys, dtype=dt)
        tr, info = o.hmc(sp, X, y, th0, T, 5e-2, np.ones(sp.num_params), 4, keys, dtype=dt, return_info=True)
        out[f"hmc_{tag}"], out[f"hmc_delta_{tag}"], out[f"hmc_accept_{tag}"] = tr, info["delta"], info["accept"]
        idx = np.stack([np.arange(12).reshape(3, 4)] * 2).astype(np.int32)
        out["adam_idx"] = idx
        out[f"adam_{tag}"] = o.deep_ensembles(sp, X, y, th0, B, 0, 1e-2, None, dtype=dt, idx=idx)
        m, v = o.predictive_moments(sp, th0, X, dtype=dt)
        out[f"pred_mean_{tag}"], out[f"pred_var_{tag}"] = m, v
        g = o.grad_estimator(sp, th0.astype(dt), X[out["idx_draw2"]].astype(dt), y[out["idx_draw2"]].astype(dt), len(X))
        out[f"grad_{tag}"] = g
    np.savez(os.path.join(HERE, "tiny.npz"), **out)
    print("wrote", os.path.join(HERE, "kat.json"), os.path.join(HERE, "tiny.npz"))


if __name__ == "__main__":
    main()
