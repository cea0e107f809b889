"""tcgen05 HMC engine vs the scalar fused engine vs the oracle: gradient accuracy and a short HMC run (GPU)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from oracle import pbnn_oracle as o  # noqa: E402
from pbnn_b200.ops import FlatSpec, default_ops  # noqa: E402

ops = default_ops()
rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
for layers, N in [((13, 50, 1), 506), ((8, 50, 1), 300), ((5, 64, 1), 100), ((15, 33, 1), 896), ((3, 7, 1), 129)]:
    X, y = o.synthetic_regression(N, layers[0], 1)
    sp, fs = o.MlpSpec(layers, o.RELU, 0.1), FlatSpec(layers, o.RELU, 0.1)
    C = 5
    th = (0.2 * np.random.default_rng(0).standard_normal((C, sp.num_params))).astype(np.float32)
    lp64, g64 = o.logposterior_full(sp, th.astype(np.float64), X.astype(np.float64), y.astype(np.float64))
    ops.reset_options()
    lp_tc, g_tc = ops.logposterior_grad(fs, X, y, th)
    torch.cuda.synchronize()
    ops.set_option("no_tc", 1)
    lp_h1, g_h1 = ops.logposterior_grad(fs, X, y, th)
    torch.cuda.synchronize()
    ops.reset_options()
    print(layers, N, "grad rel err vs f64: tc %.2e  fused %.2e | logp rel: tc %.2e fused %.2e" % (
        rel(g_tc.cpu().numpy(), g64), rel(g_h1.cpu().numpy(), g64),
        rel(lp_tc.cpu().numpy(), lp64), rel(lp_h1.cpu().numpy(), lp64)), flush=True)
    gt, g6 = g_tc.cpu().numpy(), g64
    P0 = layers[1] * (layers[0] + 1)
    print("   per-block: first-layer %.2e  second-layer %.2e" % (rel(gt[:, :P0], g6[:, :P0]), rel(gt[:, P0:], g6[:, P0:])))

# short HMC run vs oracle
layers, N = (13, 50, 1), 506
X, y = o.synthetic_regression(N, layers[0], 0)
sp, fs = o.MlpSpec(layers, o.RELU, 0.1), FlatSpec(layers, o.RELU, 0.1)
C = 4
keys = o.split(o.PRNGKey(0), C)
th0 = o.init_positions(sp, keys)
minv = np.ones(sp.num_params, np.float32)
ref, info = o.hmc(sp, X, y, th0, 3, 1e-4, minv, 5, keys, return_info=True)
out = ops.hmc(fs, X, y, th0, keys, 3, 1e-4, minv, 5, info=True)
print("hmc trace rel err", rel(out["trace"].cpu().numpy(), ref), "accept", out["accept"].cpu().numpy().tolist(),
      info["accept"].astype(int).tolist())
print("delta", out["delta"].cpu().numpy()[0], info["delta"][0])

# timing: cfg2
C, S, L = 256, 50, 20
keys = o.split(o.PRNGKey(1), C)
th0 = o.init_positions(sp, keys)
for name, env in (("tc", None), ("fused", "1")):
    if env:
        ops.set_option("no_tc", int(env))
    else:
        ops.reset_options()
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = ops.hmc(fs, X, y, th0, keys, S, 5e-4, minv, L)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(name, "cfg2: %.2f ms for %d chain-steps -> %.0f chain-steps/s, accept %.2f" % (
        dt * 1e3, C * S, C * S / dt, float(out["accept_count"].float().mean()) / S), flush=True)
