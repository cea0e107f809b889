import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, bench
from pbnn_b200.ops import FlatSpec, default_ops
w = dict(bench.WORKLOADS["hmc_cfg2"]); ops = default_ops(); fs = FlatSpec(tuple(w["layers"]), 0, 0.1)
Xh, yh, Xth, keysh, th0h = bench.make_inputs(w, 0, 1, ops)
minv = torch.ones(fs.num_params, device="cuda")
for eps in (1e-3, 5e-4, 2.5e-4, 1e-4, 5e-5):
    r = ops.hmc(fs, Xh, yh, th0h, keysh, 200, eps, minv, 20, keep_trace=False, info=True)
    acc = r["accept"].float()
    print(f"eps={eps:g} accept overall {acc.mean().item():.3f} first50 {acc[:, :50].mean().item():.3f} last50 {acc[:, -50:].mean().item():.3f} logp {r['logp'].mean().item():.1f} finite {bool(torch.isfinite(r['theta']).all())}")
