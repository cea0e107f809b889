import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import pbnn_oracle as o
from pbnn_b200.ops import FlatSpec, default_ops
ops = default_ops()
layers = (8, 50, 1)
X, y = o.synthetic_regression(1030, 8, 5)
fs = FlatSpec(layers, 0, 0.1)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4
keys = o.split(o.PRNGKey(12), C)
th0 = (0.2 * np.random.default_rng(5).standard_normal((C, 501))).astype(np.float32)
for opt in ("w1_min_chains", "no_w1"):
    ops.reset_options(); ops.set_option(opt, 1)
    runs = [ops.sgld(fs, X, y, th0, keys, 32, 1e-5, 12) for _ in range(6)]
    tr = [r["trace"].cpu().numpy() for r in runs]
    fl = [r["flags"].cpu().numpy() for r in runs]
    for i in range(1, 6):
        d = np.argwhere(tr[i] != tr[0])
        nan = np.argwhere(~np.isfinite(tr[i]))
        print(opt, "run", i, "ndiff", len(d), "first", d[:3].tolist(), "nan", len(nan), nan[:3].tolist(), "flags", fl[i].tolist()[:8])
ref = o.sgld(o.MlpSpec(layers, o.RELU, 0.1), X, y, th0[:4], 32, 1e-5, 12, keys[:4])
print("rel vs oracle per step (chain 2):", [float(np.linalg.norm(tr[0][2, t] - ref[2, t]) / np.linalg.norm(ref[2, t])) for t in range(12)])
