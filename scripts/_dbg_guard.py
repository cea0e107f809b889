import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from pbnn_b200.ops import FlatSpec, default_ops
ops = default_ops()
which = sys.argv[1]
layers, B, N, C = (9, 100, 100, 1), 128, 400, 3
rng = np.random.default_rng(0)
X = rng.standard_normal((N, layers[0])).astype(np.float32)
y = rng.standard_normal((N, 1)).astype(np.float32)
fs = FlatSpec(tuple(layers), 0, 0.1)
keys = ops.split(np.array([[0, 11]], np.uint32), C)[0]
th0 = 0.01 * ops.normal(keys, fs.num_params)
idx = np.random.default_rng(1).integers(0, N, (C, 2, B)).astype(np.int32)
f = {"sgld": lambda: ops.sgld(fs, X, y, th0, keys, B, 1e-6, 3, thin=2),
     "per_step": lambda: ops.sgld(fs, X, y, th0, keys, B, 1e-6, 2, minibatch_mode="per_step", keep_trace=False),
     "psgld": lambda: ops.psgld(fs, X, y, th0, keys, B, 1e-7, 2, 0.95),
     "sghmc": lambda: ops.sghmc(fs, X, y, th0, keys, B, 1e-7, 1, 2, momentum_resampling="sigma_sqrt_step"),
     "adam": lambda: ops.adam_ensemble(fs, X, y, th0, idx, 1e-3),
     "grad": lambda: ops.grad_estimator(fs, X, y, th0, idx[:, 0]),
     "predict": lambda: ops.predict(fs, th0, X, want_preds=True, want_moments=True)}[which]
f(); torch.cuda.synchronize(); print(which, "ok")
