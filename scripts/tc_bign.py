"""tcgen05 HMC engine on larger data sets (X streamed through TMEM per group of 4 tiles) vs the FP32 fused engine."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from pbnn_b200.ops import FlatSpec, default_ops
ops = default_ops()
for layers, N in (((8, 50, 1), 1030), ((8, 50, 1), 768), ((11, 50, 1), 1599), ((8, 50, 1), 4096)):
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N, layers[0])).astype(np.float32)
    y = np.tanh(X[:, :1]).astype(np.float32)
    fs = FlatSpec(layers, 0, 0.1)
    C = 256
    keys = ops.split(np.array([[0, 0]], np.uint32), C)[0].contiguous()
    th0 = 0.01 * ops.normal(keys, fs.num_params)
    minv = np.ones(fs.num_params, np.float32)
    for name, env in (("tc", None), ("fused", "1")):
        if env: ops.set_option("no_tc", int(env))
        else: ops.reset_options()
        for it in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = ops.hmc(fs, X, y, th0, keys, 20, 2e-4, minv, 20, keep_trace=False)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{layers} N={N} {name}: {dt*1e3:8.2f} ms -> {C*20/dt:10.0f} chain-steps/s accept {float(out['accept_count'].float().mean())/20:.2f}", flush=True)
