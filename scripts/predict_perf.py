"""Posterior-predictive kernel throughput: S stored samples x M test points (north_star subsystem 4)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from pbnn_b200.ops import FlatSpec, default_ops
ops = default_ops()
for layers, S, M in (((13, 50, 1), 51200, 1024), ((9, 100, 100, 1), 20000, 1024), ((8, 50, 50, 1), 512, 4096), ((90, 256, 256, 1), 1024, 4096)):
    fs = FlatSpec(layers, 0, 0.1)
    smp = (0.1 * torch.randn(S, fs.num_params, device="cuda")).contiguous()
    Xt = torch.randn(M, layers[0], device="cuda")
    W = sum(layers[i] * layers[i + 1] for i in range(len(layers) - 1))
    for want_preds, opts in ((False, {}), (True, {}), (False, {"no_tc2": 1})):
        ops.reset_options()
        for k_, v_ in opts.items():
            ops.set_option(k_, v_)
        for it in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = ops.predict(fs, smp, Xt, want_preds=want_preds, want_moments=True)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        fl = 2.0 * S * M * W
        print(f"{layers} S={S} M={M} preds={want_preds} {opts}: {dt*1e3:8.2f} ms  {fl/dt/1e12:6.2f} TFLOP/s  {S*M/dt/1e9:6.2f} G evals/s", flush=True)
