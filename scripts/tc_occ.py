"""Do two tcgen05 HMC CTAs share an SM?  time(C = 296) vs time(C = 148) with one CTA per chain (option no_sched)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from pbnn_b200.ops import FlatSpec, default_ops
import bench
ops = default_ops()
ops.set_option("no_sched", 1)
w = dict(bench.WORKLOADS["hmc_cfg2"])
fs = FlatSpec(tuple(w["layers"]), 0, 0.1)
X, y, Xt = bench.make_data(w)
minv = np.ones(fs.num_params, np.float32)
for C in (74, 148, 149, 222, 296, 444, 592):
    keys = ops.split(np.array([[0, 0]], np.uint32), C)[0].contiguous()
    th0 = 0.01 * ops.normal(keys, fs.num_params)
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ops.hmc(fs, X, y, th0, keys, 30, 5e-4, minv, 20, keep_trace=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"C={C:4d}: {dt*1e3:8.2f} ms  -> {C*30/dt:10.0f} chain-steps/s", flush=True)
