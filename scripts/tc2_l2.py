"""Does the tc2 engine's working set fit the L2?  Time per chain-step per CTA with 148 / 128 / 112 / 96 / 74 persistent CTAs
(option tc2_ctas), whole waves each: if fewer CTAs are faster per CTA, the full grid is thrashing the L2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from pbnn_b200.ops import FlatSpec, default_ops

ops = default_ops()
name = sys.argv[1] if len(sys.argv) > 1 else "sgld_cfg5"
waves = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for ctas in (148, 128, 112, 96, 74):
    w = dict(bench.WORKLOADS[name], name=name)
    w["C"] = ctas * waves
    fs = FlatSpec(tuple(w["layers"]), w["act"], 0.1)
    Xh, yh, Xth, keys, th0 = bench.make_inputs(w, 0, 1, ops)
    X, y = (ops._as(a, torch.float32) for a in (Xh, yh))
    step = bench.gpu_step_fn(ops, w, fs, X, y, th0, keys)
    ops.set_option("tc2_ctas", ctas)
    step(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    per = best * 1e3 / (waves * w["T"])
    print(f"{name}: {ctas:4d} CTAs x {waves} waves: {best:8.2f} ms  {per:8.1f} us per chain-step per CTA  "
          f"{w['C'] * w['T'] / best * 1e3 / 1e6:7.3f} M chain-steps/s", flush=True)
    ops.reset_options()
