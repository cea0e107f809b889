"""Quick device-timed throughput of the SG workloads of bench.py (one process, several workloads / engine switches)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from pbnn_b200 import _lib as plib
from pbnn_b200.ops import FlatSpec, default_ops

ops = default_ops()
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["sgld_cfg3", "psgld_cfg3", "sghmc_cfg3", "adam_cfg4", "sgld_cfg5"]
envs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["", "no_tc2=1"]  # library options, "a=1+b=2"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
over = dict(kv.split("=") for kv in sys.argv[4].split(",")) if len(sys.argv) > 4 else {}  # e.g. C=4096,T=500
for name in names:
    w = dict(bench.WORKLOADS[name], name=name)
    w.update({k: int(v) for k, v in over.items()})
    fs = FlatSpec(tuple(w["layers"]), w["act"], 0.1)
    Xh, yh, Xth, keys, th0 = bench.make_inputs(w, 0, 1, ops)
    X, y = (ops._as(a, torch.float32) for a in (Xh, yh))
    step = bench.gpu_step_fn(ops, w, fs, X, y, th0, keys)
    op = {"hmc": plib.OP_HMC, "sgld": plib.OP_SGLD, "psgld": plib.OP_PSGLD, "sghmc": plib.OP_SGHMC, "adam": plib.OP_ADAM}[w["algo"]]
    for env in envs:
        for kv in filter(None, env.split("+")):
            k, v = kv.split("=")
            ops.set_option(k, int(v))
        eng = ops.lib.pbnn_plan_engine(fs.c_struct(), op, w["B"], w["C"])
        r = step(); torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r = step(); b.record(); torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        best = min(ms)
        fin = bool(torch.isfinite(r["theta"]).all().item())
        print(f"{name:12s} opt[{env}] engine={eng} C={w['C']} T={w['T']} L={w['L']}: {best:9.2f} ms  "
              f"{w['C'] * w['T'] / best * 1e3 / 1e6:8.3f} M chain-steps/s  finite={fin}", flush=True)
        ops.reset_options()
