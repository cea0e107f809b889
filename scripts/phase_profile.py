"""Per-phase cycle breakdown of CTA 0 using the -DPB_PROFILE build (libpbnn_b200_prof.so).
usage: python scripts/phase_profile.py [workload] [iters] [opt=1+opt=2]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PBNN_B200_LIB"] = os.path.join(ROOT, "pbnn_b200", "libpbnn_b200_prof.so")
import numpy as np, torch
import bench
from pbnn_b200.ops import FlatSpec, default_ops

name = sys.argv[1] if len(sys.argv) > 1 else "hmc_cfg2"
w = dict(bench.WORKLOADS[name], name=name)
w["T"] = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ops = default_ops()
OPTS = dict(kv.split("=") for kv in filter(None, (sys.argv[3] if len(sys.argv) > 3 else "").split("+")))
for k_, v_ in OPTS.items():
    ops.set_option(k_, int(v_))
fs = FlatSpec(tuple(w["layers"]), w["act"], 0.1)
Xh, yh, Xth, keys, th0 = bench.make_inputs(w, 0, 1, ops)
X, y = ops._as(Xh, torch.float32), ops._as(yh, torch.float32)
step = bench.gpu_step_fn(ops, w, fs, X, y, th0, keys)
out = (ctypes.c_ulonglong * 64)()
step(); ops.lib.pbnn_debug_profile(out)
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); step(); b.record(); torch.cuda.synchronize()
ops.lib.pbnn_debug_profile(out)
if "no_tc" not in OPTS and name.startswith("hmc"):
    tcn = {1: "tc: W1 split/stage", 2: "tc: fwd MMA wait + epilogue r0", 4: "tc: bwd MMA wait + epilogue r1+", 5: "tc: bwd issue (+next)", 3: "tc: bwd wait + readout A", 6: "tc: readout B (update)",
           20: "tc mark: fwd issue", 21: "tc mark: fwd wait", 22: "tc mark: epilogue math", 23: "tc mark: bwd issue",
           24: "tc mark: bwd wait", 25: "tc mark: proxy fence"}
else:
    tcn = {}
names = {11: "hmc momentum", 12: "leapfrog pre", 13: "leapfrog post", 14: "logp reduce", 15: "accept", 16: "copy/emit", 0: "other/elementwise", 1: "fwd gemm", 2: "last fwd A", 3: "last fwd B", 4: "rank1 bwd C1", 5: "rank1 bwd C2",
         6: "dW gemm", 7: "dH gemm", 8: "gpart reduce", 9: "post", 10: "gather rows"}
names.update(tcn)
if len(w["layers"]) == 4 and "no_tc2" not in OPTS and "no_tc" not in OPTS:
    names.update({1: "tc2: vectors/restage", 2: "tc2: Z1 wait (fwd1)", 3: "tc2: H1 chunks (fwd2)", 4: "tc2: Z2 wait",
                  5: "tc2: f, r, mask", 6: "tc2: w3*M2 chunks + E wait", 7: "tc2: dZ1 chunks", 8: "tc2: D1 wait",
                  9: "tc2: D1 drain+update", 10: "tc2: Y chunks + D2 drain+update (stage X at chain start)",
                  11: "tc2: w3/b3 + end barrier", 0: "other"})
tot = sum(out[i] for i in range(32))
print(f"{name}: T={w['T']}  kernel {a.elapsed_time(b):.3f} ms; CTA0 accounted {tot} cycles; plan {ops.plan(fs, w['B'], 6 if w['algo']=='hmc' else 3)}")
for i in range(32):
    if out[i]:
        print(f"  tag {i:2d} {names.get(i,''):18s} {out[i]:12d} cyc  {100*out[i]/tot:5.1f}%  phases {out[32+i]:8d}  avg {out[i]/max(1,out[32+i]):8.0f}")
