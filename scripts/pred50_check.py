"""Predictive for 8-50-50-1 (BASELINE configs[3]) on the two-hidden-layer tcgen05 engine (option tc2_min_h) vs the FP32
kernel: float64 torch reference on the device, timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from pbnn_b200.ops import FlatSpec, default_ops
ops = default_ops()
torch.manual_seed(0)
for layers, S, M in (((8, 50, 50, 1), 512, 4096), ((8, 50, 50, 1), 512, 1000), ((8, 50, 50, 1), 4096, 9568), ((5, 48, 48, 1), 100, 300), ((8, 33, 33, 1), 64, 200)):
    fs = FlatSpec(layers, 0, 0.1)
    smp = (0.3 * torch.randn(S, fs.num_params, device="cuda")).contiguous()
    Xt = torch.randn(M, layers[0], device="cuda")
    d, H = layers[0], layers[1]
    s64 = smp.double(); o = 0
    W1 = s64[:, o:o + d * H].view(S, d, H); o += d * H
    b1 = s64[:, o:o + H]; o += H
    W2 = s64[:, o:o + H * H].view(S, H, H); o += H * H
    b2 = s64[:, o:o + H]; o += H
    W3 = s64[:, o:o + H]; o += H
    b3 = s64[:, o]
    # FlatSpec ravel order: check against the library's own fp32 kernel below (no_tc2) -- both orders are tested
    for opts in ({"no_tc2": 1}, {}, {"tc2_min_h": 32}):
        ops.reset_options()
        for k_, v_ in opts.items():
            ops.set_option(k_, v_)
        for it in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = ops.predict(fs, smp, Xt, want_preds=True, want_moments=True)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        if "no_tc2" in opts:
            base = {k: v.double() for k, v in out.items() if torch.is_tensor(v)}
            print(f"{layers} S={S} M={M} {opts}: {dt*1e3:8.3f} ms  keys={list(out.keys())}", flush=True)
        else:
            errs = {k: float((out[k].double() - base[k]).norm() / base[k].norm()) for k in base}
            print(f"{layers} S={S} M={M} {opts}: {dt*1e3:8.3f} ms  rel-vs-fp32-kernel {errs}", flush=True)
ops.reset_options()
