"""GPU check of the two-hidden-layer tcgen05 engine (pbnn_tc2.cuh) against the float64 oracle."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pbnn_oracle as o
from pbnn_b200 import _lib
from pbnn_b200.ops import FlatSpec, default_ops

ops = default_ops()
rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30))
npy = lambda t: t.detach().cpu().numpy()


def problem(layers, C, N, seed=1, scale=0.2):
    X, y = o.synthetic_regression(N, layers[0], seed)
    sp, fs = o.MlpSpec(tuple(layers), o.RELU, 0.1), FlatSpec(tuple(layers), 0, 0.1)
    keys = o.split(o.PRNGKey(seed + 100), C)
    th0 = (scale * np.random.default_rng(seed).standard_normal((C, sp.num_params))).astype(np.float32)
    return X, y, sp, fs, keys, th0


def check_grad(layers, B, C=3, N=300, scale=0.2):
    X, y, sp, fs, keys, th0 = problem(layers, C, N, scale=scale)
    eng = ops.lib.pbnn_plan_engine(fs.c_struct(), _lib.OP_GRAD, B, C)
    idx = np.random.default_rng(2).integers(0, N, (C, B)).astype(np.int32)
    g, ll = ops.grad_estimator(fs, X, y, th0, idx)
    torch.cuda.synchronize()
    g64 = o.grad_estimator(sp, th0.astype(np.float64), X[idx].astype(np.float64), y[idx].astype(np.float64), N)
    ll64, _ = o.loglik_and_grad(sp, th0.astype(np.float64), X[idx].astype(np.float64), y[idx].astype(np.float64), 1.0)
    eg, el = rel(npy(g), g64), rel(npy(ll), ll64)
    # per-layer errors
    offs, parts = 0, []
    for i, oo in zip(layers[:-1], layers[1:]):
        nb, nw = oo, i * oo
        parts.append((rel(npy(g)[:, offs:offs + nb], g64[:, offs:offs + nb]), rel(npy(g)[:, offs + nb:offs + nb + nw], g64[:, offs + nb:offs + nb + nw])))
        offs += nb + nw
    print(f"grad {layers} B={B} C={C} engine={eng}: rel g {eg:.2e} ll {el:.2e}  per layer (bias, kernel): " +
          " ".join(f"({a:.1e},{b:.1e})" for a, b in parts), flush=True)
    return eg


def check_samplers(layers, B, C=3, N=300, scale=0.2):
    X, y, sp, fs, keys, th0 = problem(layers, C, N, scale=scale)
    T = 5
    ref = o.sgld(sp, X, y, th0, B, 1e-5, T, keys)
    out = ops.sgld(fs, X, y, th0, keys, B, 1e-5, T)
    print(f"  sgld  {layers} B={B}: trace rel {rel(npy(out['trace']), ref):.2e} flags {npy(out['flags']).tolist()}", flush=True)
    ref, (gref, Vref) = o.psgld(sp, X, y, th0, B, 1e-6, T, 0.95, keys, return_state=True)
    out = ops.psgld(fs, X, y, th0, keys, B, 1e-6, T, 0.95)
    print(f"  psgld {layers} B={B}: trace rel {rel(npy(out['trace']), ref):.2e} V rel {rel(npy(out['square_avg']), Vref):.2e}", flush=True)
    ref = o.sghmc(sp, X, y, th0, B, 1e-9, 1, 2, keys)
    out = ops.sghmc(fs, X, y, th0, keys, B, 1e-9, 1, 2)
    print(f"  sghmc {layers} B={B}: trace rel {rel(npy(out['trace']), ref):.2e}", flush=True)
    # Adam: 4 steps on explicit minibatches
    idx = np.random.default_rng(5).integers(0, N, (C, 4, B)).astype(np.int32)
    th, m, v, cnt = th0.copy(), np.zeros_like(th0), np.zeros_like(th0), 0
    for s_ in range(4):
        g = -o.grad_estimator(sp, th, X[idx[:, s_]], y[idx[:, s_]], N)
        th, m, v, cnt = o.adam_step(th, g.astype(np.float32), m, v, cnt, 1e-3)
    out = ops.adam_ensemble(fs, X, y, th0, idx, 1e-3)
    print(f"  adam  {layers} B={B}: theta rel {rel(npy(out['theta']), th):.2e} m rel {rel(npy(out['m']), m):.2e}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("grad", "all"):
        check_grad((9, 100, 100, 1), 128)
        check_grad((9, 100, 100, 1), 128, C=300)
        check_grad((8, 50, 50, 1), 32)
        check_grad((8, 64, 64, 1), 100)
        check_grad((20, 48, 48, 1), 64)
        check_grad((90, 256, 256, 1), 128, C=2, N=400, scale=0.05)
        check_grad((127, 256, 256, 1), 77, C=2, N=400, scale=0.05)
        check_grad((70, 160, 160, 1), 128, C=2, N=400, scale=0.05)
    if which in ("samplers", "all"):
        check_samplers((9, 100, 100, 1), 128)
        check_samplers((8, 50, 50, 1), 32)
        check_samplers((90, 256, 256, 1), 128, C=2, N=400, scale=0.03)
        ops.set_option("no_tc2", 1)
        print("  (FP32 engine on the same problem:)")
        check_samplers((90, 256, 256, 1), 128, C=2, N=400, scale=0.03)
        ops.reset_options()
