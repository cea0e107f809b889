#!/usr/bin/env python
"""bench.py -- chain-steps/sec of the B200-native pbnn hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload hmc_cfg2] [--impl reference]

One *step* = one pass of the hot path over one batch of synthetic input: a full run of the named configuration's
job (default `hmc_cfg2`: BASELINE.json configs[1], 256 chains x 200 HMC iterations x L=20 full-batch leapfrogs on a
13-50-1 MLP, N=506).  `value` = chain-steps/s with inputs resident in HBM (CUDA events, max over ranks);
`e2e` = the same through the reference-signature call `pbnn_b200.mcmc.hmc` (pytrees in, `(positions, ravel_fn,
predict_fn)` out) with HOST inputs and the positions copied back to pinned host memory inside the timed region
(`e2e.ops_*`: the same through the flat C-ABI operator layer).  The metric is "SGLD & HMC": the default run also
times the other BASELINE configurations (`workloads`: sgld_cfg1 -- configs[0], one chain --, sgld_cfg1x, sgld_cfg3,
psgld_cfg3, sghmc_cfg3 -- configs[2] --, sgld_cfg5 -- configs[4], one GPU's share --, adam_cfg4 -- configs[3]) with their
own rooflines.
Chains shard over ranks with no data-path collective (weak scaling).  With --gpus N > 1 the run adds BASELINE
configs[4] (`sharded_cfg5`: 1024 SGLD chains per GPU of a 90-256-256-1 net -- 8192 chains on 8 GPUs -- keys
split(PRNGKey(0), C_total)[shard]), checks on the device that a shard's chains are bit-identical to the same chains
run alone, and times the one exchange step (all_gather of the predictive moments over NCCL) warm.

`--impl reference` times the CPU restatement of the reference (oracle/c, OpenMP over chains, all host threads) on
the same config -- JAX / BlackJAX are not installable here, see DESIGN.md.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# ------------------------------------------------------------------------------------------------ workloads
# algorithmic FLOPs per gradient evaluation: 6*B*W - 2*B*d*h1, W = sum in*out
WORKLOADS = {
    # BASELINE.json configs[1]
    "hmc_cfg2": dict(algo="hmc", layers=(13, 50, 1), act=0, N=506, B=506, C=256, T=200, L=20, step=5e-4, M=1024),
    # BASELINE.json configs[0] shape, many chains
    "sgld_cfg1": dict(algo="sgld", layers=(8, 50, 1), act=0, N=1030, B=32, C=1, T=100_000, L=1, step=1e-5, M=1024),
    "sgld_cfg1x": dict(algo="sgld", layers=(8, 50, 1), act=0, N=1030, B=32, C=2048, T=2000, L=1, step=1e-5, M=1024),
    # the same shape on the other samplers (warp-per-chain engine; not part of the default line)
    "psgld_cfg1x": dict(algo="psgld", layers=(8, 50, 1), act=0, N=1030, B=32, C=2048, T=2000, L=1, step=1e-6, M=1024),
    "sghmc_cfg1x": dict(algo="sghmc", layers=(8, 50, 1), act=0, N=1030, B=32, C=2048, T=200, L=10, step=1e-7, M=1024,
                        momentum="sigma_sqrt_step"),
    # BASELINE.json configs[2]
    "psgld_cfg3": dict(algo="psgld", layers=(9, 100, 100, 1), act=0, N=45730, B=128, C=1024, T=200, L=1, step=2e-8,
                       M=1024),
    "sgld_cfg3": dict(algo="sgld", layers=(9, 100, 100, 1), act=0, N=45730, B=128, C=1024, T=200, L=1, step=2e-8,
                      M=1024),
    # (momentum ~ N(0, step): the SGHMC-paper law; the "unit" law of SURVEY.md 2b random-walks to non-finite values)
    "sghmc_cfg3": dict(algo="sghmc", layers=(9, 100, 100, 1), act=0, N=45730, B=128, C=1024, T=20, L=10, step=3e-9,
                       M=1024, momentum="sigma_sqrt_step"),
    # BASELINE.json configs[4], per-GPU share (1024 of the 8192 chains); state lives in the global workspace
    "sgld_cfg5": dict(algo="sgld", layers=(90, 256, 256, 1), act=0, N=515345, B=128, C=1024, T=20, L=1, step=1e-9,
                      M=4096),
    # BASELINE.json configs[3]
    "adam_cfg4": dict(algo="adam", layers=(8, 50, 50, 1), act=0, N=9568, B=32, C=512, T=299 * 2, L=1, step=1e-2,
                      M=1024),
}


# SGHMC momentum law p = (mu0 + mu1 sqrt(step)) + (sig0 + sig1 sqrt(step)) z  (include/pbnn_b200.h)
MOMENTUM = {"unit": (0.0, 0.0, 1.0, 0.0), "sigma_sqrt_step": (0.0, 0.0, 0.0, 1.0), "mu_sqrt_step": (0.0, 1.0, 1.0, 0.0)}


def grad_flops(layers, B):
    W = sum(i * o for i, o in zip(layers[:-1], layers[1:]))
    return 6 * B * W - 2 * B * layers[0] * layers[1]


def num_params(layers):
    return sum(o * (i + 1) for i, o in zip(layers[:-1], layers[1:]))


def step_flops(w):
    """algorithmic FLOPs of one bench step (all chains), RNG integer work excluded."""
    per_chain_step = w["L"] * grad_flops(w["layers"], w["B"]) + 6 * num_params(w["layers"])
    return per_chain_step * w["C"] * w["T"]


def make_data(w):
    """Synthetic UCI-shaped regression data (seeded numpy; the same arrays feed the GPU arm and the CPU arms)."""
    N, M, d = w["N"], w["M"], w["layers"][0]
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N + M, d)).astype(np.float32)
    X = ((X - X[:N].mean(0)) / X[:N].std(0)).astype(np.float32)
    wv = rng.standard_normal(d).astype(np.float32)
    y = (np.tanh(X @ wv / np.sqrt(np.float32(d))) + 0.1 * rng.standard_normal(N + M)).astype(np.float32)[:, None]
    return X[:N], y[:N], X[N:]


def make_inputs(w, rank, world, ops):
    """Data + per-chain keys and starting points of this rank.  Keys are split(PRNGKey(0), C * world)[rank's slice] and
    theta0 = 0.01 * normal(key_c, P) (the Flax init law), both produced by the library's own threefry kernels."""
    X, y, Xt = make_data(w)
    P = num_params(w["layers"])
    root = np.array([[0, 0]], np.uint32)  # PRNGKey(0)
    keys_all = ops.split(root, w["C"] * world)[0]
    keys = keys_all[rank * w["C"]:(rank + 1) * w["C"]].contiguous()
    th0 = 0.01 * ops.normal(keys, P)
    return X, y, Xt, keys, th0


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(hi), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_run(w, n_chains, n_iters, seed_rank=0):
    """Run the C restatement on a bounded sample: n_chains chains x n_iters sampler iterations.  Returns seconds."""
    from oracle import c_oracle as co
    from oracle import pbnn_oracle as o
    X, y, _ = make_data(w)
    sp = o.MlpSpec(tuple(w["layers"]), w["act"], 0.1)
    keys = o.split(o.PRNGKey(0), n_chains)
    th0 = o.init_positions(sp, keys)
    t0 = time.perf_counter()
    if w["algo"] == "hmc":
        co.hmc_run(w["layers"], w["act"], 0.1, X, y, th0, keys, n_iters, w["step"], np.ones(sp.num_params, np.float32),
                   w["L"], trace=False)
    elif w["algo"] == "adam":
        idx = np.random.default_rng(0).integers(0, w["N"], (n_chains, n_iters, w["B"])).astype(np.int32)
        co.adam_run(w["layers"], w["act"], 0.1, X, y, th0, idx, w["step"])
    else:
        co.sg_run(w["layers"], w["act"], 0.1, w["algo"], X, y, th0, w["B"], w["step"], n_iters, keys, L=w["L"],
                  alpha=0.95 if w["algo"] == "psgld" else 0.01, trace=False,
                  momentum=MOMENTUM[w.get("momentum", "unit")])
    return time.perf_counter() - t0


def cpu_sample_plan(w, target_s):
    """Size a bounded sample of the workload for ~target_s seconds of CPU work on all host threads."""
    from oracle import c_oracle as co
    cores = co.use_all_cores()
    n_chains = min(w["C"], max(1, cores) * 2)
    probe_iters = max(1, min(w["T"], int(2e8 / max(1, w["L"] * grad_flops(w["layers"], w["B"])))))
    dt = cpu_run(w, n_chains, probe_iters)
    rate = n_chains * probe_iters / dt  # chain-steps/s
    iters = int(max(probe_iters, min(w["T"], target_s * rate / n_chains)))
    return cores, n_chains, iters


def cpu_baseline(w, target_s=12.0):
    cores, n_chains, iters = cpu_sample_plan(w, target_s)
    dt = cpu_run(w, n_chains, iters)
    return {"value": n_chains * iters / dt, "unit": "chain-steps/s", "cores": cores, "kind": "port",
            "sample": f"{n_chains} chains x {iters} iterations of {w['name']} (C restatement of the reference, OpenMP over "
                      f"chains; JAX/BlackJAX not installable)", "seconds": dt}


# ------------------------------------------------------------------------------------------------ GPU arm
def gpu_step_fn(ops, w, fs, X, y, th0, keys, keep_trace=True):
    import torch
    a = w["algo"]
    if a == "hmc":
        minv = torch.ones(fs.num_params, device="cuda")
        return lambda: ops.hmc(fs, X, y, th0, keys, w["T"], w["step"], minv, w["L"], keep_trace=keep_trace)
    if a == "sgld":
        return lambda: ops.sgld(fs, X, y, th0, keys, w["B"], w["step"], w["T"], keep_trace=False)
    if a == "psgld":
        return lambda: ops.psgld(fs, X, y, th0, keys, w["B"], w["step"], w["T"], 0.95, keep_trace=False)
    if a == "sghmc":
        return lambda: ops.sghmc(fs, X, y, th0, keys, w["B"], w["step"], w["T"], w["L"], keep_trace=False,
                                 momentum_resampling=w.get("momentum", "unit"))
    if a == "adam":
        idx = torch.randint(0, w["N"], (w["C"], w["T"], w["B"]), dtype=torch.int32, device="cuda")
        return lambda: ops.adam_ensemble(fs, X, y, th0, idx, w["step"])
    raise ValueError(a)


ENGINE_NAMES = {0: "tiled_fp32", 1: "fused_h1_fp32", 2: "tcgen05_tf32x3", 3: "tiled_fp32_global_state",
                4: "tcgen05_bf16x3_mlp2", 5: "warp_per_chain_fp32", 6: "four_warps_per_member_fp32_mlp2", 7: "four_warps_per_chain_fp32"}
OPS = {"hmc": "OP_HMC", "sgld": "OP_SGLD", "psgld": "OP_PSGLD", "sghmc": "OP_SGHMC", "adam": "OP_ADAM"}


def workload_config(name, w):
    return {"workload": f"{name}: {w['algo']} MLP {'-'.join(map(str, w['layers']))} relu, N={w['N']}, batch={w['B']}, "
                        f"{w['C']} chains/GPU x {w['T']} iterations x L={w['L']} per step, step_size={w['step']:g}"
                        + (f", momentum={w['momentum']}" if "momentum" in w else ""),
            "chains_per_gpu": w["C"], "iterations_per_step": w["T"], "leapfrogs": w["L"], "step_size": w["step"],
            "l2": "flushed (256 MiB write) before every timed step"}


def roofline_for(ops, w, fs, ms_per_step, clocks, peaks, traffic=None):
    """Roofline object of one workload: algorithmic flops per launch / CUDA-event time against the pipe the dominant
    kernel runs on, plus the explanatory sub-rooflines (FP32-FMA equivalent, threefry issue)."""
    from pbnn_b200 import _lib as plib
    flops = step_flops(w)
    achieved = flops / (ms_per_step * 1e-3) / 1e12
    sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    fma_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    engine = ops.lib.pbnn_plan_engine(fs.c_struct(), getattr(plib, OPS[w["algo"]]), w["B"], w["C"])
    bf16_peak = peaks.get("bf16_tflops", 1590.0)
    src = "measured (MEASURED_PEAKS.json, burst)" if peaks else "fallback (B200_PROFILING.md)"
    fma_eq = {"achieved": achieved, "peak": fma_peak, "frac": achieved / fma_peak,
              "note": "the same algorithmic flops against the FP32-FMA peak (148 SM x 128 lanes x 2 x clock)"}
    # injected noise: one threefry2x32-20 block + erf_inv per parameter and noisy step, ~100 issue slots per warp-normal
    # (SASS count, pbnn_rng.cuh); peak = every scheduler issuing nothing else
    P = num_params(w["layers"])
    draws = {"hmc": 1, "sgld": 1, "psgld": 1, "sghmc": w["L"] + 1, "adam": 0}[w["algo"]]
    normals = w["C"] * w["T"] * P * draws / (ms_per_step * 1e-3)
    rng_peak = 148 * 4 * 32 * sm_max * 1e6 / 100.0
    rng = {"normals_per_s": normals, "peak": rng_peak, "frac": normals / rng_peak,
           "note": "bit-exact jax.random.normal draws/s against the issue-slot bound (~100 instructions per warp-normal)"}
    common = {"achieved": achieved, "unit": "TFLOP/s", "traffic": traffic, "algorithmic_flops_per_launch": flops,
              "engine": ENGINE_NAMES.get(engine, engine), "fp32_fma_equivalent": fma_eq, "rng_issue": rng}
    if engine == plib.ENGINE_TCGEN05:
        tf32_peak = bf16_peak / 2
        ntile = (w["B"] + 127) // 128
        grads = w["C"] * w["T"] * (w["L"] + 1.0 / w["T"])
        executed = grads * ntile * (6 + 8) * (128 * 64 * 8 * 2)  # 6 forward + 8 backward MMAs 128x64x8 per tile
        return dict(common, bound="tensor", peak=tf32_peak, frac=achieved / tf32_peak, kernel="pb_hmc_tc_kernel",
                    peak_source=f"TF32 dense tensor rate = cuBLAS bf16 {src} / 2",
                    executed_tensor_flops_per_launch=executed,
                    executed_tensor_tflops=executed / (ms_per_step * 1e-3) / 1e12,
                    note="latency-bound: per gradient a dependent chain forward MMA -> epilogue -> backward MMA -> "
                         "readout; two CTAs per SM overlap each other's waits (profiles/, DESIGN.md)")
    if engine == plib.ENGINE_TCGEN05_MLP2:
        B, (d, H) = 128, w["layers"][:2]
        r16 = lambda x: (x + 15) // 16 * 16
        Kx, Hp, Hy = r16(d + 1), r16(H), r16(H + 1)
        nmt = (Hp + 127) // 128
        per_grad = 2 * B * (6 * Kx * Hp + 6 * Hp * Hp + 6 * Hp * Hp + 6 * 128 * Hp + 3 * nmt * 128 * Hy)
        executed = per_grad * w["C"] * w["T"] * w["L"]
        return dict(common, bound="tensor", peak=bf16_peak, frac=achieved / bf16_peak, kernel="pb_sg_tc2_kernel",
                    peak_source=f"dense bf16 tensor rate, cuBLAS {src}",
                    executed_tensor_flops_per_launch=executed,
                    executed_tensor_tflops=executed / (ms_per_step * 1e-3) / 1e12,
                    note="bf16x3 products (6 passes, 3 against the exact relu masks); the step is bound by the per-"
                         "parameter noise + update work on the CUDA cores (rng_issue), not by the tensor pipe")
    return dict(common, bound="fma", peak=fma_peak, frac=achieved / fma_peak,
                kernel="pb_hmc_kernel" if w["algo"] == "hmc" else
                ("pb_sgw_kernel" if engine == plib.ENGINE_WARP_H1 else "pb_sgc_kernel" if engine == plib.ENGINE_CTA4_H1 else
                 "pb_adam_w2_kernel" if engine == plib.ENGINE_WARP_MLP2 else "pb_sg_kernel"),
                peak_source=f"FP32 FMA: 148 SM x 128 lanes x 2 x {sm_max:.0f} MHz")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="hmc_cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the SGLD workloads / the sharded cfg5 run")
    ap.add_argument("--iters", type=int, default=0, help="override iterations per step (profiling runs)")
    ap.add_argument("--chains", type=int, default=0, help="override chains per GPU (profiling runs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: W >= 3
    w = dict(WORKLOADS[args.workload], name=args.workload)
    if args.iters > 0:
        w["T"] = args.iters
    if args.chains > 0:
        w["C"] = args.chains
    if args.iters > 0 or args.chains > 0:
        args.no_extras = True
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    base = {"metric": "chain-steps/sec", "unit": "chain-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (UCI-shaped regression, seeded; SURVEY.md 8d)", "config": workload_config(args.workload, w)}

    # ---------------------------------------------------------------- reference arm (CPU restatement)
    if args.impl == "reference":
        if rank != 0:
            return
        cores, n_chains, iters = cpu_sample_plan(w, 8.0)
        for _ in range(min(args.warmup, 1)):
            cpu_run(w, n_chains, max(1, iters // 4))
        t = [cpu_run(w, n_chains, iters) for _ in range(args.steps)]
        total = sum(t)
        v = args.steps * n_chains * iters / total
        cb = {"value": v, "unit": "chain-steps/s", "cores": cores, "kind": "port",
              "sample": f"each step = {n_chains} chains (2 per host thread, NOT the workload's {w['C']}; per-chain work is "
                        f"identical) x {iters} iterations of {args.workload} (C restatement of the reference, OpenMP over "
                        f"chains on all host threads; JAX/BlackJAX not installable here)"}
        out = dict(base, impl="reference", value=v, ms_per_step=1e3 * total / args.steps, cpu_baseline=cb,
                   e2e={"value": v, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                   gpu_launches=0)
        print(json.dumps(out))
        return

    # ---------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist
    from pbnn_b200.ops import FlatSpec, default_ops

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ops = default_ops()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    launches = [0]  # kernels of libpbnn_b200.so launched inside timed regions (counted per operator call below)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K, W, kernels_per_call=1):
        for _ in range(W):
            fn()
        barrier()
        evs = []
        for _ in range(K):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        barrier()
        launches[0] += K * kernels_per_call
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    traffic_db = {}
    for fn_ in ("r2_traffic.json", "r1_traffic.json"):
        try:
            for k_, v_ in json.load(open(os.path.join(ROOT, "profiles", fn_))).items():
                traffic_db.setdefault(k_, v_)
        except (OSError, ValueError):
            pass

    def traffic_of(name):
        tr = traffic_db.get(name)
        if tr and not args.iters and not args.chains:
            return tr["dram_bytes_read"] + tr["dram_bytes_write"]
        return None

    # ---- the headline workload -----------------------------------------------------------------------
    fs = FlatSpec(tuple(w["layers"]), w["act"], 0.1)
    Xh, yh, Xth, keys, th0 = make_inputs(w, rank, world, ops)
    X, y, Xt = (ops._as(a, torch.float32) for a in (Xh, yh, Xth))
    keysh, th0h = keys.cpu().numpy(), th0.cpu().numpy()
    step = gpu_step_fn(ops, w, fs, X, y, th0, keys)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    chain_steps = w["C"] * w["T"] * world
    value = args.steps * chain_steps / (ms_total * 1e-3)
    ms_per_step = ms_total / args.steps

    # ---- end to end ------------------------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        P = fs.num_params
        res_shape = (w["C"], w["T"] if w["algo"] == "hmc" else 1, P)
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a.view(np.int32) if a.dtype == np.uint32 else a)).pin_memory()
        hX, hy, hk, ht = pin(Xh), pin(yh), pin(keysh), pin(th0h)
        hout = torch.empty(res_shape, dtype=torch.float32).pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in (hX, hy, hk, ht))
        d2h = hout.numel() * 4

        # (1) the flat operator layer (C ABI through pbnn_b200.ops), one step at a time and as a 2-stream pipeline
        def ops_step(dst=None, stream=None):
            dX, dy = hX.to("cuda", non_blocking=True), hy.to("cuda", non_blocking=True)
            dk, dt_ = hk.to("cuda", non_blocking=True), ht.to("cuda", non_blocking=True)
            r = gpu_step_fn(ops, w, fs, dX, dy, dt_, dk)()
            src = r["trace"] if (w["algo"] == "hmc" and r.get("trace") is not None) else r["theta"].unsqueeze(1)
            if stream is None:
                (dst if dst is not None else hout).copy_(src, non_blocking=True)
                return
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            stream.wait_event(ev)
            with torch.cuda.stream(stream):
                dst.copy_(src, non_blocking=True)
            src.record_stream(stream)

        ms_ops_serial = timed(ops_step, args.steps, args.warmup)

        copy_stream = torch.cuda.Stream()
        houts = [hout, torch.empty(res_shape, dtype=torch.float32).pin_memory()]

        def pipelined(fn_, K):
            for i in range(K):
                flush.fill_(1)
                fn_(houts[i % 2], copy_stream)
            torch.cuda.current_stream().wait_stream(copy_stream)

        def time_pipelined(fn_):
            pipelined(fn_, args.warmup)
            barrier()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            pipelined(fn_, args.steps)
            eb.record()
            barrier()
            launches[0] += args.steps
            ms = ea.elapsed_time(eb)
            if world > 1:
                t = torch.tensor([ms], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms

        ms_ops_pipe = time_pipelined(ops_step)
        to_rate = lambda ms: args.steps * chain_steps / (ms * 1e-3)
        e2e = {"unit": "chain-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ops_value": to_rate(ms_ops_pipe), "ops_serial_value": to_rate(ms_ops_serial)}

        # (2) the reference-signature call a pbnn user makes (src/pbnn/mcmc/hamiltonian.py:31-39 / langevin.py:21-33):
        # host arrays and Flax-layout pytrees in, (positions pytree, ravel_fn, predict_fn) out, positions to the host
        api_step = None
        if w["algo"] in ("hmc", "sgld"):
            from pbnn_b200 import logprobs, mcmc, pytree
            from pbnn_b200.models import MLP
            net = MLP.from_widths(w["layers"], "relu")
            lik = logprobs.GaussianMLPLikelihood(net, 0.1)
            prior = logprobs.StdNormalPrior()
            init_tree = pytree.unravel(ht, tuple(w["layers"]))  # pinned host leaves, leading chain axis

            def api_step(dst=None, stream=None):
                if w["algo"] == "hmc":
                    lp = logprobs.FullBatchLogPosterior(lik, prior, hX, hy)
                    pos, ravel_fn, _ = mcmc.hmc(lp, init_tree, w["T"], w["step"], np.ones(P, np.float32), w["L"], keysh)
                else:
                    pos, ravel_fn, _ = mcmc.sgld(hX, hy, lik, prior, init_tree, w["B"], w["step"], w["T"], keysh)
                src = ravel_fn(pos)
                if w["algo"] != "hmc":
                    src = src[:, -1:, :]
                src = src.contiguous()
                if stream is None:
                    (dst if dst is not None else hout).copy_(src, non_blocking=True)
                    return
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                stream.wait_event(ev)
                with torch.cuda.stream(stream):
                    dst.copy_(src, non_blocking=True)
                src.record_stream(stream)

        if api_step is not None:
            try:
                ms_api_serial = timed(api_step, args.steps, args.warmup)
                ms_api_pipe = time_pipelined(api_step)
                e2e.update(value=to_rate(ms_api_pipe), serial_value=to_rate(ms_api_serial),
                           ms_per_step=ms_api_pipe / args.steps,
                           api=f"pbnn_b200.mcmc.{w['algo']} (the reference's signature: pytrees in, (positions, ravel_fn, "
                               "predict_fn) out) with host inputs; positions copied to pinned host memory; value = steps "
                               "issued as a 2-stream pipeline (the D2H of step i overlaps the kernel of step i+1), "
                               "serial_value = one step at a time; ops_* = the same through the flat C-ABI operators")
            except Exception as exc:  # noqa: BLE001  (never lose the line over the wrapper layer)
                e2e["api_error"] = f"{type(exc).__name__}: {exc}"
        if "value" not in e2e:
            e2e.update(value=e2e["ops_value"], serial_value=e2e["ops_serial_value"], ms_per_step=ms_ops_pipe / args.steps,
                       api="pbnn_b200.ops (C ABI) with pinned host buffers; positions returned to the host; 2-stream "
                           "pipeline (value) / one step at a time (serial_value)")

    # ---- the single exchange step: predictive moments of each rank's final states, all_gather'ed -------
    from pbnn_b200 import dist as pdist
    last = step()
    final = last["theta"]
    mom = ops.predict(fs, final, Xt, want_preds=False, want_moments=True)
    pm, pv, S_tot = pdist.pooled_moments(mom["sum"], mom["sumsq"], w["C"])  # first call: communicator set-up included
    payload = torch.cat([mom["sum"].reshape(-1), mom["sumsq"].reshape(-1), torch.ones(1, device="cuda")])
    gathered = torch.empty(world * payload.numel(), device="cuda")

    def exchange():
        if world > 1:
            dist.all_gather_into_tensor(gathered, payload)
        else:
            gathered.copy_(payload)

    for _ in range(5):
        exchange()
    barrier()
    xa, xb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    xa.record()
    for _ in range(20):
        exchange()
    xb.record()
    barrier()
    exchange_us = xa.elapsed_time(xb) * 1e3 / 20
    exchange_info = {"what": "all_gather of the per-rank predictive moments (sum f, sum f^2, count), warm, CUDA events, "
                             "20 back-to-back calls" + (" over NCCL" if world > 1 else " (one rank: a device copy)"),
                     "us_per_exchange": exchange_us, "bytes_per_rank": int(payload.numel() * 4),
                     "pred_mean_abs": float(pm.abs().mean()), "pred_var_mean": float(pv.mean()), "samples_pooled": S_tot}

    # ---- posterior predictive over stored samples (north_star subsystem 4), timed on its own -----------
    psmp = last["trace"].reshape(-1, fs.num_params) if (w["algo"] == "hmc" and last.get("trace") is not None) else final
    ops.predict(fs, psmp, Xt, want_preds=False, want_moments=True)
    pa, pb_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pa.record()
    for _ in range(3):
        ops.predict(fs, psmp, Xt, want_preds=False, want_moments=True)
    pb_.record()
    torch.cuda.synchronize()
    pred_ms = pa.elapsed_time(pb_) / 3
    predictive = {"samples": int(psmp.shape[0]), "test_points": int(Xt.shape[0]), "ms": pred_ms,
                  "evals_per_s": psmp.shape[0] * Xt.shape[0] / (pred_ms * 1e-3),
                  "engine": "tcgen05" if (len(w["layers"]) == 3 and w["act"] == 0 and w["layers"][1] <= 64
                                          and w["layers"][0] + 1 <= 16 and not ops.get_option("no_tc")) else "fp32",
                  "note": "mean / variance over the samples for every test point; not part of `value`"}
    main_finite = bool(torch.isfinite(final).all().item())
    accept = float(last["accept_count"].float().mean().item() / w["T"]) if w["algo"] == "hmc" else None
    del last, final, psmp, step
    torch.cuda.empty_cache()

    # ---- the metric is "SGLD & HMC": the SGLD configurations, device-timed like the headline -------------
    extras = {}
    if not args.no_extras:
        for name in ("sgld_cfg1", "sgld_cfg1x", "sgld_cfg3", "psgld_cfg3", "sghmc_cfg3", "sgld_cfg5", "adam_cfg4"):
            if name == args.workload:
                continue
            we = dict(WORKLOADS[name], name=name)
            fse = FlatSpec(tuple(we["layers"]), we["act"], 0.1)
            Xe, ye, Xte_h, ke, the = make_inputs(we, rank, world, ops)
            Xd, yd = ops._as(Xe, torch.float32), ops._as(ye, torch.float32)
            fe = gpu_step_fn(ops, we, fse, Xd, yd, the, ke)
            ks = max(3, min(args.steps, 5))
            ms_e = timed(fe, ks, 3) / ks
            fin_theta = fe()["theta"]
            fin = bool(torch.isfinite(fin_theta).all().item())
            # posterior predictive moments of the final states on the workload's M test points (north_star subsystem 4)
            Xte = ops._as(Xte_h, torch.float32)
            ops.predict(fse, fin_theta, Xte, want_preds=False, want_moments=True)
            qa, qb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            qa.record()
            for _ in range(3):
                ops.predict(fse, fin_theta, Xte, want_preds=False, want_moments=True)
            qb.record()
            torch.cuda.synchronize()
            pms = qa.elapsed_time(qb) / 3
            Wsum = sum(i * o for i, o in zip(we["layers"][:-1], we["layers"][1:]))
            pred_e = {"samples": int(fin_theta.shape[0]), "test_points": int(Xte.shape[0]), "ms": pms,
                      "evals_per_s": fin_theta.shape[0] * Xte.shape[0] / (pms * 1e-3),
                      "algorithmic_tflops": 2.0 * fin_theta.shape[0] * Xte.shape[0] * Wsum / (pms * 1e-3) / 1e12}
            del Xte, fin_theta
            extras[name] = {"value": we["C"] * we["T"] * world / (ms_e * 1e-3),
                            "unit": "member-steps/s" if we["algo"] == "adam" else "chain-steps/s",
                            "ms_per_step": ms_e, "steps": ks, "warmup": 3, "config": workload_config(name, we),
                            "roofline": roofline_for(ops, we, fse, ms_e, clocks, peaks, traffic_of(name)),
                            "predictive": pred_e, "chains_finite": fin}
            del Xd, yd, fe, the, ke
            torch.cuda.empty_cache()

    # ---- N > 1: BASELINE configs[4], chains sharded through pbnn_b200.dist ----------------------------
    sharded = None
    if world > 1 and not args.no_extras:
        w5 = dict(WORKLOADS["sgld_cfg5"], name="sgld_cfg5")
        fs5 = FlatSpec(tuple(w5["layers"]), w5["act"], 0.1)
        C_total = w5["C"] * world
        X5h, y5h, X5t = make_data(w5)
        X5, y5, X5t = (ops._as(a, torch.float32) for a in (X5h, y5h, X5t))
        root = np.array([[0, 0]], np.uint32)
        k5 = pdist.shard_chain_keys(ops, root, C_total)            # split(PRNGKey(0), C_total)[lo:hi]
        lo, hi = pdist.shard_range(C_total)
        th5 = 0.01 * ops.normal(k5, fs5.num_params)
        f5 = lambda: ops.sgld(fs5, X5, y5, th5, k5, w5["B"], w5["step"], w5["T"], keep_trace=False)
        ms5 = timed(f5, 3, 3) / 3
        out5 = f5()["theta"]
        # sharding check on the device: every rank publishes its first 4 chains; rank r re-runs the first 4 chains of
        # rank (r + 1) % world from the global key table, alone, and compares bit for bit
        nchk = 4
        mine = out5[:nchk].contiguous()
        allfirst = torch.empty((world,) + tuple(mine.shape), device="cuda")
        dist.all_gather_into_tensor(allfirst, mine)
        peer = (rank + 1) % world
        kall = ops.split(root, C_total)[0]
        kp = kall[peer * w5["C"]:peer * w5["C"] + nchk].contiguous()
        thp = 0.01 * ops.normal(kp, fs5.num_params)
        solo = ops.sgld(fs5, X5, y5, thp, kp, w5["B"], w5["step"], w5["T"], keep_trace=False)["theta"]
        same = torch.tensor([1.0 if torch.equal(solo, allfirst[peer]) else 0.0], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        m5 = ops.predict(fs5, out5, X5t, want_preds=False, want_moments=True)
        za, zb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        za.record()
        for _ in range(3):
            ops.predict(fs5, out5, X5t, want_preds=False, want_moments=True)
        zb.record()
        torch.cuda.synchronize()
        pred5_ms = za.elapsed_time(zb) / 3
        pm5, pv5, S5 = pdist.pooled_moments(m5["sum"], m5["sumsq"], w5["C"])
        pay5 = torch.cat([m5["sum"].reshape(-1), m5["sumsq"].reshape(-1), torch.ones(1, device="cuda")])
        g5 = torch.empty(world * pay5.numel(), device="cuda")
        for _ in range(5):
            dist.all_gather_into_tensor(g5, pay5)
        barrier()
        ya, yb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ya.record()
        for _ in range(20):
            dist.all_gather_into_tensor(g5, pay5)
        yb.record()
        barrier()
        sharded = {"what": f"BASELINE configs[4]: {C_total} SGLD chains sharded over {world} GPUs "
                           f"({w5['C']} per GPU, rank r owns chains [{w5['C']}r, {w5['C']}(r+1)) of split(PRNGKey(0), "
                           f"{C_total})), MLP 90-256-256-1, N=515345, batch 128, {w5['T']} iterations per step",
                   "value": C_total * w5["T"] / (ms5 * 1e-3), "unit": "chain-steps/s", "ms_per_step": ms5,
                   "per_gpu_value": w5["C"] * w5["T"] / (ms5 * 1e-3), "scaling": "weak",
                   "sharding_check": {"chains_compared_per_rank": nchk, "ranks": world,
                                      "bit_identical_to_solo_run": bool(same.item() == 1.0),
                                      "how": "rank r re-ran the first 4 chains of rank r+1 alone from the global key "
                                             "table; torch.equal on the final states, MIN-reduced over ranks"},
                   "moments_allgather_us": ya.elapsed_time(yb) * 1e3 / 20, "moments_bytes_per_rank": int(pay5.numel() * 4),
                   "predictive_ms_per_rank": pred5_ms, "predictive_test_points": int(X5t.shape[0]),
                   "pooled_samples": S5, "pred_var_mean": float(pv5.mean()),
                   "roofline": roofline_for(ops, w5, fs5, ms5, clocks, peaks, traffic_of("sgld_cfg5"))}

    if rank == 0:
        roofline = roofline_for(ops, w, fs, ms_per_step, clocks, peaks, traffic_of(args.workload))
        trace_bytes = w["C"] * w["T"] * fs.num_params * 4 if w["algo"] == "hmc" else 0
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        roofline["hbm"] = {"achieved": trace_bytes / (ms_per_step * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                           "frac": trace_bytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
                           "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"}
        try:  # practical FP32 ceiling measured in the same run (scalar FFMA and packed FFMA2 probes)
            import ctypes
            ops.lib.pbnn_probe_fp32.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 2
            sink = torch.zeros(4, device="cuda")
            best = 0.0
            for mode in (0, 1):
                ops.lib.pbnn_probe_fp32(mode, 200, 148 * 8, sink.data_ptr(), None)
                ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ea.record()
                ops.lib.pbnn_probe_fp32(mode, 10000, 148 * 8, sink.data_ptr(), None)
                eb.record()
                torch.cuda.synchronize()
                best = max(best, 148 * 8 * 256 * 10000 * 64 / (ea.elapsed_time(eb) * 1e-3) / 1e12)
            roofline["peak_measured_fp32_probe"] = best
        except Exception:  # noqa: BLE001
            roofline["peak_measured_fp32_probe"] = None
        out = dict(base, value=value, ms_per_step=ms_per_step, clocks=clocks, roofline=roofline,
                   gpu_launches=launches[0], predictive=predictive, exchange=exchange_info, chains_finite=main_finite)
        if accept is not None:
            out["hmc_accept_rate"] = accept
        if e2e:
            out["e2e"] = e2e
        if extras:
            out["workloads"] = extras
        if sharded:
            out["sharded_cfg5"] = sharded
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(w)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
