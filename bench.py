#!/usr/bin/env python
"""bench.py -- chain-steps/sec of the B200-native pbnn hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload hmc_cfg2] [--impl reference]

One *step* = one pass of the hot path over one batch of synthetic input: a full run of the named configuration's
job (default `hmc_cfg2`: BASELINE.json configs[1], 256 chains x 200 HMC iterations x L=20 full-batch leapfrogs on a
13-50-1 MLP, N=506).  `value` = chain-steps/s with inputs resident in HBM (CUDA events, max over ranks);
`e2e` = the same through the public sampler call with pinned HOST buffers (H2D of data/positions/keys and D2H of
the positions inside the timed region).  Chains shard over ranks with no data-path collective (weak scaling); the
only exchange is an all_gather of the predictive moments after the timed region.

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
    # BASELINE.json configs[2]
    "psgld_cfg3": dict(algo="psgld", layers=(9, 100, 100, 1), act=0, N=45730, B=128, C=1024, T=200, L=1, step=2e-8,
                       M=1024),
    "sgld_cfg3": dict(algo="sgld", layers=(9, 100, 100, 1), act=0, N=45730, B=128, C=1024, T=200, L=1, step=2e-8,
                      M=1024),
    "sghmc_cfg3": dict(algo="sghmc", layers=(9, 100, 100, 1), act=0, N=45730, B=128, C=1024, T=20, L=10, step=1e-9,
                       M=1024),
    # BASELINE.json configs[4], per-GPU share (1024 of the 8192 chains); state lives in the global workspace
    "sgld_cfg5": dict(algo="sgld", layers=(90, 256, 256, 1), act=0, N=515345, B=128, C=1024, T=20, L=1, step=1e-9,
                      M=4096),
    # BASELINE.json configs[3]
    "adam_cfg4": dict(algo="adam", layers=(8, 50, 50, 1), act=0, N=9568, B=32, C=512, T=299 * 2, L=1, step=1e-2,
                      M=1024),
}


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
                  alpha=0.95 if w["algo"] == "psgld" else 0.01, trace=False)
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
        return lambda: ops.sghmc(fs, X, y, th0, keys, w["B"], w["step"], w["T"], w["L"], keep_trace=False)
    if a == "adam":
        idx = torch.randint(0, w["N"], (w["C"], w["T"], w["B"]), dtype=torch.int32, device="cuda")
        return lambda: ops.adam_ensemble(fs, X, y, th0, idx, w["step"])
    raise ValueError(a)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="hmc_cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    base = {"metric": "chain-steps/sec", "unit": "chain-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (UCI-shaped regression, seeded; SURVEY.md 8d)",
            "config": {"workload": f"{args.workload}: {w['algo']} MLP {'-'.join(map(str, w['layers']))} relu, N={w['N']}, "
                                   f"batch={w['B']}, {w['C']} chains/GPU x {w['T']} iterations x L={w['L']} per step",
                       "chains_per_gpu": w["C"], "iterations_per_step": w["T"], "leapfrogs": w["L"],
                       "l2": "flushed (256 MiB write) before every timed step"}}

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
              "sample": f"each step = {n_chains} chains x {iters} iterations of {args.workload} (C restatement of the "
                        f"reference, OpenMP over chains on all host threads; JAX/BlackJAX not installable here)"}
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
    fs = FlatSpec(tuple(w["layers"]), w["act"], 0.1)
    Xh, yh, Xth, keys, th0 = make_inputs(w, rank, world, ops)
    X, y, Xt = (ops._as(a, torch.float32) for a in (Xh, yh, Xth))
    keysh, th0h = keys.cpu().numpy(), th0.cpu().numpy()
    step = gpu_step_fn(ops, w, fs, X, y, th0, keys)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K, W):
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
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    chain_steps = w["C"] * w["T"] * world
    value = args.steps * chain_steps / (ms_total * 1e-3)
    ms_per_step = ms_total / args.steps

    # end-to-end: public call with pinned host buffers; H2D of inputs and D2H of the result inside the timed region
    e2e = None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a.view(np.int32) if a.dtype == np.uint32 else a)).pin_memory()
        hX, hy, hk, ht = pin(Xh), pin(yh), pin(keysh), pin(th0h)
        P = fs.num_params
        res_shape = (w["C"], w["T"] if w["algo"] == "hmc" else 1, P)
        hout = torch.empty(res_shape, dtype=torch.float32).pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in (hX, hy, hk, ht))
        d2h = hout.numel() * 4

        def e2e_step():
            dX, dy = hX.to("cuda", non_blocking=True), hy.to("cuda", non_blocking=True)
            dk, dt_ = hk.to("cuda", non_blocking=True), ht.to("cuda", non_blocking=True)
            r = gpu_step_fn(ops, w, fs, dX, dy, dt_, dk)()
            src = r["trace"] if (w["algo"] == "hmc" and r.get("trace") is not None) else r["theta"].unsqueeze(1)
            hout.copy_(src, non_blocking=True)

        ms_serial = timed(e2e_step, args.steps, args.warmup)

        # the same calls as a stream pipeline: the D2H of step i (copy stream, double-buffered pinned output) overlaps
        # the kernel of step i+1; every step still uploads its inputs and downloads its whole result
        copy_stream = torch.cuda.Stream()
        houts = [hout, torch.empty(res_shape, dtype=torch.float32).pin_memory()]

        def pipelined(K):
            cur = torch.cuda.current_stream()
            for i in range(K):
                flush.fill_(1)
                dX, dy = hX.to("cuda", non_blocking=True), hy.to("cuda", non_blocking=True)
                dk, dt_ = hk.to("cuda", non_blocking=True), ht.to("cuda", non_blocking=True)
                r = gpu_step_fn(ops, w, fs, dX, dy, dt_, dk)()
                src = r["trace"] if (w["algo"] == "hmc" and r.get("trace") is not None) else r["theta"].unsqueeze(1)
                ev = torch.cuda.Event()
                ev.record(cur)
                copy_stream.wait_event(ev)
                with torch.cuda.stream(copy_stream):
                    houts[i % 2].copy_(src, non_blocking=True)
                src.record_stream(copy_stream)
            cur.wait_stream(copy_stream)

        pipelined(args.warmup)
        barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        pipelined(args.steps)
        eb.record()
        barrier()
        ms_e2e = ea.elapsed_time(eb)
        if world > 1:
            t = torch.tensor([ms_e2e], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t.item())
        e2e = {"value": args.steps * chain_steps / (ms_e2e * 1e-3), "unit": "chain-steps/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
               "serial_value": args.steps * chain_steps / (ms_serial * 1e-3),
               "api": "pbnn_b200.ops (C ABI) with pinned host buffers; positions returned to the host; value = steps "
                      "issued as a 2-stream pipeline (D2H of step i overlaps the kernel of step i+1), serial_value = "
                      "one step at a time"}

    # the single exchange step: predictive moments of each rank's final states, all_gather'ed (not in the timed region)
    last = step()
    final = last["theta"]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mom = ops.predict(fs, final, Xt, want_preds=False, want_moments=True)
    from pbnn_b200 import dist as pdist
    pm, pv, S_tot = pdist.pooled_moments(mom["sum"], mom["sumsq"], w["C"])  # all_gather over NCCL when world > 1
    torch.cuda.synchronize()
    exchange_ms = 1e3 * (time.perf_counter() - t0)

    # posterior predictive over stored samples (north_star subsystem 4), timed on its own: every stored position of the
    # HMC workload (C x T samples), else the final states, on the M test points
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

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        flops = step_flops(w)
        achieved = flops / (ms_per_step * 1e-3) / 1e12
        sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        peak_nominal = 148 * 128 * 2 * sm_max * 1e6 / 1e12
        trace_bytes = w["C"] * w["T"] * fs.num_params * 4 if w["algo"] == "hmc" else 0
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get(args.workload)
            if tr and not args.iters and not args.chains:
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
        except (OSError, KeyError, ValueError):
            pass
        from pbnn_b200 import _lib as plib
        op = {"hmc": plib.OP_HMC, "sgld": plib.OP_SGLD, "psgld": plib.OP_PSGLD, "sghmc": plib.OP_SGHMC,
              "adam": plib.OP_ADAM}[w["algo"]]
        engine = ops.lib.pbnn_plan_engine(fs.c_struct(), op, w["B"], w["C"])
        hbm = {"achieved": trace_bytes / (ms_per_step * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
               "frac": trace_bytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
               "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"}
        if engine == plib.ENGINE_TCGEN05:
            # the two GEMMs of every gradient run on the tensor cores as split-TF32 products:
            # achieved = ALGORITHMIC flops / time against the TF32 tensor peak (half the measured dense bf16 rate)
            tf32_peak = peaks.get("bf16_tflops", 2250.0 * 0.73) / 2
            ntile = (w["B"] + 127) // 128
            grads = w["C"] * w["T"] * (w["L"] + 1.0 / w["T"])
            executed = grads * ntile * (6 + 8) * (128 * 64 * 8 * 2)  # 6 forward + 8 backward MMAs 128x64x8 per tile
            roofline = {"bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                        "frac": achieved / tf32_peak, "traffic": traffic,
                        "peak_source": "TF32 dense tensor rate = measured cuBLAS bf16 (MEASURED_PEAKS.json burst) / 2"
                                       if peaks else "fallback: 0.73 x 2250 / 2",
                        "kernel": "pb_hmc_tc_kernel", "algorithmic_flops_per_launch": flops,
                        "executed_tensor_flops_per_launch": executed,
                        "executed_tensor_tflops": executed / (ms_per_step * 1e-3) / 1e12,
                        "fp32_fma_equivalent": {"achieved": achieved, "peak": peak_nominal, "frac": achieved / peak_nominal,
                                                "note": "the same algorithmic flops against the FP32-FMA peak the "
                                                        "scalar engines are bound by (148 SM x 128 lanes x 2 x clock)"},
                        "note": "latency-bound: per gradient a dependent chain of forward MMA -> epilogue -> backward "
                                "MMA -> readout; two CTAs per SM overlap each other's waits; see profiles/ and DESIGN.md",
                        "hbm": hbm}
        else:
            roofline = {"bound": "fma", "achieved": achieved, "peak": peak_nominal, "unit": "TFLOP/s",
                        "frac": achieved / peak_nominal, "traffic": traffic,
                        "peak_source": f"FP32 FMA: 148 SM x 128 lanes x 2 x {sm_max:.0f} MHz (the kernel is FP32-FMA "
                                       f"bound, neither HBM nor tensor; DESIGN.md)",
                        "kernel": "pb_hmc_kernel" if w["algo"] == "hmc" else "pb_sg_kernel",
                        "algorithmic_flops_per_launch": flops, "hbm": hbm}
        roofline["engine"] = ["tiled", "fused_h1", "tcgen05", "tiled_global_state"][engine] if 0 <= engine < 4 else engine
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
            roofline["frac_of_measured_fp32"] = achieved / best
        except Exception as exc:  # noqa: BLE001
            roofline["peak_measured_fp32_probe"] = None
        out = dict(base, value=value, ms_per_step=ms_per_step, clocks=clocks, roofline=roofline, gpu_launches=args.steps,
                   predictive=predictive,
                   exchange={"what": "predictive moments (sum, sumsq over chains) all_gather", "ms": exchange_ms,
                             "pred_mean_abs": float(pm.abs().mean()), "pred_var_mean": float(pv.mean())},
                   chains_finite=bool(torch.isfinite(final).all().item()))
        if w["algo"] == "hmc":
            out["hmc_accept_rate"] = float(last["accept_count"].float().mean().item() / w["T"])
        if e2e:
            out["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(w)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
