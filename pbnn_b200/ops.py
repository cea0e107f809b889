"""Flat-array operator layer over the C ABI: torch CUDA tensors in, torch CUDA tensors out.

PyTorch is used only for device memory and streams.  Every method launches the hand-written
sm_100a kernels of ``libpbnn_b200.so``; there is no eager / CPU path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import ACT_RELU, ACT_TANH, MB_PER_STEP, MB_REFERENCE_FIXED, check, make_spec

_MB_MODES = {"reference_fixed": MB_REFERENCE_FIXED, "per_step": MB_PER_STEP}

# SGHMC momentum resampling p = (mu0 + mu1 sqrt(step)) + (sig0 + sig1 sqrt(step)) z  (include/pbnn_b200.h):
# blackjax.sghmc is not vendored with the reference, so which of the three the pinned 1.2.5 implements cannot be read
# here; the default is the restatement of SURVEY.md 2b.
MOMENTUM_RESAMPLING = {
    "unit": (0.0, 0.0, 1.0, 0.0),             # p ~ N(0, 1): generate_gaussian_noise(rng_key, position)
    "sigma_sqrt_step": (0.0, 0.0, 0.0, 1.0),  # p ~ N(0, step): sigma=sqrt(step_size) (Chen et al. 2014)
    "mu_sqrt_step": (0.0, 1.0, 1.0, 0.0),     # p ~ N(sqrt(step), 1): sqrt(step_size) landing in the `mu` slot
}


def momentum_coefficients(momentum):
    if isinstance(momentum, str):
        if momentum not in MOMENTUM_RESAMPLING:
            raise ValueError(f"momentum_resampling must be one of {sorted(MOMENTUM_RESAMPLING)} or 4 coefficients")
        return MOMENTUM_RESAMPLING[momentum]
    coef = tuple(float(v) for v in momentum)
    if len(coef) != 4:
        raise ValueError("momentum_resampling: expected (mu, mu_sqrt_step, sigma, sigma_sqrt_step)")
    return coef


@dataclass(frozen=True)
class FlatSpec:
    """Widths + densities of the MLP log-posterior (the ``pbnn_mlp_spec`` struct)."""

    layers: tuple
    activation: int = ACT_RELU
    noise_level: float = 0.1
    prior_scale: float = 1.0
    likelihood: int = 0  # _lib.LIK_HOMOSCEDASTIC | _lib.LIK_HETEROSCEDASTIC

    @property
    def target_width(self) -> int:
        return 1 if self.likelihood == _lib.LIK_HETEROSCEDASTIC else self.layers[-1]

    @property
    def num_params(self) -> int:
        return sum(o * (i + 1) for i, o in zip(self.layers[:-1], self.layers[1:]))

    def c_struct(self):
        return make_spec(self.layers, self.activation, self.noise_level, self.prior_scale, self.likelihood)


def trace_len(T: int, thin: int, burnin: int) -> int:
    if thin < 1 or burnin < 0:
        raise ValueError("need thin >= 1 and burnin >= 0")
    return (T - burnin + thin - 1) // thin if T > burnin else 0


class Ops:
    """Binding of one loaded library.  Array handling is isolated in ``_ptr`` / ``_empty`` / ``_as`` so that the
    test-suite's kernel-logic emulator (tests/emu, host pointers) can reuse the marshalling; the product
    instance (``default_ops()``) only accepts CUDA tensors."""

    def __init__(self, lib=None):
        self.lib = lib if lib is not None else _lib.load_library()

    # ---- explicit engine-selection options (the library reads nothing from the environment) ------------
    def set_option(self, name: str, value: int) -> None:
        check(self.lib, self.lib.pbnn_set_option(name.encode(), int(value)), "set_option")

    def get_option(self, name: str):
        v = _lib.c_int(0)
        rc = self.lib.pbnn_get_option(name.encode(), v)
        if rc < 0:
            check(self.lib, rc, "get_option")
        return None if rc == 1 else int(v.value)

    def reset_options(self) -> None:
        self.lib.pbnn_reset_options()

    # ---- array handling (CUDA) -----------------------------------------------------------------
    def _device(self):
        if not torch.cuda.is_available():
            raise RuntimeError("pbnn_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())

    def _as(self, a, dtype):
        """Move to the device as a contiguous tensor of ``dtype`` (torch dtype)."""
        if a is None:
            return None
        if isinstance(a, torch.Tensor):
            t = a
        else:
            arr = np.asarray(a)
            if arr.dtype == np.uint32:
                arr = arr.view(np.int32)
            t = torch.from_numpy(np.ascontiguousarray(arr))
        if t.dtype == torch.uint32:
            t = t.view(torch.int32)
        return t.to(device=self._device(), dtype=dtype, non_blocking=True).contiguous()

    def _empty(self, shape, dtype):
        return torch.empty(tuple(int(s) for s in shape), dtype=dtype, device=self._device())

    def _zeros(self, shape, dtype):
        return torch.zeros(tuple(int(s) for s in shape), dtype=dtype, device=self._device())

    def _ptr(self, t):
        return None if t is None else t.data_ptr()

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    f32 = torch.float32
    i32 = torch.int32

    # ---- helpers -------------------------------------------------------------------------------
    def _keys(self, keys):
        k = self._as(keys, self.i32)
        if k.ndim == 1:
            k = k.reshape(1, 2)
        if k.ndim != 2 or k.shape[1] != 2:
            raise ValueError("rng keys must have shape (2,) or (C, 2) (legacy uint32 threefry keys)")
        return k.contiguous()

    def _steps(self, step_size, T):
        if callable(step_size):
            vals = [float(step_size(t)) for t in range(1, T + 1)]  # scheduled_sgld: schedule_fn(1..T)
        elif np.ndim(step_size) == 0:
            vals = [float(step_size)]
        else:
            vals = [float(v) for v in np.asarray(step_size).reshape(-1)]
        if len(vals) not in (1, T):
            raise ValueError("step_size must be a scalar, a schedule callable or T values")
        return self._as(np.asarray(vals, dtype=np.float32), self.f32), len(vals)

    def _theta(self, theta, spec: FlatSpec, C: int):
        th = self._as(theta, self.f32)
        th = th.reshape(-1, th.shape[-1]) if th.ndim > 1 else th.reshape(1, -1)
        if th.shape != (C, spec.num_params):
            raise ValueError(f"positions must have shape ({C}, {spec.num_params}), got {tuple(th.shape)}")
        return th.clone()

    def _data(self, X, y, spec: FlatSpec):
        X = self._as(X, self.f32)
        y = self._as(y, self.f32)
        if X.ndim != 2 or X.shape[1] != spec.layers[0]:
            raise ValueError(f"X must be (N, {spec.layers[0]}), got {tuple(X.shape)}")
        if y.ndim == 1:
            y = y.reshape(-1, 1)
        if y.shape != (X.shape[0], spec.target_width):
            raise ValueError(f"y must be (N, {spec.target_width}), got {tuple(y.shape)}")
        return X, y

    def _idx(self, idx, C, B):
        if idx is None:
            return None, 0
        ix = self._as(idx, self.i32)
        if ix.ndim == 2:
            ix = ix.reshape(C, 1, B)
        if ix.ndim != 3 or ix.shape[0] != C or ix.shape[2] != B:
            raise ValueError(f"idx must be (C, steps, B) = ({C}, *, {B}), got {tuple(ix.shape)}")
        return ix.contiguous(), int(ix.shape[1])

    # ---- jax.random primitives -----------------------------------------------------------------
    def split(self, keys, n: int = 2):
        k = self._keys(keys)
        out = self._empty((k.shape[0], n, 2), self.i32)
        check(self.lib, self.lib.pbnn_threefry_split(self._ptr(k), k.shape[0], n, self._ptr(out), self._stream()), "split")
        return out

    def fold_in(self, keys, data: int):
        k = self._keys(keys)
        out = self._empty((k.shape[0], 2), self.i32)
        check(self.lib, self.lib.pbnn_threefry_fold_in(self._ptr(k), k.shape[0], int(data) & 0xFFFFFFFF, self._ptr(out),
                                                       self._stream()), "fold_in")
        return out

    def bits(self, keys, n: int):
        k = self._keys(keys)
        out = self._empty((k.shape[0], n), self.i32)
        check(self.lib, self.lib.pbnn_threefry_bits(self._ptr(k), k.shape[0], n, self._ptr(out), self._stream()), "bits")
        return out

    def uniform(self, keys, n: int):
        k = self._keys(keys)
        out = self._empty((k.shape[0], n), self.f32)
        check(self.lib, self.lib.pbnn_threefry_uniform(self._ptr(k), k.shape[0], n, self._ptr(out), self._stream()),
              "uniform")
        return out

    def normal(self, keys, n: int):
        k = self._keys(keys)
        out = self._empty((k.shape[0], n), self.f32)
        check(self.lib, self.lib.pbnn_threefry_normal(self._ptr(k), k.shape[0], n, self._ptr(out), self._stream()),
              "normal")
        return out

    def minibatch_indices(self, keys, draw: int, B: int, N: int):
        k = self._keys(keys)
        out = self._empty((k.shape[0], B), self.i32)
        check(self.lib, self.lib.pbnn_minibatch_indices(self._ptr(k), k.shape[0], draw, B, N, self._ptr(out),
                                                        self._stream()), "minibatch_indices")
        return out

    def permutation(self, keys, N: int):
        k = self._keys(keys)
        C = k.shape[0]
        out = self._empty((C, N), self.i32)
        ws = self._empty((3 * C * N * 2,), self.i32)  # 3*C*N uint64
        check(self.lib, self.lib.pbnn_permutation(self._ptr(k), C, N, self._ptr(out), self._ptr(ws), self._stream()),
              "permutation")
        return out

    def _workspace(self, spec, op, rows, C):
        """Device workspace for models whose state does not fit shared memory (None when not needed)."""
        need = self.lib.pbnn_workspace_bytes(spec.c_struct(), op, int(rows), int(C))
        if need < 0:
            check(self.lib, int(need), "workspace_bytes")
        if need == 0:
            return None, 0
        return self._empty(((need + 3) // 4,), self.f32), int(need)

    # ---- samplers ------------------------------------------------------------------------------
    def _sg_common(self, spec, X, y, theta, keys, B, step_size, T, idx, thin, burnin, keep_trace):
        keys = self._keys(keys)
        C = keys.shape[0]
        X, y = self._data(X, y, spec)
        th = self._theta(theta, spec, C)
        ix, idx_steps = self._idx(idx, C, B)
        steps, nsteps = self._steps(step_size, T)
        T_out = trace_len(T, thin, burnin)
        trace = self._empty((C, T_out, spec.num_params), self.f32) if keep_trace else None
        flags = self._zeros((C,), self.i32)
        return keys, C, X, y, th, ix, idx_steps, steps, nsteps, trace, flags

    def _cv(self, cv, spec, C):
        if cv is None:
            return None, None
        return self._theta(cv[0], spec, C), self._theta(cv[1], spec, C)

    def sgld(self, spec: FlatSpec, X, y, theta, keys, B: int, step_size, T: int, *, idx=None,
             minibatch_mode="reference_fixed", t0=0, thin=1, burnin=0, keep_trace=True, cv=None, temperature=1.0):
        keys, C, X, y, th, ix, idx_steps, steps, nsteps, trace, flags = self._sg_common(
            spec, X, y, theta, keys, B, step_size, T, idx, thin, burnin, keep_trace)
        cs = spec.c_struct()
        ws, wsb = self._workspace(spec, _lib.OP_SGLD, B, C)
        cvc, cve = self._cv(cv, spec, C)
        rc = self.lib.pbnn_sgld_run(cs, self._ptr(X), self._ptr(y), X.shape[0], self._ptr(ix), idx_steps,
                                    _MB_MODES[minibatch_mode], B, self._ptr(keys), self._ptr(th), self._ptr(steps),
                                    nsteps, t0, T, C, thin, burnin, self._ptr(trace), self._ptr(flags), self._ptr(cvc),
                                    self._ptr(cve), float(temperature), self._ptr(ws), wsb, self._stream())
        check(self.lib, rc, "sgld")
        return {"theta": th, "trace": trace, "flags": flags}

    def psgld(self, spec: FlatSpec, X, y, theta, keys, B: int, step_size, T: int, alpha: float, *, idx=None,
              minibatch_mode="reference_fixed", t0=0, thin=1, burnin=0, keep_trace=True, state=None):
        keys, C, X, y, th, ix, idx_steps, steps, nsteps, trace, flags = self._sg_common(
            spec, X, y, theta, keys, B, step_size, T, idx, thin, burnin, keep_trace)
        if state is None:
            g, V, has = self._zeros(th.shape, self.f32), self._zeros(th.shape, self.f32), 0
        else:
            g, V, has = self._theta(state[0], spec, C), self._theta(state[1], spec, C), 1
        cs = spec.c_struct()
        ws, wsb = self._workspace(spec, _lib.OP_PSGLD, B, C)
        rc = self.lib.pbnn_psgld_run(cs, self._ptr(X), self._ptr(y), X.shape[0], self._ptr(ix), idx_steps,
                                     _MB_MODES[minibatch_mode], B, self._ptr(keys), self._ptr(th), self._ptr(g),
                                     self._ptr(V), has, float(alpha), self._ptr(steps), nsteps, t0, T, C, thin, burnin,
                                     self._ptr(trace), self._ptr(flags), self._ptr(ws), wsb, self._stream())
        check(self.lib, rc, "psgld")
        return {"theta": th, "trace": trace, "flags": flags, "grad": g, "square_avg": V}

    def sghmc(self, spec: FlatSpec, X, y, theta, keys, B: int, step_size, T: int, L: int, *, alpha=0.01, beta=0.0,
              idx=None, minibatch_mode="reference_fixed", t0=0, thin=1, burnin=0, keep_trace=True, cv=None,
              temperature=1.0, momentum_resampling="unit"):
        mom = momentum_coefficients(momentum_resampling)
        keys, C, X, y, th, ix, idx_steps, steps, nsteps, trace, flags = self._sg_common(
            spec, X, y, theta, keys, B, step_size, T, idx, thin, burnin, keep_trace)
        cs = spec.c_struct()
        ws, wsb = self._workspace(spec, _lib.OP_SGHMC, B, C)
        cvc, cve = self._cv(cv, spec, C)
        rc = self.lib.pbnn_sghmc_run(cs, self._ptr(X), self._ptr(y), X.shape[0], self._ptr(ix), idx_steps,
                                     _MB_MODES[minibatch_mode], B, self._ptr(keys), self._ptr(th), int(L), float(alpha),
                                     float(beta), mom[0], mom[1], mom[2], mom[3], self._ptr(steps), nsteps, t0, T, C,
                                     thin, burnin, self._ptr(trace),
                                     self._ptr(flags), self._ptr(cvc), self._ptr(cve), float(temperature),
                                     self._ptr(ws), wsb, self._stream())
        check(self.lib, rc, "sghmc")
        return {"theta": th, "trace": trace, "flags": flags}

    def hmc(self, spec: FlatSpec, X, y, theta, keys, num_samples: int, step_size: float, inv_mass, L: int, *, t0=0,
            thin=1, burnin=0, keep_trace=True, info=False):
        keys = self._keys(keys)
        C = keys.shape[0]
        X, y = self._data(X, y, spec)
        th = self._theta(theta, spec, C)
        im = self._as(inv_mass, self.f32).reshape(-1)
        if im.shape[0] != spec.num_params:
            raise ValueError(f"inverse_mass_matrix must be the flat diagonal of length {spec.num_params}")
        S_out = trace_len(num_samples, thin, burnin)
        trace = self._empty((C, S_out, spec.num_params), self.f32) if keep_trace else None
        logp = self._empty((C,), self.f32)
        acc = self._zeros((C,), self.i32)
        flags = self._zeros((C,), self.i32)
        idelta = self._empty((C, num_samples), self.f32) if info else None
        iacc = self._empty((C, num_samples), self.i32) if info else None
        cs = spec.c_struct()
        ws, wsb = self._workspace(spec, _lib.OP_HMC, X.shape[0], C)
        rc = self.lib.pbnn_hmc_run(cs, self._ptr(X), self._ptr(y), X.shape[0], self._ptr(keys), self._ptr(th),
                                   self._ptr(im), float(step_size), int(L), t0, num_samples, C, thin, burnin,
                                   self._ptr(trace), self._ptr(logp), self._ptr(acc), self._ptr(idelta), self._ptr(iacc),
                                   self._ptr(flags), self._ptr(ws), wsb, self._stream())
        check(self.lib, rc, "hmc")
        if wsb > 0 and bool((flags & 2).any().item()):  # balanced schedule only: a hand-over timed out (segment skipped)
            raise _lib.PbnnError("hmc: a segment of the balanced tcgen05 schedule timed out waiting for its "
                                 "predecessor; the affected chains have bit 1 set in `flags` and are incomplete")
        return {"theta": th, "trace": trace, "flags": flags, "logp": logp, "accept_count": acc, "delta": idelta,
                "accept": iacc}

    def adam_ensemble(self, spec: FlatSpec, X, y, theta, idx, lr: float, *, m=None, v=None, count0=0, b1=0.9,
                      b2=0.999, eps=1e-8):
        th0 = self._as(theta, self.f32)
        C = th0.shape[0]
        X, y = self._data(X, y, spec)
        th = self._theta(th0, spec, C)
        ix = self._as(idx, self.i32)
        if ix.ndim != 3 or ix.shape[0] != C:
            raise ValueError("idx must be (C, n_steps, B)")
        n_steps, B = int(ix.shape[1]), int(ix.shape[2])
        m = self._zeros(th.shape, self.f32) if m is None else self._theta(m, spec, C)
        v = self._zeros(th.shape, self.f32) if v is None else self._theta(v, spec, C)
        flags = self._zeros((C,), self.i32)
        cs = spec.c_struct()
        ws, wsb = self._workspace(spec, _lib.OP_ADAM, B, C)
        rc = self.lib.pbnn_adam_ensemble_run(cs, self._ptr(X), self._ptr(y), X.shape[0], self._ptr(ix.contiguous()),
                                             n_steps, B, self._ptr(th), self._ptr(m), self._ptr(v), int(count0),
                                             float(lr), float(b1), float(b2), float(eps), C, self._ptr(flags),
                                             self._ptr(ws), wsb, self._stream())
        check(self.lib, rc, "adam_ensemble")
        return {"theta": th, "m": m, "v": v, "flags": flags, "count": count0 + n_steps}

    def grad_estimator(self, spec: FlatSpec, X, y, theta, idx):
        th = self._as(theta, self.f32)
        th = th.reshape(-1, th.shape[-1])
        C = th.shape[0]
        X, y = self._data(X, y, spec)
        ix = self._as(idx, self.i32).reshape(C, -1).contiguous()
        B = int(ix.shape[1])
        g = self._empty(th.shape, self.f32)
        ll = self._empty((C,), self.f32)
        cs = spec.c_struct()
        ws, wsb = self._workspace(spec, _lib.OP_GRAD, B, C)
        rc = self.lib.pbnn_grad_estimator(cs, self._ptr(X), self._ptr(y), X.shape[0], self._ptr(ix), B,
                                          self._ptr(th.contiguous()), C, self._ptr(g), self._ptr(ll), self._ptr(ws), wsb,
                                          self._stream())
        check(self.lib, rc, "grad_estimator")
        return g, ll

    def logposterior_grad(self, spec: FlatSpec, X, y, theta):
        """(logp [C], grad [C,P]) of the full-batch log-posterior (HMC target; centring gradient of control variates)."""
        th = self._as(theta, self.f32)
        th = th.reshape(-1, th.shape[-1]).contiguous()
        C = th.shape[0]
        X, y = self._data(X, y, spec)
        g = self._empty(th.shape, self.f32)
        lp = self._empty((C,), self.f32)
        ws, wsb = self._workspace(spec, _lib.OP_HMC, X.shape[0], C)
        rc = self.lib.pbnn_logposterior_grad(spec.c_struct(), self._ptr(X), self._ptr(y), X.shape[0], self._ptr(th), C,
                                             self._ptr(g), self._ptr(lp), self._ptr(ws), wsb, self._stream())
        check(self.lib, rc, "logposterior_grad")
        return lp, g

    def predict(self, spec: FlatSpec, samples, X_test, *, want_preds=True, want_moments=False):
        smp = self._as(samples, self.f32)
        smp = smp.reshape(-1, smp.shape[-1]).contiguous()
        S = smp.shape[0]
        if smp.shape[1] != spec.num_params:
            raise ValueError(f"samples must be (S, {spec.num_params})")
        Xt = self._as(X_test, self.f32)
        if Xt.ndim != 2 or Xt.shape[1] != spec.layers[0]:
            raise ValueError(f"X_test must be (M, {spec.layers[0]})")
        M, s = Xt.shape[0], spec.layers[-1]
        preds = self._empty((S, M, s), self.f32) if want_preds else None
        ssum = ssq = ws = None
        if want_moments:
            ssum, ssq = self._empty((M, s), self.f32), self._empty((M, s), self.f32)
        cs = spec.c_struct()
        wsb = int(self.lib.pbnn_predict_workspace_bytes(cs, S, M, 1 if want_moments else 0))
        if wsb < 0:
            check(self.lib, wsb, "predict_workspace_bytes")
        if wsb > 0:
            ws = self._empty(((wsb + 3) // 4,), self.f32)
        rc = self.lib.pbnn_predict(cs, self._ptr(smp), S, self._ptr(Xt), M, self._ptr(preds), self._ptr(ssum),
                                   self._ptr(ssq), self._ptr(ws), wsb, self._stream())
        check(self.lib, rc, "predict")
        out = {"preds": preds}
        if want_moments:
            mean = ssum / S
            out.update(sum=ssum, sumsq=ssq, mean=mean, var=(ssq / S - mean * mean).clamp_min(0), count=S)
        return out

    def thinning(self, positions, size: int):
        """Indices (size,) selected by the greedy energy-distance thinning of ``positions`` (n, D) -- utils/misc.py:58-125."""
        pos = self._as(positions, self.f32)
        pos = pos.reshape(pos.shape[0], -1).contiguous()
        n, D = int(pos.shape[0]), int(pos.shape[1])
        out = self._empty((int(size),), self.i32)
        ws = self._empty((3 * n + 2064,), self.f32)
        rc = self.lib.pbnn_thinning(self._ptr(pos), n, D, int(size), self._ptr(out), self._ptr(ws), self._stream())
        check(self.lib, rc, "thinning")
        return out

    def plan(self, spec: FlatSpec, rows: int, n_state: int):
        out = (_lib.c_int32 * 4)()
        check(self.lib, self.lib.pbnn_plan_query(spec.c_struct(), rows, n_state, out), "plan_query")
        return {"threads": out[0], "batch_tile": out[1], "smem_bytes": out[2], "padded_params": out[3]}


_DEFAULT: Optional[Ops] = None


def set_default_ops(ops: Optional[Ops]) -> None:
    """Test hook: lets the CPU test-suite drive the host-side API through the kernel-logic emulator."""
    global _DEFAULT
    _DEFAULT = ops


def default_ops() -> Ops:
    """The product instance: CUDA library only; raises ImportError when it has not been built."""
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = Ops()
    return _DEFAULT
