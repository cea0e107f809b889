"""ctypes binding of the C restatement (oracle/c/pbnn_oracle.c) -- TEST INFRASTRUCTURE ONLY.
Used by tests (validated against pbnn_oracle.py) and by bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_float, c_int, c_int32, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libpbnn_oracle_c.so")


class SpecC(ctypes.Structure):
    _fields_ = [("nl", c_int), ("dims", c_int * 9), ("act", c_int), ("sigma", c_float), ("prior_scale", c_float)]


def build(force=False):
    src = os.path.join(_HERE, "c", "pbnn_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(src) > os.path.getmtime(LIB):
        subprocess.check_call(["make", "-s", "-B", "-C", os.path.join(_HERE, "c")])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        # -march=native objects do not travel between hosts: rebuild when the library cannot be loaded/run here
        if not os.path.exists(LIB):
            build()
        _lib = ctypes.CDLL(LIB)
        _lib.pbo_num_threads.restype = c_int
    return _lib


def _spec(layers, act, sigma, prior_scale=1.0):
    s = SpecC()
    s.nl = len(layers) - 1
    for i, v in enumerate(layers):
        s.dims[i] = int(v)
    s.act, s.sigma, s.prior_scale = int(act), float(sigma), float(prior_scale)
    return s


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads():
    return lib().pbo_num_threads()


def use_all_cores():
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline is meant to use every host core it may run on."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().pbo_set_num_threads(c_int(n))
    return num_threads()


def normal(keys, n):
    keys = np.ascontiguousarray(keys, np.uint32).reshape(-1, 2)
    out = np.empty((len(keys), n), np.float32)
    lib().pbo_normal(_p(keys), c_int(len(keys)), c_int(n), _p(out))
    return out


def reference_draw(keys, draw, B, N):
    keys = np.ascontiguousarray(keys, np.uint32).reshape(-1, 2)
    out = np.empty((len(keys), B), np.int32)
    lib().pbo_reference_draw(_p(keys), c_int(len(keys)), c_int(draw), c_int(B), c_int(N), _p(out))
    return out


def sg_run(layers, act, sigma, algo, X, y, theta0, B, step, T, keys, L=1, alpha=0.01, beta=0.0, trace=True,
           momentum=(0.0, 0.0, 1.0, 0.0)):
    """algo: 'sgld' | 'psgld' | 'sghmc'; for psgld alpha is the preconditioning factor; momentum = SGHMC resampling
    coefficients (mu, mu_sqrt_step, sigma, sigma_sqrt_step)."""
    s = _spec(layers, act, sigma)
    X, y = _f32(X), _f32(y)
    th = _f32(theta0).copy()
    C, P = th.shape
    keys = np.ascontiguousarray(keys, np.uint32).reshape(C, 2)
    tr = np.empty((C, T, P), np.float32) if trace else None
    lib().pbo_sg_run(ctypes.byref(s), c_int({"sgld": 0, "psgld": 1, "sghmc": 2}[algo]), _p(X), _p(y), c_int(len(X)),
                     c_int(B), _p(keys), _p(th), c_float(step), c_int(T), c_int(C), c_int(L), c_float(alpha),
                     c_float(beta), _p(np.asarray(momentum, np.float32)), _p(tr))
    return th, tr


def hmc_run(layers, act, sigma, X, y, theta0, keys, S, step, inv_mass, L, trace=True):
    s = _spec(layers, act, sigma)
    X, y = _f32(X), _f32(y)
    th = _f32(theta0).copy()
    C, P = th.shape
    keys = np.ascontiguousarray(keys, np.uint32).reshape(C, 2)
    im = _f32(inv_mass)
    tr = np.empty((C, S, P), np.float32) if trace else None
    acc = np.empty((C, S), np.int32)
    lib().pbo_hmc_run(ctypes.byref(s), _p(X), _p(y), c_int(len(X)), _p(keys), _p(th), _p(im), c_float(step), c_int(L),
                      c_int(S), c_int(C), _p(tr), _p(acc))
    return th, tr, acc


def adam_run(layers, act, sigma, X, y, theta0, idx, lr):
    s = _spec(layers, act, sigma)
    X, y = _f32(X), _f32(y)
    th = _f32(theta0).copy()
    idx = np.ascontiguousarray(idx, np.int32)
    C, n_steps, B = idx.shape
    lib().pbo_adam_run(ctypes.byref(s), _p(X), _p(y), c_int(len(X)), _p(idx), c_int(n_steps), c_int(B), _p(th),
                       c_float(lr), c_int(C))
    return th


def predict_moments(layers, act, sigma, samples, Xt):
    s = _spec(layers, act, sigma)
    samples, Xt = _f32(samples), _f32(Xt)
    M, so = len(Xt), layers[-1]
    sm, sq = np.empty((M, so), np.float32), np.empty((M, so), np.float32)
    lib().pbo_predict_moments(ctypes.byref(s), _p(samples), c_int(len(samples)), _p(Xt), c_int(M), _p(sm), _p(sq))
    return sm, sq
