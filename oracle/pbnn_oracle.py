"""CPU restatement of pbnn's SG-MCMC / HMC / deep-ensemble hot path  --  TEST INFRASTRUCTURE ONLY.

This file is the parity *oracle*.  It is imported only by ``tests/``, by
``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs.  Nothing under ``pbnn_b200/`` (the product) may import it.

PARITY STATUS: **parity unpinned by the reference** -- the reference's own test-suite is
``assert True`` (``/root/reference/tests/test_dummpy.py:1-2``) and JAX / BlackJAX / Flax / Optax
cannot be installed in this environment, so the reference cannot be executed.  The RNG layer
is pinned against the Random123 threefry2x32 known-answer vectors and against outputs printed
in the public JAX documentation (``tests/test_oracle.py``); everything above the RNG is a
restatement of the formulas in the files cited per function.

Conventions
-----------
* All arithmetic is float32 unless ``dtype=np.float64`` is requested (used as the
  high-precision yard-stick for tolerances).
* Parameters are *flat* vectors in ``jax.flatten_util.ravel_pytree`` order of the Flax tree
  ``{"Dense_k": {"bias": (out,), "kernel": (in, out)}}``: dict keys sorted, so for every layer
  the bias comes first, then the kernel row-major ``[in][out]``.
* ``rng_key`` is a legacy ``uint32[2]`` threefry key, ``jax_threefry_partitionable=True``
  (the default since jax 0.5.0; the reference pins jax 0.7.1 in ``uv.lock``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Iterator, Sequence

import numpy as np

U32 = np.uint32
_M32 = 0xFFFFFFFF

# ----------------------------------------------------------------------------------------------
# threefry2x32 and the jax.random primitives used on the path
# (jax/_src/prng.py: threefry2x32, _threefry_split_foldlike, _threefry_random_bits_partitionable;
#  jax/_src/random.py: _uniform, _normal_real, _randint, _shuffle.  Call sites in pbnn:
#  src/pbnn/mcmc/langevin.py:69, src/pbnn/utils/data.py:70-76, src/pbnn/deep_ensembles.py:64,85,88-89)
# ----------------------------------------------------------------------------------------------

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return ((x << U32(r)) | (x >> U32(32 - r))).astype(U32)


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds (Salmon et al. 2011).  All arguments broadcastable uint32."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=U32)
        k1 = np.asarray(k1, dtype=U32)
        x0 = np.asarray(x0, dtype=U32).copy()
        x1 = np.asarray(x1, dtype=U32).copy()
        ks = (k0, k1, (k0 ^ k1 ^ U32(0x1BD11BDA)).astype(U32))
        x0 = (x0 + ks[0]).astype(U32)
        x1 = (x1 + ks[1]).astype(U32)
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = (x0 + x1).astype(U32)
                x1 = _rotl(x1, r)
                x1 = (x1 ^ x0).astype(U32)
            x0 = (x0 + ks[(i + 1) % 3]).astype(U32)
            x1 = (x1 + ks[(i + 2) % 3] + U32(i + 1)).astype(U32)
    return x0, x1


def PRNGKey(seed: int) -> np.ndarray:
    """jax.random.PRNGKey: ``[seed >> 32, seed & 0xffffffff]``."""
    seed = int(seed)
    return np.array([(seed >> 32) & _M32, seed & _M32], dtype=U32)


def split(key, n: int = 2) -> np.ndarray:
    """jax.random.split (partitionable): ``out[i] = threefry(key, (hi=0, lo=i))`` -> (n, 2)."""
    key = np.asarray(key, dtype=U32)
    i = np.arange(n, dtype=np.uint64)
    x0, x1 = threefry2x32(key[0], key[1], (i >> np.uint64(32)).astype(U32), (i & np.uint64(_M32)).astype(U32))
    return np.stack([x0, x1], axis=-1)


def fold_in(key, data: int) -> np.ndarray:
    """jax.random.fold_in: ``threefry(key, (0, data))``."""
    key = np.asarray(key, dtype=U32)
    x0, x1 = threefry2x32(key[0], key[1], U32(0), U32(int(data) & _M32))
    return np.array([x0, x1], dtype=U32)


def random_bits(key, n: int) -> np.ndarray:
    """32 random bits per element: ``x0 ^ x1`` of ``threefry(key, (i>>32, i&0xffffffff))``."""
    key = np.asarray(key, dtype=U32)
    i = np.arange(n, dtype=np.uint64)
    x0, x1 = threefry2x32(key[0], key[1], (i >> np.uint64(32)).astype(U32), (i & np.uint64(_M32)).astype(U32))
    return (x0 ^ x1).astype(U32)


def random_bits_legacy(key, n: int) -> np.ndarray:
    """jax < 0.5 default (``jax_threefry_partitionable=False``) -- kept only to detect a
    mis-configured oracle in the tests."""
    key = np.asarray(key, dtype=U32)
    m = max(1, n)
    cnt = np.arange(m, dtype=U32)
    if m % 2:
        cnt = np.concatenate([cnt, np.zeros(1, U32)])  # threefry_2x32 pads odd inputs with a zero
    half = len(cnt) // 2
    y0, y1 = threefry2x32(key[0], key[1], cnt[:half], cnt[half:])
    return np.concatenate([y0, y1])[:n].astype(U32)


def bits_to_uniform(bits: np.ndarray) -> np.ndarray:
    """float32 in [0,1): ``bitcast((bits >> 9) | 0x3f800000) - 1``."""
    f = ((bits >> U32(9)) | U32(0x3F800000)).astype(U32).view(np.float32)
    return (f - np.float32(1.0)).astype(np.float32)


def uniform(key, n: int) -> np.ndarray:
    u = bits_to_uniform(random_bits(key, n))
    return np.maximum(np.float32(0.0), u)


_ERFINV_LT5 = np.array(
    [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087,
     -0.00125372503, -0.00417768164, 0.246640727, 1.50140941], dtype=np.float32)
_ERFINV_GE5 = np.array(
    [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773,
     -0.0076224613, 0.00943887047, 1.00167406, 2.83297682], dtype=np.float32)


def erf_inv_f32(x: np.ndarray) -> np.ndarray:
    """XLA's float32 ErfInv (Giles' single-precision polynomial; xla/hlo/builder/lib/math.cc)."""
    x = np.asarray(x, dtype=np.float32)
    w = (-np.log1p((-(x * x)).astype(np.float32))).astype(np.float32)
    lt = w < np.float32(5.0)
    w2 = np.where(lt, w - np.float32(2.5), np.sqrt(np.maximum(w, 0)).astype(np.float32) - np.float32(3.0)).astype(np.float32)
    p = np.where(lt, _ERFINV_LT5[0], _ERFINV_GE5[0]).astype(np.float32)
    for i in range(1, 9):
        p = (np.where(lt, _ERFINV_LT5[i], _ERFINV_GE5[i]).astype(np.float32) + p * w2).astype(np.float32)
    res = (p * x).astype(np.float32)
    return np.where(np.abs(x) == 1, x * np.float32(np.inf), res).astype(np.float32)


_NORMAL_LO = np.nextafter(np.float32(-1.0), np.float32(0.0), dtype=np.float32)


def bits_to_normal(bits: np.ndarray) -> np.ndarray:
    """jax.random.normal float32: ``sqrt(2) * erf_inv(uniform(lo=nextafter(-1,0), hi=1))``."""
    u = bits_to_uniform(bits)
    span = np.float32(np.float32(1.0) - _NORMAL_LO)  # rounds to exactly 2.0f
    v = np.maximum(_NORMAL_LO, (u * span + _NORMAL_LO).astype(np.float32))
    return (np.float32(np.sqrt(2)) * erf_inv_f32(v)).astype(np.float32)


def normal(key, n: int) -> np.ndarray:
    return bits_to_normal(random_bits(key, n))


def randint(key, n: int, minval: int, maxval: int) -> np.ndarray:
    """jax.random.randint for int32 (used by ``choice(replace=True)``: utils/data.py:70-76)."""
    k1, k2 = split(key, 2)
    hi = random_bits(k1, n).astype(np.uint64)
    lo = random_bits(k2, n).astype(np.uint64)
    span = (maxval - minval) if maxval > minval else 1
    mult = (1 << 16) % span
    mult = ((mult * mult) & _M32) % span  # lax.mul in uint32: wraps to 0 when span > 2**16 (as JAX does)
    with np.errstate(over="ignore"):
        off = (((hi % span) * mult) & _M32) + (lo % span)  # uint32 wrap-around arithmetic
        off = (off & _M32) % span
    return (minval + off).astype(np.int32)


def choice_with_replacement(key, N: int, B: int) -> np.ndarray:
    return randint(key, B, 0, N)


def permutation(key, N: int) -> np.ndarray:
    """jax.random.permutation(key, N): repeated stable sort by fresh random keys."""
    x = np.arange(N, dtype=np.int32)
    rounds = int(np.ceil(3 * np.log(max(1, N)) / np.log(np.iinfo(np.uint32).max)))
    key = np.asarray(key, dtype=U32)
    for _ in range(rounds):
        key, sub = split(key, 2)
        sk = random_bits(sub, N)
        order = np.argsort(sk, kind="stable")
        x = x[order]
    return x


# ----------------------------------------------------------------------------------------------
# MLP, densities, gradient  (src/pbnn/models.py:16-35; benchmark/pipelines/logprobs.py:6-17;
# blackjax grad_estimator == src/pbnn/utils/misc.py:34-55)
# ----------------------------------------------------------------------------------------------

RELU, TANH = 0, 1


@dataclass(frozen=True)
class MlpSpec:
    layers: tuple  # (d, h1, ..., out)
    act: int = RELU
    noise_level: float = 0.1
    prior_scale: float = 1.0
    heteroscedastic: bool = False  # outputs (mu, log sigma^2), logprobs.py:37-42

    @property
    def n_layers(self) -> int:
        return len(self.layers) - 1

    def offsets(self):
        """[(bias_off, kernel_off, in, out)] in ravel_pytree order (bias before kernel)."""
        out, off = [], 0
        names = sorted(range(self.n_layers), key=lambda k: f"Dense_{k}")
        table = {}
        for k in names:
            i, o = self.layers[k], self.layers[k + 1]
            table[k] = (off, off + o, i, o)
            off += o + i * o
        for k in range(self.n_layers):
            out.append(table[k])
        return out

    @property
    def num_params(self) -> int:
        return sum(o + i * o for i, o in zip(self.layers[:-1], self.layers[1:]))


def unravel(spec: MlpSpec, theta: np.ndarray):
    """flat (..., P) -> list of (bias (..., out), kernel (..., in, out))."""
    res = []
    for (bo, wo, i, o) in spec.offsets():
        b = theta[..., bo:bo + o]
        w = theta[..., wo:wo + i * o].reshape(theta.shape[:-1] + (i, o))
        res.append((b, w))
    return res


def to_pytree(spec: MlpSpec, theta: np.ndarray):
    return {f"Dense_{k}": {"bias": b, "kernel": w} for k, (b, w) in enumerate(unravel(spec, theta))}


def _act(spec, z):
    return np.maximum(z, 0) if spec.act == RELU else np.tanh(z)


def _act_grad_from_out(spec, h):
    return (h > 0).astype(h.dtype) if spec.act == RELU else (1 - h * h)


def mlp_forward(spec: MlpSpec, theta: np.ndarray, X: np.ndarray, return_hidden=False):
    """theta (C, P), X (C, B, d) or (B, d) -> f (C, B, out).  models.py:16-35 generalised to a
    ``layers`` list."""
    dt = theta.dtype
    h = np.broadcast_to(X, (theta.shape[0],) + X.shape[-2:]).astype(dt) if X.ndim == 2 else X.astype(dt)
    hs = [h]
    params = unravel(spec, theta)
    for k, (b, w) in enumerate(params):
        z = np.matmul(h, w) + b[:, None, :]
        h = z if k == spec.n_layers - 1 else _act(spec, z)
        hs.append(h)
    return (h, hs) if return_hidden else h


def loglik_and_grad(spec: MlpSpec, theta: np.ndarray, X: np.ndarray, y: np.ndarray, scale: float):
    """Returns (sum_i loglik_i, scale * d/dtheta sum_i loglik_i) for the homoscedastic Gaussian
    likelihood ``-sum 0.5 (y - f)^2 / sigma^2`` (logprobs.py:12-17).  theta (C,P); X (C,B,d)|(B,d);
    y (C,B,s)|(B,s)."""
    dt = theta.dtype
    C = theta.shape[0]
    f, hs = mlp_forward(spec, theta, X, return_hidden=True)
    yb = np.broadcast_to(y, (C,) + y.shape[-2:]).astype(dt) if y.ndim == 2 else y.astype(dt)
    if spec.heteroscedastic:
        mu, ls = f[..., 0], f[..., 1]
        res = yb[..., 0] - mu
        e = np.exp(-ls)
        ll = (-dt.type(0.5) * ls - dt.type(0.5) * res * res * e).sum(axis=1)
        dz = np.stack([res * e, dt.type(-0.5) + dt.type(0.5) * res * res * e], axis=-1) * dt.type(scale)
    else:
        inv_s2 = dt.type(1.0) / (dt.type(spec.noise_level) * dt.type(spec.noise_level))
        res = yb - f
        ll = -(dt.type(0.5) * res * res * inv_s2).sum(axis=(1, 2))
        dz = res * inv_s2 * dt.type(scale)  # d(scale*ll)/df
    g = np.zeros_like(theta)
    params = unravel(spec, theta)
    offs = spec.offsets()
    for k in range(spec.n_layers - 1, -1, -1):
        bo, wo, i, o = offs[k]
        hin = hs[k]
        g[:, bo:bo + o] = dz.sum(axis=1)
        g[:, wo:wo + i * o] = np.matmul(np.swapaxes(hin, 1, 2), dz).reshape(C, i * o)
        if k > 0:
            dh = np.matmul(dz, np.swapaxes(params[k][1], 1, 2))
            dz = dh * _act_grad_from_out(spec, hin)
    return ll, g


def logprior(spec: MlpSpec, theta: np.ndarray):
    """sum_j norm.logpdf(theta_j) (logprobs.py:6-9); prior_scale=1 in the reference."""
    dt = theta.dtype
    s = dt.type(spec.prior_scale)
    t = theta / s
    return (-(t * t) / dt.type(2) - dt.type(math.log(math.sqrt(2 * math.pi))) - dt.type(math.log(spec.prior_scale))).sum(axis=-1)


def grad_estimator(spec: MlpSpec, theta, Xb, yb, N: int):
    """grad of ``logprior + N * mean_i loglik_i``  (blackjax grad_estimator; misc.py:48-53)."""
    dt = theta.dtype
    B = Xb.shape[-2]
    _, g = loglik_and_grad(spec, theta, Xb, yb, dt.type(N) / dt.type(B))
    s2 = dt.type(spec.prior_scale) ** 2
    return g - theta / s2


def logposterior_full(spec: MlpSpec, theta, X, y):
    """``logprior + sum_i loglik_i`` and its gradient (benchmark/pipelines/hmc.py:59-62)."""
    ll, g = loglik_and_grad(spec, theta, X, y, 1.0)
    s2 = theta.dtype.type(spec.prior_scale) ** 2
    return logprior(spec, theta) + ll, g - theta / s2


# ----------------------------------------------------------------------------------------------
# minibatch generator (src/pbnn/utils/data.py:47-78) with the trace-time semantics of SURVEY Q1
# ----------------------------------------------------------------------------------------------

def batch_index_stream(rng_key, B: int, N: int) -> Iterator[np.ndarray]:
    key = np.asarray(rng_key, dtype=U32)
    while True:
        key = split(key, 2)[1]
        yield choice_with_replacement(key, N, B)


def reference_draw(rng_key, B: int, N: int, draw: int) -> np.ndarray:
    """Indices of the ``draw``-th ``next(batches)`` (1-based)."""
    it = batch_index_stream(rng_key, B, N)
    idx = None
    for _ in range(draw):
        idx = next(it)
    return idx


# ----------------------------------------------------------------------------------------------
# samplers.  Every function advances C independent chains: chain c == one reference call with
# (rng_key[c], theta0[c]).  ``which_draw`` is the minibatch the traced loop re-uses (Q1).
# ----------------------------------------------------------------------------------------------

def _as_keys(keys):
    keys = np.asarray(keys, dtype=U32)
    return keys.reshape(-1, 2)


def _fixed_batches(spec, X, y, keys, B, draw):
    N = X.shape[0]
    idx = np.stack([reference_draw(k, B, N, draw) for k in keys])  # (C, B)
    return X[idx], y[idx], idx


def _step_keys(keys, t):
    """keys[c] -> split(keys[c], T)[t] for every chain (langevin.py:69)."""
    x0, x1 = threefry2x32(keys[:, 0], keys[:, 1], U32(0), U32(t))
    return np.stack([x0, x1], axis=-1)


def _normals(step_keys, P, dt):
    return np.stack([normal(k, P) for k in step_keys]).astype(dt)


def _minibatch(X, y, keys, B, mode, draw, t, gens):
    N = X.shape[0]
    if mode == "reference_fixed":
        raise AssertionError
    idx = np.stack([next(g) for g in gens])
    return X[idx], y[idx]


def sgld(spec, X, y, theta0, B, step_size, T, keys, *, dtype=np.float32, which_draw=2,
         minibatch_mode="reference_fixed", t0=0, idx=None):
    """src/pbnn/mcmc/langevin.py:21-78 + blackjax.sgld: ``theta += eps*g + sqrt(2 eps)*xi``.
    Returns positions (C, T, P)."""
    keys = _as_keys(keys)
    dt = np.dtype(dtype)
    theta = np.array(theta0, dtype=dt).reshape(len(keys), -1)
    C, P = theta.shape
    N = X.shape[0]
    Xd, yd = X.astype(dt), y.astype(dt)
    gens = None
    if idx is not None:
        pass
    elif minibatch_mode == "reference_fixed":
        Xb, yb, _ = _fixed_batches(spec, Xd, yd, keys, B, which_draw)
    else:
        gens = [batch_index_stream(k, B, N) for k in keys]
        for g in gens:
            next(g)  # draw #1 is consumed by kern.init (langevin.py:60)
    eps = dt.type(np.float32(step_size))
    noise_scale = dt.type(np.sqrt(np.float32(2.0) * np.float32(step_size)))
    out = np.empty((C, T, P), dtype=dt)
    for t in range(T):
        if idx is not None:
            ii = idx[:, t if idx.shape[1] > 1 else 0]
            Xb, yb = Xd[ii], yd[ii]
        elif gens is not None:
            ii = np.stack([next(g) for g in gens])
            Xb, yb = Xd[ii], yd[ii]
        g = grad_estimator(spec, theta, Xb, yb, N)
        xi = _normals(_step_keys(keys, t0 + t), P, dt)
        theta = theta + eps * g + noise_scale * xi
        out[:, t] = theta
    return out


def psgld_init(spec, X, y, theta0, B, keys, *, dtype=np.float32, init_draw=1):
    """src/pbnn/mcmc/sgmcmc/psgld.py:25-29: g0 = grad(theta0, draw #1), V0 = g0^2."""
    keys = _as_keys(keys)
    dt = np.dtype(dtype)
    theta = np.array(theta0, dtype=dt).reshape(len(keys), -1)
    Xb, yb, _ = _fixed_batches(spec, X.astype(dt), y.astype(dt), keys, B, init_draw)
    g = grad_estimator(spec, theta, Xb, yb, X.shape[0])
    return g, g * g


def psgld(spec, X, y, theta0, B, step_size, T, alpha, keys, *, dtype=np.float32, which_draw=2,
          init_draw=1, t0=0, state=None, return_state=False):
    """src/pbnn/mcmc/langevin.py:81-140 + sgmcmc/diffusions.py:33-60 (stale-gradient RMSprop
    preconditioned Euler-Langevin, no Gamma term)."""
    keys = _as_keys(keys)
    dt = np.dtype(dtype)
    theta = np.array(theta0, dtype=dt).reshape(len(keys), -1)
    C, P = theta.shape
    N = X.shape[0]
    Xd, yd = X.astype(dt), y.astype(dt)
    if state is None:
        g, V = psgld_init(spec, X, y, theta, B, keys, dtype=dtype, init_draw=init_draw)
    else:
        g, V = (np.array(s, dtype=dt) for s in state)
    Xb, yb, _ = _fixed_batches(spec, Xd, yd, keys, B, which_draw)
    eps = dt.type(np.float32(step_size))
    a = dt.type(np.float32(alpha))
    out = np.empty((C, T, P), dtype=dt)
    for t in range(T):
        G = dt.type(1.0) / (np.sqrt(V) + dt.type(1e-6))
        xi = _normals(_step_keys(keys, t0 + t), P, dt)
        theta = theta + eps * G * g + np.sqrt(dt.type(2) * G * eps) * xi
        g = grad_estimator(spec, theta, Xb, yb, N)
        V = a * V + (dt.type(1.0) - a) * np.square(g)
        out[:, t] = theta
    return (out, (g, V)) if return_state else out


MOMENTUM_UNIT = (0.0, 0.0, 1.0, 0.0)             # p ~ N(0, 1): generate_gaussian_noise(rng_key, position)
MOMENTUM_SIGMA_SQRT_STEP = (0.0, 0.0, 0.0, 1.0)  # p ~ N(0, step): sigma = sqrt(step_size) (Chen et al. 2014)
MOMENTUM_MU_SQRT_STEP = (0.0, 1.0, 1.0, 0.0)     # p ~ N(sqrt(step), 1): sqrt(step_size) in the positional `mu` slot


def _momentum(momentum, step_size, dt):
    mu0, mu1, s0, s1 = (np.float32(v) for v in momentum)
    se = np.sqrt(np.float32(step_size))
    return dt.type(mu0 + mu1 * se), dt.type(s0 + s1 * se)


def sghmc(spec, X, y, theta0, B, step_size, T, L, keys, *, alpha=0.01, beta=0.0, dtype=np.float32,
          which_draw=2, t0=0, momentum=MOMENTUM_UNIT, _drop_drift=False):
    """src/pbnn/mcmc/hamiltonian.py:80-140 + blackjax.sghmc (1.2.5): fresh momentum from
    ``keys[t]``, ``split(keys[t], L)`` for the L inner noise draws; gradient at the position
    *before* the position update.

    ``momentum`` = (mu, mu_sqrt_step, sigma, sigma_sqrt_step): the resampled momentum is
    ``(mu + mu_sqrt_step sqrt(eps)) + (sigma + sigma_sqrt_step sqrt(eps)) z``.  The upstream body
    (blackjax/sgmcmc/sghmc.py, ``generate_gaussian_noise(rng_key, position, ...)``) is not vendored with the reference
    and cannot be read here, so -- like ``which_draw`` -- the variant is a parameter of the oracle, not a constant."""
    keys = _as_keys(keys)
    dt = np.dtype(dtype)
    theta = np.array(theta0, dtype=dt).reshape(len(keys), -1)
    C, P = theta.shape
    N = X.shape[0]
    Xb, yb, _ = _fixed_batches(spec, X.astype(dt), y.astype(dt), keys, B, which_draw)
    eps = dt.type(np.float32(step_size))
    fric = dt.type(1.0) - dt.type(np.float32(alpha))
    ns = dt.type(np.sqrt(np.float32(2.0) * np.float32(step_size) * (np.float32(alpha) - np.float32(beta))))
    mom_mu, mom_sig = _momentum(momentum, step_size, dt)
    out = np.empty((C, T, P), dtype=dt)
    for t in range(T):
        kt = _step_keys(keys, t0 + t)
        p = mom_mu + mom_sig * _normals(kt, P, dt)
        for l in range(L):
            sub = _step_keys(kt, l)
            g = grad_estimator(spec, theta, Xb, yb, N)
            xi = _normals(sub, P, dt)
            theta = theta + p
            p = fric * p + (dt.type(0) if _drop_drift else eps * g) + ns * xi  # _drop_drift: test-sensitivity probe
        out[:, t] = theta
    return out


def hmc(spec, X, y, theta0, num_samples, step_size, inv_mass, L, keys, *, dtype=np.float32, t0=0,
        return_info=False):
    """src/pbnn/mcmc/hamiltonian.py:18-77 + blackjax.hmc (1.2.5) with the full-batch target of
    benchmark/pipelines/hmc.py:59-62.  Velocity Verlet, one value_and_grad per leapfrog,
    Metropolis accept with one uniform.  Returns positions (C, S, P) (rejected => repeated row)."""
    keys = _as_keys(keys)
    dt = np.dtype(dtype)
    theta = np.array(theta0, dtype=dt).reshape(len(keys), -1)
    C, P = theta.shape
    Xd, yd = X.astype(dt), y.astype(dt)
    Minv = np.asarray(inv_mass, dtype=dt).reshape(P)
    msqrt = np.sqrt(dt.type(1.0) / Minv)
    eps = dt.type(np.float32(step_size))
    heps = dt.type(np.float32(step_size) * np.float32(0.5))
    lp, gl = logposterior_full(spec, theta, Xd, yd)
    out = np.empty((C, num_samples, P), dtype=dt)
    info = {"delta": np.empty((C, num_samples), dt), "accept": np.empty((C, num_samples), bool),
            "u": np.empty((C, num_samples), np.float32)}
    for s in range(num_samples):
        ks = _step_keys(keys, t0 + s)
        km = np.stack([split(k, 2)[0] for k in ks])
        ka = np.stack([split(k, 2)[1] for k in ks])
        p = msqrt * _normals(km, P, dt)
        th, pp, lp1, gl1 = theta.copy(), p.copy(), lp.copy(), gl.copy()
        for _ in range(L):
            pp = pp + heps * gl1
            th = th + eps * (Minv * pp)
            lp1, gl1 = logposterior_full(spec, th, Xd, yd)
            pp = pp + heps * gl1
        e0 = -lp + dt.type(0.5) * ((Minv * p) * p).sum(axis=-1)
        e1 = -lp1 + dt.type(0.5) * ((Minv * pp) * pp).sum(axis=-1)
        delta = e0 - e1
        delta = np.where(np.isnan(delta), -np.inf, delta).astype(dt)
        with np.errstate(over="ignore"):
            pacc = np.minimum(dt.type(1.0), np.exp(delta))
        u = np.array([uniform(k, 1)[0] for k in ka], dtype=np.float32)
        acc = u.astype(dt) < pacc
        theta = np.where(acc[:, None], th, theta)
        gl = np.where(acc[:, None], gl1, gl)
        lp = np.where(acc, lp1, lp)
        out[:, s] = theta
        info["delta"][:, s], info["accept"][:, s], info["u"][:, s] = delta, acc, u
    return (out, info) if return_info else out


# ----------------------------------------------------------------------------------------------
# control variates / SVRG / cyclical variants
# blackjax.sgmcmc.gradients.control_variates (1.2.5): g_hat(theta) = g_full(c) + g_mb(theta) - g_mb(c), with
# g_full(c) = grad_estimator(c, full data) (N * mean over N rows = the full-data log-posterior gradient).
# ----------------------------------------------------------------------------------------------

def control_variate_terms(spec, X, y, centering, Xb, yb, dtype=np.float32):
    dt = np.dtype(dtype)
    c = np.array(centering, dtype=dt)
    _, g_full = logposterior_full(spec, c, X.astype(dt), y.astype(dt))
    g_est = grad_estimator(spec, c, Xb, yb, X.shape[0])
    return g_full, g_est


def sgld_cv(spec, X, y, theta0, B, step_size, T, centering, keys, *, batch_keys=None, dtype=np.float32, which_draw=2,
            temperature=1.0):
    """src/pbnn/mcmc/langevin.py:207-270.  ``batch_keys``: keys of the minibatch generator when they differ from the
    step keys (the SVRG outer loop, :273-354)."""
    keys = _as_keys(keys)
    dt = np.dtype(dtype)
    theta = np.array(theta0, dtype=dt).reshape(len(keys), -1)
    C, P = theta.shape
    N = X.shape[0]
    Xb, yb, _ = _fixed_batches(spec, X.astype(dt), y.astype(dt), _as_keys(batch_keys if batch_keys is not None else keys),
                               B, which_draw)
    g_c, g_ce = control_variate_terms(spec, X, y, np.array(centering, dtype=dt).reshape(C, P), Xb, yb, dtype)
    eps = dt.type(np.float32(step_size))
    ns = dt.type(np.sqrt(np.float32(2.0) * np.float32(temperature) * np.float32(step_size)))
    out = np.empty((C, T, P), dtype=dt)
    for t in range(T):
        g = (g_c + grad_estimator(spec, theta, Xb, yb, N)) - g_ce
        theta = theta + eps * g + ns * _normals(_step_keys(keys, t), P, dt)
        out[:, t] = theta
    return out


def sghmc_cv(spec, X, y, theta0, B, step_size, T, centering, L, keys, *, batch_keys=None, alpha=0.01, beta=0.0,
             dtype=np.float32, which_draw=2, momentum=MOMENTUM_UNIT):
    """src/pbnn/mcmc/hamiltonian.py:207-273."""
    keys = _as_keys(keys)
    dt = np.dtype(dtype)
    theta = np.array(theta0, dtype=dt).reshape(len(keys), -1)
    C, P = theta.shape
    N = X.shape[0]
    Xb, yb, _ = _fixed_batches(spec, X.astype(dt), y.astype(dt), _as_keys(batch_keys if batch_keys is not None else keys),
                               B, which_draw)
    g_c, g_ce = control_variate_terms(spec, X, y, np.array(centering, dtype=dt).reshape(C, P), Xb, yb, dtype)
    eps = dt.type(np.float32(step_size))
    fric = dt.type(1.0) - dt.type(np.float32(alpha))
    ns = dt.type(np.sqrt(np.float32(2.0) * np.float32(step_size) * (np.float32(alpha) - np.float32(beta))))
    mom_mu, mom_sig = _momentum(momentum, step_size, dt)
    out = np.empty((C, T, P), dtype=dt)
    for t in range(T):
        kt = _step_keys(keys, t)
        p = mom_mu + mom_sig * _normals(kt, P, dt)
        for l in range(L):
            g = (g_c + grad_estimator(spec, theta, Xb, yb, N)) - g_ce
            xi = _normals(_step_keys(kt, l), P, dt)
            theta = theta + p
            p = fric * p + eps * g + ns * xi
        out[:, t] = theta
    return out


def svrg(algo, spec, X, y, theta0, B, step_size, centering, num_cv, num_svrg, keys, *, L=None, dtype=np.float32):
    """sgld_svrg (langevin.py:273-354) / sghmc_svrg (hamiltonian.py:276-359): outer keys = split(rng_key, num_svrg);
    every outer iteration re-centres on the last position; the minibatch generator stays keyed by rng_key."""
    keys = _as_keys(keys)
    theta = np.array(theta0, dtype=dtype).reshape(len(keys), -1)
    c = np.array(centering, dtype=dtype).reshape(theta.shape)
    outs = []
    for i in range(num_svrg):
        ki = _step_keys(keys, i)
        if algo == "sgld":
            tr = sgld_cv(spec, X, y, theta, B, step_size, num_cv, c, ki, batch_keys=keys, dtype=dtype)
        else:
            tr = sghmc_cv(spec, X, y, theta, B, step_size, num_cv, c, L, ki, batch_keys=keys, dtype=dtype)
        outs.append(tr)
        theta = tr[:, -1]
        c = theta
    return np.concatenate(outs, axis=1)


# ----------------------------------------------------------------------------------------------
# deep ensembles (src/pbnn/deep_ensembles.py:16-107; optax.adam 0.2.5 defaults)
# ----------------------------------------------------------------------------------------------

def adam_step(theta, g, m, v, count, lr, b1=0.9, b2=0.999, eps=1e-8):
    dt = theta.dtype
    b1, b2, eps, lr = dt.type(b1), dt.type(b2), dt.type(eps), dt.type(np.float32(lr))
    m = (dt.type(1) - b1) * g + b1 * m
    v = (dt.type(1) - b2) * (g * g) + b2 * v
    count = count + 1
    mh = m / (dt.type(1) - b1 ** dt.type(count))
    vh = v / (dt.type(1) - b2 ** dt.type(count))
    theta = theta + (-lr) * (mh / (np.sqrt(vh) + eps))
    return theta, m, v, count


def ensemble_epoch_indices(epoch_key, N, B):
    """train_epoch (deep_ensembles.py:59-66): permutation(rng, N)[:steps*B].reshape(steps, B)."""
    steps = N // B
    return permutation(epoch_key, N)[: steps * B].reshape(steps, B)


def ensemble_member_keys(rng_key, num_networks: int):
    """Key flow of deep_ensembles.py:83-99 (quirk Q4): returns [(train_key, init_key)] per member."""
    rk, ik = split(rng_key, 2)
    res = []
    for m in range(num_networks):
        res.append((rk, ik))
        rk, ik = split(PRNGKey(m + 1), 2)
    return res


def deep_ensembles(spec, X, y, init_params, B, num_epochs, lr, member_train_keys, *, dtype=np.float32,
                   idx=None):
    """Adam on ``-logposterior`` for ``num_epochs + 1`` epochs (quirk Q3) per member.
    ``init_params`` (K, P) are inputs (Flax's path-hashed init keys are not restated).
    ``idx`` optionally overrides the index stream: (K, n_steps, B)."""
    dt = np.dtype(dtype)
    theta = np.array(init_params, dtype=dt)
    K, P = theta.shape
    N = X.shape[0]
    Xd, yd = X.astype(dt), y.astype(dt)
    if idx is None:
        rows = []
        for k in range(K):
            ek = split(member_train_keys[k], num_epochs + 1)
            rows.append(np.concatenate([ensemble_epoch_indices(e, N, B) for e in ek]))
        idx = np.stack(rows)
    m = np.zeros_like(theta)
    v = np.zeros_like(theta)
    count = 0
    for s in range(idx.shape[1]):
        ii = idx[:, s]
        g = -grad_estimator(spec, theta, Xd[ii], yd[ii], N)
        theta, m, v, count = adam_step(theta, g, m, v, count, lr)
    return theta


# ----------------------------------------------------------------------------------------------
# thinning (src/pbnn/utils/misc.py:58-125)
# ----------------------------------------------------------------------------------------------

def thinning(positions, size: int, dtype=np.float32):
    """Greedy energy-distance thinning; returns the indices in the reference's order (picks 1..size-1, then the
    initial pick appended last, misc.py:121-123)."""
    X = np.asarray(positions, dtype=dtype)
    xx = (X * X).sum(1)
    nrm = np.sqrt(xx)
    gram = X @ X.T
    K = nrm[:, None] + nrm[None, :] - np.sqrt(np.clip(xx[:, None] + xx[None, :] - 2 * gram, 0, None))
    k0_mean = K.mean(axis=1)
    obj = 2.0 * nrm - 2.0 * k0_mean
    init = int(np.argmin(obj))
    idx, picks = init, []
    for _ in range(1, size):
        ki = nrm[idx] + nrm - np.sqrt(np.clip(xx[idx] + xx - 2 * (X @ X[idx]), 0, None))
        obj = obj + 2.0 * ki - 2.0 * k0_mean
        idx = int(np.argmin(obj))
        picks.append(idx)
    return np.array(picks + [init], dtype=np.int32)


# ----------------------------------------------------------------------------------------------
# predictive (predict_fn: langevin.py:75-76; mean: benchmark/pipelines/metrics_prediction_intervals.py:188)
# ----------------------------------------------------------------------------------------------

def predict(spec, samples, X_test, dtype=np.float32):
    """samples (S, P), X_test (M, d) -> (S, M, out)."""
    return mlp_forward(spec, np.asarray(samples, dtype=dtype), np.asarray(X_test, dtype=dtype))


def predictive_moments(spec, samples, X_test, dtype=np.float32):
    F = predict(spec, samples, X_test, dtype=dtype)
    return F.mean(axis=0), F.var(axis=0)


# ----------------------------------------------------------------------------------------------
# synthetic UCI-shaped data -- shared by tests and bench so both sides see the
# same inputs.
# ----------------------------------------------------------------------------------------------

def synthetic_regression(N: int, d: int, seed: int = 0, M: int = 0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N + M, d)).astype(np.float32)
    X = ((X - X[:N].mean(0)) / X[:N].std(0)).astype(np.float32)
    w = rng.standard_normal(d).astype(np.float32)
    y = (np.tanh(X @ w / np.sqrt(np.float32(d))) + 0.1 * rng.standard_normal(N + M)).astype(np.float32)[:, None]
    if M:
        return X[:N], y[:N], X[N:], y[N:]
    return X, y


def init_positions(spec: MlpSpec, keys) -> np.ndarray:
    """theta_c ~ N(0, 0.01^2) (Flax init law of models.py:21-22) from per-chain keys."""
    keys = _as_keys(keys)
    return np.stack([np.float32(0.01) * normal(k, spec.num_params) for k in keys]).astype(np.float32)
