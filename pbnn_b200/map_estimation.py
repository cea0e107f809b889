"""``pbnn.map_estimation.train_fn`` mirror (src/pbnn/map_estimation.py:42-108): MAP estimate of one network with Adam
or SGD on ``-logposterior``.  The reference pre-batches the data ONCE with a shuffling torch ``DataLoader``
(``drop_last=True``) and replays the same batch order every epoch (:85-100); the same sampler calls are made here so
that, under the same torch seed, the batch order is the reference's.  The optimisation steps run in the fused CUDA
chain kernel (Adam: the ensemble kernel with one member; SGD: the temperature-0 Langevin step)."""
from __future__ import annotations

import numpy as np
import torch

from . import pytree
from .mcmc.kernels import grad_estimator
from .ops import default_ops


def build_logposterior_estimator_fn(logprior_fn, loglikelihood_fn, data_size: int) -> grad_estimator:
    """``pbnn.utils.misc.build_logposterior_estimator_fn`` (src/pbnn/utils/misc.py:34-55) as a spec object:
    ``logprior + data_size * mean_i loglik_i``."""
    return grad_estimator(logprior_fn, loglikelihood_fn, data_size)


def _loader_batches(n: int, batch_size: int) -> np.ndarray:
    """Batch order of ``NumpyLoader(dataset, batch_size, shuffle=True, drop_last=True)``: torch's RandomSampler seeds a
    fresh generator from the global RNG and draws one ``randperm``."""
    seed = int(torch.empty((), dtype=torch.int64).random_().item())
    gen = torch.Generator()
    gen.manual_seed(seed)
    perm = torch.randperm(n, generator=gen).numpy()
    steps = n // batch_size
    return perm[: steps * batch_size].reshape(steps, batch_size).astype(np.int32)


def train_fn(logposterior_fn, network, train_ds, batch_size, num_epochs, learning_rate, rng_key, optimizer="adam", *,
             init_params=None, batch_indices=None):
    """Returns the MAP parameters (pytree), like the reference.  Keyword extensions: ``init_params`` (skip
    ``network.init``) and ``batch_indices`` ((steps, batch_size) rows, replayed every epoch)."""
    if not isinstance(logposterior_fn, grad_estimator):
        raise TypeError("logposterior_fn must come from pbnn_b200.map_estimation.build_logposterior_estimator_fn")
    if optimizer not in ("adam", "sgd"):
        raise ValueError("optimizer must be 'adam' or 'sgd'")
    ops = default_ops()
    X, y = train_ds["x"], train_ds["y"]
    n, d = int(np.shape(X)[0]), int(np.shape(X)[1])
    layers = network.widths(d)
    if int(logposterior_fn.data_size) != n:
        # the Adam / SGD kernels scale the minibatch log-likelihood by len(X) / batch_size; the reference would use
        # data_size * mean(loglik) -- a different posterior.  Refuse instead of silently training on another target.
        raise ValueError(f"build_logposterior_estimator_fn was given data_size={logposterior_fn.data_size}, but the "
                         f"training set has {n} rows; the fused trainers support data_size == len(X) only")
    spec = logposterior_fn.spec_for(layers)
    if init_params is None:
        init_rng = ops.split(rng_key, 2)[:, 1, :]
        init_params = network.init(init_rng[0], np.zeros((d,), np.float32))["params"]
    flat = pytree.ravel(init_params, layers).reshape(1, -1)
    batches = _loader_batches(n, int(batch_size)) if batch_indices is None else np.asarray(batch_indices, np.int32)
    idx = np.tile(batches, (int(num_epochs), 1))[None]  # (1, epochs * steps, B)
    if optimizer == "adam":
        res = ops.adam_ensemble(spec, X, y, flat, idx, float(learning_rate))
    else:
        T = idx.shape[1]
        res = ops.sgld(spec, X, y, flat, np.zeros((1, 2), np.uint32), int(batch_size), float(learning_rate), T, idx=idx,
                       temperature=0.0, keep_trace=False)
    return pytree.unravel(res["theta"][0], layers)
