"""``pbnn.deep_ensembles.deep_ensembles_fn`` mirror (src/pbnn/deep_ensembles.py:16-107).

The reference trains the members one after the other (a sequential ``lax.scan``, :97-99); here all members advance
together, one CTA per member, with the same key flow: member k's epoch keys are ``split(train_key_k,
num_epochs + 1)`` (quirk Q3: one extra epoch), member 0 derives (train_key, init_key) from ``split(rng_key)``, members
k >= 1 from ``split(PRNGKey(k))`` (quirk Q4, :88-89); each epoch's order is ``jax.random.permutation`` reproduced on
device (``pbnn_permutation``).
"""
from __future__ import annotations

import numpy as np
import torch

from . import pytree
from .logprobs import require_specs
from .mcmc._common import make_predict_fn, make_ravel_fn
from .ops import default_ops


def member_keys(rng_key, num_networks: int):
    """(train_keys (K,2), init_keys (K,2)) following deep_ensembles.py:83-99."""
    ops = default_ops()
    base = [ops._keys(rng_key)]
    for k in range(1, num_networks):
        base.append(ops._keys(np.array([0, k], dtype=np.uint32)))  # jax.random.PRNGKey(k), k < 2^32
    sp = ops.split(torch.cat(base, 0), 2)  # (K, 2, 2)
    return sp[:, 0, :].contiguous(), sp[:, 1, :].contiguous()


def epoch_indices(train_keys, N: int, batch_size: int, num_epochs: int):
    """(K, (num_epochs+1) * steps, B) minibatch rows: permutation(epoch_key, N)[:steps*B] per epoch (:59-66)."""
    ops = default_ops()
    K = train_keys.shape[0]
    steps = N // batch_size
    ek = ops.split(train_keys, num_epochs + 1).reshape(-1, 2)  # (K*(E+1), 2)
    perms = ops.permutation(ek, N)  # (K*(E+1), N)
    idx = perms[:, : steps * batch_size].reshape(K, (num_epochs + 1) * steps, batch_size)
    return idx.contiguous()


def deep_ensembles_fn(X, y, loglikelihood_fn, logprior_fn, network, batch_size, num_epochs, step_size, num_networks,
                      rng_key, *, init_params=None):
    """Returns ``(params stacked over members, ravel_fn, predict_fn)`` like deep_ensembles.py:101-107.
    ``init_params`` (keyword extension): pytree with a leading ``num_networks`` axis replacing ``network.init``."""
    require_specs(loglikelihood_fn, logprior_fn)
    ops = default_ops()
    N, d = int(np.shape(X)[0]), int(np.shape(X)[1])
    layers = network.widths(d)
    spec = loglikelihood_fn.flat_spec(d, logprior_fn)
    train_keys, init_keys = member_keys(rng_key, int(num_networks))
    if init_params is None:
        trees = [network.init(init_keys[k], np.zeros((d,), np.float32))["params"] for k in range(int(num_networks))]
        flat = torch.stack([pytree.ravel(t, layers) for t in trees])
    else:
        flat = pytree.ravel(init_params, layers)
    idx = epoch_indices(train_keys, N, int(batch_size), int(num_epochs))
    res = ops.adam_ensemble(spec, X, y, flat, idx, float(step_size))
    return pytree.unravel(res["theta"], layers), make_ravel_fn(layers), make_predict_fn(layers)
