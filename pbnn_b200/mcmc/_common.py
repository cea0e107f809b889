"""Shared plumbing of the sampler wrappers: pytree <-> flat, chain batching, ravel_fn / predict_fn closures."""
from __future__ import annotations

import numpy as np
import torch

from .. import pytree
from ..ops import FlatSpec, default_ops


def chain_layout(init_positions, rng_key):
    """Returns (flat (C,P) tensor, widths, batched?) -- a single chain is (2,)-shaped key + unbatched leaves,
    exactly the reference call; C chains are a (C,2) key array + leaves with a leading C."""
    ops = default_ops()
    keys = ops._keys(rng_key)
    batched = np.ndim(rng_key) == 2 if not isinstance(rng_key, torch.Tensor) else rng_key.ndim == 2
    C = keys.shape[0]
    layers = pytree.widths_of(init_positions)
    flat = pytree.ravel(init_positions, layers)
    if flat.ndim == 1:
        flat = flat.reshape(1, -1).expand(C, -1)
    if flat.shape[0] != C:
        raise ValueError(f"init_positions carry {flat.shape[0]} chains but rng_key has {C}")
    return flat, layers, keys, batched


def setup(loglikelihood_fn, logprior_fn, init_positions, rng_key):
    """Common front end of the SG wrappers: spec objects checked, chains laid out, network widths verified."""
    from ..logprobs import require_specs
    require_specs(loglikelihood_fn, logprior_fn)
    flat, layers, keys, batched = chain_layout(init_positions, rng_key)
    spec = loglikelihood_fn.flat_spec(layers[0], logprior_fn)
    if tuple(spec.layers) != tuple(layers):
        raise ValueError(f"init_positions widths {layers} do not match the network {spec.layers}")
    return default_ops(), spec, flat, layers, keys, batched


def check_flags(res, who: str):
    """The kernels report per-chain conditions in ``flags`` (include/pbnn_b200.h): bit 0 non-finite state (the
    reference's pipelines detect NaNs after the fact, benchmark/pipelines/sgmcmc.py:207-209), bit 2 clamped minibatch
    index.  A warning, not an error: a diverged chain is a legitimate sampler outcome.  (One small D2H read.)"""
    import warnings
    flags = res.get("flags")
    if flags is None:
        return
    f = flags.detach().cpu().numpy()
    if (f & 1).any():
        warnings.warn(f"{who}: {int((f & 1).sum())} of {f.size} chains became non-finite (flags bit 0)", RuntimeWarning,
                      stacklevel=3)
    if (f & 4).any():
        warnings.warn(f"{who}: minibatch indices outside [0, N) were clamped (flags bit 2)", RuntimeWarning, stacklevel=3)


def positions_tree(trace: torch.Tensor, layers, batched: bool):
    """(C, T, P) trace -> pytree with leading (T, ...) (single chain) or (C, T, ...)."""
    t = trace if batched else trace[0]
    return pytree.unravel(t, layers)


def make_ravel_fn(layers):
    def ravel_fn(tree):
        """``jax.vmap(lambda t: ravel_pytree(t)[0])(pytree)`` (langevin.py:72-73): leading axes kept, -> (..., P)."""
        return pytree.ravel(tree, layers)

    return ravel_fn


def make_predict_fn(layers):
    def predict_fn(network, params, X_test):
        """``vmap(lambda p: network.apply({"params": p}, X_test))(params)`` -> (S, M, out) (langevin.py:75-76);
        runs the batched predictive kernel over samples x test points."""
        ops = default_ops()
        if tuple(network.widths(layers[0])) != tuple(layers):
            raise ValueError("network does not match the sampled parameters")
        flat = pytree.ravel(params, layers)
        lead = tuple(flat.shape[:-1])
        out = ops.predict(FlatSpec(layers, network.act_code), flat.reshape(-1, flat.shape[-1]), X_test)["preds"]
        return out.reshape(lead + tuple(out.shape[1:]))

    return predict_fn


def predictive_moments(network, params, X_test):
    """mean / variance (ddof=0) over the leading sample axes without materialising (S, M, out)
    (mean: benchmark/pipelines/metrics_prediction_intervals.py:188; variance: our extension)."""
    ops = default_ops()
    layers = pytree.widths_of(params)
    flat = pytree.ravel(params, layers)
    res = ops.predict(FlatSpec(layers, network.act_code), flat.reshape(-1, flat.shape[-1]), X_test, want_preds=False,
                      want_moments=True)
    return res["mean"], res["var"]
