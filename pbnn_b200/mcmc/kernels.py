"""``pbnn.mcmc.kernels.psgld`` mirror (src/pbnn/mcmc/kernels.py:18-94; sgmcmc/psgld.py:18-48;
sgmcmc/diffusions.py:22-62): the BlackJAX-style ``SamplingAlgorithm(init, step)`` pair of the preconditioned
SGLD kernel, executed by the fused CUDA chain-step kernel one transition at a time."""
from __future__ import annotations

from typing import NamedTuple

import numpy as np
import torch

from .. import pytree
from ..logprobs import GaussianMLPLikelihood, StdNormalPrior, require_specs
from ..ops import check, default_ops

__all__ = ["psgld", "grad_estimator", "pSGLDState", "SamplingAlgorithm"]


class SamplingAlgorithm(NamedTuple):
    init: callable
    step: callable


class pSGLDState(NamedTuple):  # sgmcmc/psgld.py:18-22
    step: int
    position: dict
    logprob_grad: dict
    square_avg: dict


class grad_estimator:
    """``blackjax.sgmcmc.gradients.grad_estimator(logprior_fn, loglikelihood_fn, data_size)`` as a spec object:
    gradient of ``logprior + data_size * mean_i loglik_i`` (== src/pbnn/utils/misc.py:34-55).  Callable:
    ``grad_fn(position, (X_batch, y_batch))`` -> gradient pytree, evaluated by the CUDA engine."""

    def __init__(self, logprior_fn, loglikelihood_fn, data_size: int):
        require_specs(loglikelihood_fn, logprior_fn)
        self.logprior_fn, self.loglikelihood_fn, self.data_size = logprior_fn, loglikelihood_fn, int(data_size)

    def spec_for(self, layers):
        return self.loglikelihood_fn.flat_spec(layers[0], self.logprior_fn)

    def __call__(self, position, batch):
        state = psgld.init(position, batch, self)
        return state.logprob_grad


def _flat_single(tree):
    layers = pytree.widths_of(tree)
    return pytree.ravel(tree, layers).reshape(1, -1), layers


class psgld:
    """``psgld(grad_estimator_fn, alpha=0.95)`` -> ``SamplingAlgorithm(init_fn(position, batch),
    step_fn(rng_key, state, batch, step_size))`` (kernels.py:80-94)."""

    @staticmethod
    def init(position, batch, grad_estimator_fn):
        ops = default_ops()
        flat, layers = _flat_single(position)
        spec = grad_estimator_fn.spec_for(layers)
        th = ops._theta(flat, spec, 1)
        Xb, yb = ops._data(batch[0], batch[1], spec)
        g, V = ops._zeros(th.shape, ops.f32), ops._zeros(th.shape, ops.f32)
        rc = ops.lib.pbnn_psgld_init(spec.c_struct(), ops._ptr(Xb), ops._ptr(yb), Xb.shape[0],
                                     grad_estimator_fn.data_size, ops._ptr(th), ops._ptr(g), ops._ptr(V), 1,
                                     ops._stream())
        check(ops.lib, rc, "psgld.init")
        return pSGLDState(0, position, pytree.unravel(g[0], layers), pytree.unravel(V[0], layers))

    @staticmethod
    def kernel(grad_estimator_fn):
        def one_step(rng_key, state, data_batch, step_size, alpha):
            ops = default_ops()
            flat, layers = _flat_single(state.position)
            spec = grad_estimator_fn.spec_for(layers)
            th = ops._theta(flat, spec, 1)
            g = ops._theta(pytree.ravel(state.logprob_grad, layers), spec, 1)
            V = ops._theta(pytree.ravel(state.square_avg, layers), spec, 1)
            Xb, yb = ops._data(data_batch[0], data_batch[1], spec)
            keys = ops._keys(rng_key)
            rc = ops.lib.pbnn_psgld_step(spec.c_struct(), ops._ptr(Xb), ops._ptr(yb), Xb.shape[0],
                                         grad_estimator_fn.data_size, ops._ptr(keys), ops._ptr(th), ops._ptr(g),
                                         ops._ptr(V), float(alpha), float(step_size), 1, ops._stream())
            check(ops.lib, rc, "psgld.step")
            return pSGLDState(state.step + 1, pytree.unravel(th[0], layers), pytree.unravel(g[0], layers),
                              pytree.unravel(V[0], layers))

        return one_step

    def __new__(cls, grad_estimator_fn, alpha: float = 0.95):
        if not isinstance(grad_estimator_fn, grad_estimator):
            raise TypeError("grad_estimator_fn must be a pbnn_b200.mcmc.kernels.grad_estimator spec object")
        step = cls.kernel(grad_estimator_fn)

        def init_fn(position, data_batch):
            return cls.init(position, data_batch, grad_estimator_fn)

        def step_fn(rng_key, state, data_batch, step_size):
            return step(rng_key, state, data_batch, step_size, alpha)

        return SamplingAlgorithm(init_fn, step_fn)
