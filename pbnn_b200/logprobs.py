"""Density *spec objects* for the CUDA path (mirror of benchmark/pipelines/logprobs.py:6-17).

pbnn passes ``loglikelihood_fn`` / ``logprior_fn`` / ``logprob_fn`` as Python closures that JAX differentiates.
A hand-written kernel can only honour a *described* density, so the drop-in takes these spec objects in the
same argument slots.  They are also callables (evaluated with the CUDA kernels) so user code that calls them
keeps working; any other callable is rejected with ``TypeError``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from . import pytree
from .models import MLP
from .ops import FlatSpec, default_ops


@dataclass(frozen=True)
class StdNormalPrior:
    """``logprior_fn``: ``sum_j norm.logpdf(theta_j / scale)`` (logprobs.py:6-9; scale = 1 in the reference)."""

    scale: float = 1.0

    def __call__(self, parameters):
        flat = pytree.ravel(parameters).to(torch.float32)
        t = flat / self.scale
        return (-(t * t) / 2 - math.log(math.sqrt(2 * math.pi)) - math.log(self.scale)).sum(-1)


@dataclass(frozen=True)
class GaussianMLPLikelihood:
    """``homoscedastic_loglike_fn(parameters, data, network, noise_level)`` with network / noise_level bound
    (logprobs.py:12-17; ``partial(...)`` at benchmark/pipelines/experiments.py:149-166)."""

    network: MLP
    noise_level: float

    def flat_spec(self, d: int, prior: StdNormalPrior = StdNormalPrior()) -> FlatSpec:
        return FlatSpec(self.network.widths(d), self.network.act_code, float(self.noise_level), float(prior.scale))

    def __call__(self, parameters, data):
        X, y = data
        f = self.network.apply({"params": parameters}, X)
        y = default_ops()._as(y, torch.float32).reshape(f.shape)
        return -(0.5 * (y - f) ** 2 / self.noise_level ** 2).sum()


@dataclass(frozen=True)
class HeteroscedasticGaussianMLPLikelihood(GaussianMLPLikelihood):
    """``heteroscedastic_loglike_fn(parameters, data, network)`` (logprobs.py:37-42): the network has two outputs
    ``(mu, log sigma^2)``; ``-0.5 log sigma^2 - 0.5 (y - mu)^2 / sigma^2`` per datum (Experiment 3,
    benchmark/pipelines/experiments.py:149-151)."""

    noise_level: float = 1.0  # unused

    def flat_spec(self, d: int, prior: StdNormalPrior = StdNormalPrior()) -> FlatSpec:
        from ._lib import LIK_HETEROSCEDASTIC
        widths = self.network.widths(d)
        if widths[-1] != 2:
            raise ValueError("the heteroscedastic likelihood needs a network with out_features=2")
        return FlatSpec(widths, self.network.act_code, 1.0, float(prior.scale), LIK_HETEROSCEDASTIC)

    def __call__(self, parameters, data):
        X, y = data
        f = self.network.apply({"params": parameters}, X)
        f = f.reshape(-1, 2)
        y = default_ops()._as(y, torch.float32).reshape(-1)
        mu, ls = f[:, 0], f[:, 1]
        return (-0.5 * ls - 0.5 * (y - mu) ** 2 * torch.exp(-ls)).sum()


def heteroscedastic_loglike_fn(network: MLP) -> HeteroscedasticGaussianMLPLikelihood:
    return HeteroscedasticGaussianMLPLikelihood(network)


def homoscedastic_loglike_fn(network: MLP, noise_level: float) -> GaussianMLPLikelihood:
    return GaussianMLPLikelihood(network, noise_level)


logprior_fn = StdNormalPrior()


@dataclass(frozen=True)
class FullBatchLogPosterior:
    """``logprob_fn`` of the HMC pipeline: ``logprior + sum_i loglik_i`` over the full data
    (benchmark/pipelines/hmc.py:59-62)."""

    loglikelihood_fn: GaussianMLPLikelihood
    logprior_fn: StdNormalPrior
    X: object
    y: object

    def flat_spec(self) -> FlatSpec:
        return self.loglikelihood_fn.flat_spec(int(np.shape(self.X)[-1]), self.logprior_fn)

    def __call__(self, parameters):
        return self.logprior_fn(parameters) + self.loglikelihood_fn(parameters, (self.X, self.y))


def require_specs(loglikelihood_fn, logprior_fn):
    if not isinstance(loglikelihood_fn, GaussianMLPLikelihood):
        raise TypeError(
            "pbnn_b200 runs hand-written kernels for a described density: loglikelihood_fn must be a "
            "pbnn_b200.logprobs.GaussianMLPLikelihood (got %r); arbitrary callables cannot be compiled" %
            (type(loglikelihood_fn).__name__,))
    if not isinstance(logprior_fn, StdNormalPrior):
        raise TypeError("logprior_fn must be a pbnn_b200.logprobs.StdNormalPrior (got %r)" %
                        (type(logprior_fn).__name__,))
