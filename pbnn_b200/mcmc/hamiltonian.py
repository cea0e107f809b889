"""Hamiltonian-family wrappers with the signatures of src/pbnn/mcmc/hamiltonian.py (hmc :31-39, sghmc :80-91)."""
from __future__ import annotations

import torch

from ..logprobs import FullBatchLogPosterior, require_specs
from ..ops import default_ops
from ._common import chain_layout, check_flags, make_predict_fn, make_ravel_fn, positions_tree, setup


def hmc(logprob_fn, init_positions, num_samples, step_size, inverse_mass_matrix, num_integration_steps, rng_key, *,
        thin=1, burnin=0, return_info=False):
    """pbnn.mcmc.hmc (hamiltonian.py:31-77) over blackjax.hmc.  ``logprob_fn`` must be a
    ``FullBatchLogPosterior`` (the closure of benchmark/pipelines/hmc.py:59-62 as a spec object)."""
    if not isinstance(logprob_fn, FullBatchLogPosterior):
        raise TypeError("logprob_fn must be a pbnn_b200.logprobs.FullBatchLogPosterior (got %r); arbitrary "
                        "callables cannot be compiled to the CUDA path" % (type(logprob_fn).__name__,))
    ops = default_ops()
    flat, layers, keys, batched = chain_layout(init_positions, rng_key)
    spec = logprob_fn.flat_spec()
    if tuple(spec.layers) != tuple(layers):
        raise ValueError(f"init_positions widths {layers} do not match the network {spec.layers}")
    res = ops.hmc(spec, logprob_fn.X, logprob_fn.y, flat, keys, int(num_samples), float(step_size),
                  inverse_mass_matrix, int(num_integration_steps), thin=thin, burnin=burnin, info=return_info)
    check_flags(res, "hmc")
    out = (positions_tree(res["trace"], layers, batched), make_ravel_fn(layers), make_predict_fn(layers))
    if return_info:
        info = {"acceptance_rate": res["accept_count"].to(torch.float32) / max(1, int(num_samples)),
                "delta_energy": res["delta"], "is_accepted": res["accept"], "logdensity": res["logp"],
                "flags": res["flags"]}
        return out + (info,)
    return out


def sghmc(X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, step_size, num_iterations,
          num_integration_steps, rng_key, *, thin=1, burnin=0, minibatch_mode="reference_fixed", alpha=0.01, beta=0.0,
          momentum_resampling="unit"):
    """pbnn.mcmc.sghmc (hamiltonian.py:80-140) over blackjax.sghmc(grad_fn, L) (alpha=0.01, beta=0 defaults).

    ``momentum_resampling``: how the fresh momentum of every iteration is drawn -- "unit" (N(0, 1), the restatement of
    SURVEY.md 2b), "sigma_sqrt_step" (N(0, step_size), the SGHMC paper) or "mu_sqrt_step" (N(sqrt(step_size), 1)), or
    the four coefficients (mu, mu_sqrt_step, sigma, sigma_sqrt_step).  blackjax 1.2.5 is not vendored with the
    reference, so the variant it implements is a parameter here, not a constant."""
    ops, spec, flat, layers, keys, batched = setup(loglikelihood_fn, logprior_fn, init_positions, rng_key)
    res = ops.sghmc(spec, X, y, flat, keys, int(batch_size), step_size, int(num_iterations),
                    int(num_integration_steps), alpha=alpha, beta=beta, minibatch_mode=minibatch_mode, thin=thin,
                    burnin=burnin, momentum_resampling=momentum_resampling)
    check_flags(res, "sghmc")
    return positions_tree(res["trace"], layers, batched), make_ravel_fn(layers), make_predict_fn(layers)


def sghmc_cv(X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, step_size, num_iterations,
             centering_positions, num_integration_steps, rng_key, *, thin=1, burnin=0, momentum_resampling="unit"):
    """pbnn.mcmc.sghmc_cv (hamiltonian.py:207-273)."""
    from .langevin import _cv_terms
    ops, spec, flat, layers, keys, batched = setup(loglikelihood_fn, logprior_fn, init_positions, rng_key)
    cflat, _, _, _ = chain_layout(centering_positions, rng_key)
    g_full, g_est, idx = _cv_terms(ops, spec, X, y, cflat, keys, batch_size)
    res = ops.sghmc(spec, X, y, flat, keys, int(batch_size), step_size, int(num_iterations),
                    int(num_integration_steps), idx=idx, cv=(g_full, g_est), thin=thin, burnin=burnin,
                    momentum_resampling=momentum_resampling)
    check_flags(res, "sghmc_cv")
    return positions_tree(res["trace"], layers, batched), make_ravel_fn(layers), make_predict_fn(layers)


def sghmc_svrg(X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, step_size, centering_positions,
               num_cv_iterations, num_svrg_iterations, num_integration_steps, rng_key, *,
               momentum_resampling="unit"):
    """pbnn.mcmc.sghmc_svrg (hamiltonian.py:276-359)."""
    from .langevin import _cv_terms
    ops, spec, flat, layers, keys, batched = setup(loglikelihood_fn, logprior_fn, init_positions, rng_key)
    cflat, _, _, _ = chain_layout(centering_positions, rng_key)
    outer = ops.split(keys, int(num_svrg_iterations))
    traces, theta, centre = [], flat, cflat
    for i in range(int(num_svrg_iterations)):
        g_full, g_est, idx = _cv_terms(ops, spec, X, y, centre, keys, batch_size)
        res = ops.sghmc(spec, X, y, theta, outer[:, i, :].contiguous(), int(batch_size), step_size,
                        int(num_cv_iterations), int(num_integration_steps), idx=idx, cv=(g_full, g_est),
                        momentum_resampling=momentum_resampling)
        traces.append(res["trace"])
        theta = centre = res["theta"]
    check_flags(res, "sghmc_svrg")
    trace = torch.cat(traces, dim=1)
    return positions_tree(trace, layers, batched), make_ravel_fn(layers), make_predict_fn(layers)
