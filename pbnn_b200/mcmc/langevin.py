"""SGLD-family wrappers with the signatures of src/pbnn/mcmc/langevin.py (sgld :21-31, pSGLD :81-92,
scheduled_sgld :143-153).  Same positional arguments; keyword-only extensions: ``thin``, ``burnin``,
``minibatch_mode`` ("reference_fixed" reproduces the single minibatch the reference's traced scan re-uses --
SURVEY.md quirk Q1 -- "per_step" draws a fresh minibatch every iteration as the docstrings promise), and
chain batching (``rng_key`` of shape (C, 2) + ``init_positions`` leaves with a leading C).
"""
from __future__ import annotations

from ..logprobs import require_specs
from ..ops import default_ops
from ._common import chain_layout, check_flags, make_predict_fn, make_ravel_fn, positions_tree, setup


def _run(algo, X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, step_size, num_iterations, rng_key,
         thin, burnin, minibatch_mode, **extra):
    ops, spec, flat, layers, keys, batched = setup(loglikelihood_fn, logprior_fn, init_positions, rng_key)
    fn = getattr(ops, algo)
    res = fn(spec, X, y, flat, keys, int(batch_size), step_size, int(num_iterations), minibatch_mode=minibatch_mode,
             thin=thin, burnin=burnin, **extra)
    check_flags(res, algo)
    return positions_tree(res["trace"], layers, batched), make_ravel_fn(layers), make_predict_fn(layers)


def sgld(X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, step_size, num_iterations, rng_key, *,
         thin=1, burnin=0, minibatch_mode="reference_fixed"):
    """pbnn.mcmc.sgld (langevin.py:21-78): returns (positions, ravel_fn, predict_fn)."""
    return _run("sgld", X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, step_size, num_iterations,
                rng_key, thin, burnin, minibatch_mode)


def pSGLD(X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, step_size, num_iterations,
          preconditioning_factor, rng_key, *, thin=1, burnin=0, minibatch_mode="reference_fixed"):
    """pbnn.mcmc.pSGLD (langevin.py:81-140)."""
    ops, spec, flat, layers, keys, batched = setup(loglikelihood_fn, logprior_fn, init_positions, rng_key)
    res = ops.psgld(spec, X, y, flat, keys, int(batch_size), step_size, int(num_iterations),
                    float(preconditioning_factor), minibatch_mode=minibatch_mode, thin=thin, burnin=burnin)
    check_flags(res, "pSGLD")
    return positions_tree(res["trace"], layers, batched), make_ravel_fn(layers), make_predict_fn(layers)


def scheduled_sgld(X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, schedule_fn, num_iterations,
                   rng_key, *, thin=1, burnin=0, minibatch_mode="reference_fixed"):
    """pbnn.mcmc.scheduled_sgld (langevin.py:143-204): ``step_size_t = schedule_fn(t)``, t = 1..T (:186-196)."""
    if not callable(schedule_fn):
        raise TypeError("schedule_fn must be callable")
    return _run("sgld", X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, schedule_fn, num_iterations,
                rng_key, thin, burnin, minibatch_mode)


def _cv_terms(ops, spec, X, y, centering_flat, batch_keys, batch_size):
    """(grad_center, grad_center_est, idx): the two constant terms of blackjax's control_variates estimator and the
    loop's fixed minibatch (generator draw #2 of ``batch_keys``)."""
    N = int(X.shape[0]) if hasattr(X, "shape") else len(X)
    idx = ops.minibatch_indices(batch_keys, 2, int(batch_size), N)
    _, g_full = ops.logposterior_grad(spec, X, y, centering_flat)
    g_est, _ = ops.grad_estimator(spec, X, y, centering_flat, idx)
    return g_full, g_est, idx


def sgld_cv(X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, step_size, num_iterations,
            centering_positions, rng_key, *, thin=1, burnin=0):
    """pbnn.mcmc.sgld_cv (langevin.py:207-270): SGLD with blackjax's control-variates gradient estimator."""
    ops, spec, flat, layers, keys, batched = setup(loglikelihood_fn, logprior_fn, init_positions, rng_key)
    cflat, _, _, _ = chain_layout(centering_positions, rng_key)
    g_full, g_est, idx = _cv_terms(ops, spec, X, y, cflat, keys, batch_size)
    res = ops.sgld(spec, X, y, flat, keys, int(batch_size), step_size, int(num_iterations), idx=idx,
                   cv=(g_full, g_est), thin=thin, burnin=burnin)
    check_flags(res, "sgld_cv")
    return positions_tree(res["trace"], layers, batched), make_ravel_fn(layers), make_predict_fn(layers)


def sgld_svrg(X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, step_size, centering_positions,
              num_cv_iterations, num_svrg_iterations, rng_key):
    """pbnn.mcmc.sgld_svrg (langevin.py:273-354): ``num_svrg_iterations`` re-centred blocks of ``num_cv_iterations``
    control-variate SGLD steps; positions of all blocks concatenated."""
    import torch
    ops, spec, flat, layers, keys, batched = setup(loglikelihood_fn, logprior_fn, init_positions, rng_key)
    cflat, _, _, _ = chain_layout(centering_positions, rng_key)
    outer = ops.split(keys, int(num_svrg_iterations))  # (C, n_svrg, 2)
    traces, theta, centre = [], flat, cflat
    for i in range(int(num_svrg_iterations)):
        g_full, g_est, idx = _cv_terms(ops, spec, X, y, centre, keys, batch_size)
        res = ops.sgld(spec, X, y, theta, outer[:, i, :].contiguous(), int(batch_size), step_size,
                       int(num_cv_iterations), idx=idx, cv=(g_full, g_est))
        traces.append(res["trace"])
        theta = centre = res["theta"]
    check_flags(res, "sgld_svrg")
    trace = torch.cat(traces, dim=1)
    return positions_tree(trace, layers, batched), make_ravel_fn(layers), make_predict_fn(layers)


def cyclical_sgld(X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, step_size, num_cycles, rng_key,
                  num_sgd_steps=1_000, num_sgld_steps=20_000, burnin_sgld=10_000):
    """pbnn.mcmc.cyclical_sgld (langevin.py:364-470): per cycle an SGD exploration phase (a temperature-0 Langevin
    step, step size = cosine schedule) restarted from ``init_positions`` (the reference scans from ``init_state`` in
    every cycle) followed by an SGLD sampling phase whose first ``burnin_sgld`` positions are dropped.
    Assumption (not verifiable without JAX): every phase uses the first draw of its own minibatch generator, i.e. the
    scan body is re-traced for every phase."""
    import math

    import torch
    ops, spec, flat, layers, keys, batched = setup(loglikelihood_fn, logprior_fn, init_positions, rng_key)
    N = int(X.shape[0]) if hasattr(X, "shape") else len(X)
    cycle_length = int(num_sgd_steps) + int(num_sgld_steps)

    def schedule(n):
        return [0.5 * (math.cos(math.pi * (it % cycle_length) / cycle_length) + 1.0) * float(step_size)
                for it in range(n)]

    key = keys
    kept = []
    for _ in range(int(num_cycles)):
        key = ops.split(key, 2)[:, 1, :].contiguous()
        idx = ops.minibatch_indices(key, 1, int(batch_size), N)
        res = ops.sgld(spec, X, y, flat, key, int(batch_size), schedule(int(num_sgd_steps)), int(num_sgd_steps),
                       idx=idx, temperature=0.0, keep_trace=False)
        key = ops.split(key, 2)[:, 1, :].contiguous()
        idx = ops.minibatch_indices(key, 1, int(batch_size), N)
        res = ops.sgld(spec, X, y, res["theta"], key, int(batch_size), schedule(int(num_sgld_steps)),
                       int(num_sgld_steps), idx=idx, burnin=int(burnin_sgld))
        kept.append(res["trace"])
    check_flags(res, "cyclical_sgld")
    trace = torch.cat(kept, dim=1)
    return positions_tree(trace, layers, batched), make_ravel_fn(layers), make_predict_fn(layers)
