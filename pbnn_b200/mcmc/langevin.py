"""SGLD-family wrappers with the signatures of src/pbnn/mcmc/langevin.py (sgld :21-31, pSGLD :81-92,
scheduled_sgld :143-153).  Same positional arguments; keyword-only extensions: ``thin``, ``burnin``,
``minibatch_mode`` ("reference_fixed" reproduces the single minibatch the reference's traced scan re-uses --
SURVEY.md quirk Q1 -- "per_step" draws a fresh minibatch every iteration as the docstrings promise), and
chain batching (``rng_key`` of shape (C, 2) + ``init_positions`` leaves with a leading C).
"""
from __future__ import annotations

from ..logprobs import require_specs
from ..ops import default_ops
from ._common import chain_layout, make_predict_fn, make_ravel_fn, positions_tree


def _run(algo, X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, step_size, num_iterations, rng_key,
         thin, burnin, minibatch_mode, **extra):
    require_specs(loglikelihood_fn, logprior_fn)
    ops = default_ops()
    flat, layers, keys, batched = chain_layout(init_positions, rng_key)
    spec = loglikelihood_fn.flat_spec(layers[0], logprior_fn)
    if tuple(spec.layers) != tuple(layers):
        raise ValueError(f"init_positions widths {layers} do not match the network {spec.layers}")
    fn = getattr(ops, algo)
    res = fn(spec, X, y, flat, keys, int(batch_size), step_size, int(num_iterations), minibatch_mode=minibatch_mode,
             thin=thin, burnin=burnin, **extra)
    return positions_tree(res["trace"], layers, batched), make_ravel_fn(layers), make_predict_fn(layers)


def sgld(X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, step_size, num_iterations, rng_key, *,
         thin=1, burnin=0, minibatch_mode="reference_fixed"):
    """pbnn.mcmc.sgld (langevin.py:21-78): returns (positions, ravel_fn, predict_fn)."""
    return _run("sgld", X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, step_size, num_iterations,
                rng_key, thin, burnin, minibatch_mode)


def pSGLD(X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, step_size, num_iterations,
          preconditioning_factor, rng_key, *, thin=1, burnin=0, minibatch_mode="reference_fixed"):
    """pbnn.mcmc.pSGLD (langevin.py:81-140)."""
    require_specs(loglikelihood_fn, logprior_fn)
    ops = default_ops()
    flat, layers, keys, batched = chain_layout(init_positions, rng_key)
    spec = loglikelihood_fn.flat_spec(layers[0], logprior_fn)
    res = ops.psgld(spec, X, y, flat, keys, int(batch_size), step_size, int(num_iterations),
                    float(preconditioning_factor), minibatch_mode=minibatch_mode, thin=thin, burnin=burnin)
    return positions_tree(res["trace"], layers, batched), make_ravel_fn(layers), make_predict_fn(layers)


def scheduled_sgld(X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, schedule_fn, num_iterations,
                   rng_key, *, thin=1, burnin=0, minibatch_mode="reference_fixed"):
    """pbnn.mcmc.scheduled_sgld (langevin.py:143-204): ``step_size_t = schedule_fn(t)``, t = 1..T (:186-196)."""
    if not callable(schedule_fn):
        raise TypeError("schedule_fn must be callable")
    return _run("sgld", X, y, loglikelihood_fn, logprior_fn, init_positions, batch_size, schedule_fn, num_iterations,
                rng_key, thin, burnin, minibatch_mode)
