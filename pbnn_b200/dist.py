"""Multi-GPU plumbing: chains / ensemble members shard contiguously over ranks, one process per GPU.

Nothing on the sampling path communicates.  The only exchange step is the pooling of per-rank predictive moments:
an ``all_gather`` of ``(sum_s f, sum_s f^2, count)`` followed by the pooled mean / variance.  Keys are derived as
``split(PRNGKey(seed), C_total)[c]`` -- counter based -- so a chain's trajectory does not depend on the world size.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(num_chains: int, rank: int = None, world_size: int = None) -> Tuple[int, int]:
    """Contiguous block partition: rank r owns chains [lo, hi); sizes differ by at most one."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(num_chains), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_chain_keys(ops, rng_key, num_chains: int, rank: int = None, world_size: int = None):
    """Keys of this rank's chains: ``split(rng_key, num_chains)[lo:hi]`` (independent of the world size)."""
    lo, hi = shard_range(num_chains, rank, world_size)
    return ops.split(rng_key, num_chains)[0, lo:hi].contiguous()


def pooled_moments(local_sum: torch.Tensor, local_sumsq: torch.Tensor, local_count: int):
    """all_gather the per-rank (sum, sumsq, count) and return the pooled (mean, var[ddof=0], count)."""
    rank, ws = world()
    payload = torch.cat([local_sum.reshape(-1), local_sumsq.reshape(-1),
                         torch.tensor([float(local_count)], dtype=local_sum.dtype, device=local_sum.device)])
    if ws > 1:
        parts = [torch.empty_like(payload) for _ in range(ws)]
        dist.all_gather(parts, payload)  # NCCL over NVLink on GPUs, gloo in the CPU tests
        total = torch.stack(parts).sum(0)  # fixed rank order => deterministic
    else:
        total = payload
    n = local_sum.numel()
    count = float(total[-1].item())
    mean = (total[:n] / count).reshape(local_sum.shape)
    var = (total[n:2 * n] / count).reshape(local_sum.shape) - mean * mean
    return mean, var.clamp_min(0), int(round(count))


def predictive_moments_sharded(ops, spec, local_samples, X_test):
    """Predictive mean / variance over the samples of ALL ranks; every rank returns the same result."""
    S = int(local_samples.reshape(-1, local_samples.shape[-1]).shape[0])
    if S > 0:
        r = ops.predict(spec, local_samples, X_test, want_preds=False, want_moments=True)
        s, q = r["sum"], r["sumsq"]
    else:
        M = int(X_test.shape[0])
        s = ops._zeros((M, spec.layers[-1]), ops.f32)
        q = ops._zeros((M, spec.layers[-1]), ops.f32)
    return pooled_moments(s, q, S)
