"""``pbnn.utils.misc`` mirror: ``thinning_fn`` (src/pbnn/utils/misc.py:58-125) on the device, and the ``.npz`` wire
format the benchmark pipelines write after every chain (benchmark/pipelines/sgmcmc.py:213-232)."""
from __future__ import annotations

import numpy as np

from ..ops import default_ops


def thinning_fn(positions, size: int, memory_efficient: bool = False):
    """Indices (``size``,) of the greedy energy-distance thinning of ``positions`` (N, D), in the reference's order
    (picks 1..size-1, then the initial pick).  ``memory_efficient`` is accepted for signature compatibility: the
    device implementation never materialises the N x N kernel matrix."""
    del memory_efficient
    return default_ops().thinning(positions, int(size))


def save_results(path, positions, predictions, time):
    """``jnp.savez(path, positions=(S,P), predictions=(S,M), time=...)`` -- the format read by
    benchmark/pipelines/metrics_prediction_intervals.py."""
    to_np = lambda a: a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    np.savez(path, positions=to_np(positions), predictions=to_np(predictions), time=np.asarray(time))
