"""Flax-layout parameter pytrees <-> flat vectors (``jax.flatten_util.ravel_pytree`` order).

The reference moves ``{"Dense_k": {"bias": (out,), "kernel": (in, out)}}`` dict-of-dicts around
(src/pbnn/models.py:16-35) and flattens them with ``ravel_pytree`` (src/pbnn/mcmc/langevin.py:72-73):
dict keys sorted, leaves row-major -- so inside every layer the bias precedes the kernel.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np
import torch


def layer_names(n_layers: int):
    return [f"Dense_{k}" for k in range(n_layers)]


def flat_offsets(layers: Sequence[int]):
    """[(name, bias_off, kernel_off, in, out)] in Dense_0..Dense_{L-1} order."""
    n = len(layers) - 1
    if n > 10:
        raise ValueError("more than 10 Dense layers: string-sorted key order (Dense_10 < Dense_2) is not supported")
    out, off = [], 0
    for k in range(n):
        i, o = int(layers[k]), int(layers[k + 1])
        out.append((f"Dense_{k}", off, off + o, i, o))
        off += o * (i + 1)
    return out


def widths_of(tree: Dict) -> Tuple[int, ...]:
    """Recover (d, h1, ..., out) from a parameter pytree (leaves may carry leading batch dims)."""
    names = sorted(tree.keys(), key=lambda s: int(s.split("_")[1]))
    if names != layer_names(len(names)):
        raise TypeError(f"expected Flax auto-named Dense layers Dense_0..Dense_k, got {sorted(tree.keys())}")
    w = []
    for k, nme in enumerate(names):
        kern = tree[nme]["kernel"]
        i, o = int(kern.shape[-2]), int(kern.shape[-1])
        if k == 0:
            w.append(i)
        elif w[-1] != i:
            raise ValueError("inconsistent layer widths in the parameter pytree")
        w.append(o)
    return tuple(w)


def _to_tensor(a):
    if isinstance(a, torch.Tensor):
        return a
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32)))


def ravel(tree: Dict, layers: Sequence[int] = None) -> torch.Tensor:
    """pytree with leaves (*batch, ...) -> (*batch, P) tensor."""
    layers = tuple(layers) if layers is not None else widths_of(tree)
    parts = []
    for name, _, _, i, o in flat_offsets(layers):
        b = _to_tensor(tree[name]["bias"])
        w = _to_tensor(tree[name]["kernel"])
        batch = tuple(w.shape[:-2])
        parts.append(b.reshape(batch + (o,)))
        parts.append(w.reshape(batch + (i * o,)))
    return torch.cat(parts, dim=-1)


def unravel(flat: torch.Tensor, layers: Sequence[int]) -> Dict:
    """(*batch, P) -> pytree with leaves (*batch, out) / (*batch, in, out)."""
    batch = tuple(flat.shape[:-1])
    tree = {}
    for name, bo, wo, i, o in flat_offsets(layers):
        tree[name] = {"bias": flat[..., bo:bo + o], "kernel": flat[..., wo:wo + i * o].reshape(batch + (i, o))}
    return tree
