"""``pbnn.models.MLP`` mirror (src/pbnn/models.py:9-35), generalised to any number of hidden layers.

The reference MLP is always ``Dense(H) -> relu -> Dense(H) -> relu -> Dense(out)``; the BASELINE configs also use
one-hidden-layer nets and the notebooks use tanh, so ``num_hidden`` / ``hidden_layers`` and ``activation`` are
keyword extensions.  ``init`` / ``apply`` keep the Flax calling convention
(``network.init(rng, x)["params"]``, ``network.apply({"params": p}, X)``).
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import pytree
from ._lib import ACT_RELU, ACT_TANH

_ACTS = {"relu": ACT_RELU, "tanh": ACT_TANH, ACT_RELU: ACT_RELU, ACT_TANH: ACT_TANH}


def _flax_fold_hash(parts) -> int:
    """flax.core.scope._fold_in_static: sha1 over the path parts (str utf-8 / int big-endian minimal bytes),
    first 4 digest bytes big-endian.  Restated from memory of flax 0.11 -- NOT verifiable here."""
    m = hashlib.sha1()
    for x in parts:
        if isinstance(x, str):
            m.update(x.encode("utf-8"))
        else:
            m.update(int(x).to_bytes((int(x).bit_length() + 7) // 8, byteorder="big"))
    return int.from_bytes(m.digest()[:4], byteorder="big")


@dataclass(frozen=True)
class MLP:
    hidden_features: int
    out_features: int
    num_hidden: int = 2
    activation: str = "relu"
    hidden_layers: Optional[Tuple[int, ...]] = None  # explicit widths override hidden_features/num_hidden

    def __post_init__(self):
        if self.activation not in _ACTS:
            raise ValueError("activation must be 'relu' or 'tanh'")

    @property
    def act_code(self) -> int:
        return _ACTS[self.activation]

    def widths(self, d: int) -> Tuple[int, ...]:
        hid = tuple(self.hidden_layers) if self.hidden_layers is not None else (self.hidden_features,) * self.num_hidden
        return (int(d),) + tuple(int(h) for h in hid) + (int(self.out_features),)

    @classmethod
    def from_widths(cls, widths: Sequence[int], activation: str = "relu") -> "MLP":
        widths = tuple(int(w) for w in widths)
        hid = widths[1:-1]
        return cls(hid[0] if hid else 0, widths[-1], len(hid), activation, hid)

    # -- Flax-style API ---------------------------------------------------------------------------
    def init(self, rng_key, x):
        """``nn.initializers.normal()`` (stddev 0.01) for kernels AND biases (models.py:19-22).  Each leaf draws
        ``0.01 * normal(fold_in(rng, hash(path)))``; the path hash follows Flax's scheme as restated above."""
        from .ops import default_ops
        ops = default_ops()
        d = int(np.shape(x)[-1])
        layers = self.widths(d)
        base = ops._keys(rng_key)
        tree = {}
        for name, _, _, i, o in pytree.flat_offsets(layers):
            leaves = {}
            for cnt, (leaf, n, shape) in enumerate((("kernel", i * o, (i, o)), ("bias", o, (o,)))):
                h = _flax_fold_hash((name, cnt))
                k = ops.fold_in(base, h)
                leaves[leaf] = (0.01 * ops.normal(k, n)).reshape(shape)
            tree[name] = leaves
        return {"params": tree}

    def apply(self, variables, X):
        from .ops import FlatSpec, default_ops
        ops = default_ops()
        params = variables["params"]
        layers = pytree.widths_of(params)
        flat = pytree.ravel(params, layers).reshape(1, -1)
        Xt = ops._as(X, ops.f32)
        single = Xt.ndim == 1
        out = ops.predict(FlatSpec(layers, self.act_code), flat, Xt.reshape(-1, layers[0]))["preds"][0]
        return out[0] if single else out
