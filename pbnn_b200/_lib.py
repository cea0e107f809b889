"""ctypes binding of ``libpbnn_b200.so`` (the C ABI declared in ``include/pbnn_b200.h``).

There is no fallback: if the CUDA library has not been built (``python __graft_entry__.py build``)
importing the ops fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

PBNN_MAX_LAYERS = 8
ACT_RELU, ACT_TANH = 0, 1
OP_SGLD, OP_PSGLD, OP_SGHMC, OP_ADAM, OP_GRAD, OP_HMC = range(6)
ENGINE_TILED, ENGINE_FUSED_H1, ENGINE_TCGEN05, ENGINE_TILED_GLOBAL_STATE, ENGINE_TCGEN05_MLP2, ENGINE_WARP_H1, ENGINE_WARP_MLP2, ENGINE_CTA4_H1 = range(8)
MB_REFERENCE_FIXED, MB_PER_STEP = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libpbnn_b200.so")


class MlpSpecC(ctypes.Structure):
    """``pbnn_mlp_spec`` of include/pbnn_b200.h."""

    _fields_ = [
        ("n_layers", c_int32),
        ("dims", c_int32 * (PBNN_MAX_LAYERS + 1)),
        ("activation", c_int32),
        ("noise_level", c_float),
        ("prior_scale", c_float),
        ("likelihood", c_int32),
    ]


class PbnnError(RuntimeError):
    pass


_P = c_void_p
_SPEC = POINTER(MlpSpecC)

# name -> (restype, argtypes); every symbol include/pbnn_b200.h declares
PROTOTYPES = {
    "pbnn_version": (c_int, []),
    "pbnn_last_error": (c_char_p, []),
    "pbnn_set_option": (c_int, [c_char_p, c_int]),
    "pbnn_get_option": (c_int, [c_char_p, POINTER(c_int)]),
    "pbnn_reset_options": (None, []),
    "pbnn_num_params": (c_int, [_SPEC]),
    "pbnn_plan_query": (c_int, [_SPEC, c_int, c_int, POINTER(c_int32 * 4)]),
    "pbnn_plan_engine": (c_int, [_SPEC, c_int, c_int, c_int]),
    "pbnn_threefry_split": (c_int, [_P, c_int, c_int, _P, _P]),
    "pbnn_threefry_fold_in": (c_int, [_P, c_int, ctypes.c_uint32, _P, _P]),
    "pbnn_threefry_bits": (c_int, [_P, c_int, c_int, _P, _P]),
    "pbnn_threefry_uniform": (c_int, [_P, c_int, c_int, _P, _P]),
    "pbnn_threefry_normal": (c_int, [_P, c_int, c_int, _P, _P]),
    "pbnn_minibatch_indices": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P]),
    "pbnn_permutation": (c_int, [_P, c_int, c_int, _P, _P, _P]),
    "pbnn_workspace_bytes": (c_int64, [_SPEC, c_int, c_int, c_int]),
    "pbnn_predict_workspace_bytes": (c_int64, [_SPEC, c_int, c_int, c_int]),
    "pbnn_sgld_run": (c_int, [_SPEC, _P, _P, c_int, _P, c_int, c_int, c_int, _P, _P, _P, c_int, c_int64, c_int, c_int,
                              c_int, c_int, _P, _P, _P, _P, c_float, _P, c_int64, _P]),
    "pbnn_logposterior_grad": (c_int, [_SPEC, _P, _P, c_int, _P, c_int, _P, _P, _P, c_int64, _P]),
    "pbnn_psgld_run": (c_int, [_SPEC, _P, _P, c_int, _P, c_int, c_int, c_int, _P, _P, _P, _P, c_int, c_float, _P, c_int,
                               c_int64, c_int, c_int, c_int, c_int, _P, _P, _P, c_int64, _P]),
    "pbnn_psgld_init": (c_int, [_SPEC, _P, _P, c_int, c_int, _P, _P, _P, c_int, _P]),
    "pbnn_psgld_step": (c_int, [_SPEC, _P, _P, c_int, c_int, _P, _P, _P, _P, c_float, c_float, c_int, _P]),
    "pbnn_sghmc_run": (c_int, [_SPEC, _P, _P, c_int, _P, c_int, c_int, c_int, _P, _P, c_int, c_float, c_float, c_float,
                               c_float, c_float, c_float, _P, c_int,
                               c_int64, c_int, c_int, c_int, c_int, _P, _P, _P, _P, c_float, _P, c_int64, _P]),
    "pbnn_hmc_run": (c_int, [_SPEC, _P, _P, c_int, _P, _P, _P, c_float, c_int, c_int64, c_int, c_int, c_int, c_int, _P,
                             _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "pbnn_adam_ensemble_run": (c_int, [_SPEC, _P, _P, c_int, _P, c_int, c_int, _P, _P, _P, c_int, c_float, c_float,
                                       c_float, c_float, c_int, _P, _P, c_int64, _P]),
    "pbnn_predict_chunks": (c_int, [c_int]),
    "pbnn_predict": (c_int, [_SPEC, _P, c_int, _P, c_int, _P, _P, _P, _P, c_int64, _P]),
    "pbnn_thinning": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P]),
    "pbnn_grad_estimator": (c_int, [_SPEC, _P, _P, c_int, _P, c_int, _P, c_int, _P, _P, _P, c_int64, _P]),
}


def load_library(path: str | None = None) -> ctypes.CDLL:
    path = path or os.environ.get("PBNN_B200_LIB", DEFAULT_LIB)
    if not os.path.exists(path):
        raise ImportError(
            f"pbnn_b200: native library not found at {path}. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` from the repo root "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


LIK_HOMOSCEDASTIC, LIK_HETEROSCEDASTIC = 0, 1


def make_spec(layers, activation: int, noise_level: float, prior_scale: float = 1.0, likelihood: int = 0) -> MlpSpecC:
    layers = [int(v) for v in layers]
    if not (2 <= len(layers) <= PBNN_MAX_LAYERS + 1):
        raise ValueError(f"an MLP needs 1..{PBNN_MAX_LAYERS} Dense layers, got widths {layers}")
    s = MlpSpecC()
    s.n_layers = len(layers) - 1
    for i, v in enumerate(layers):
        s.dims[i] = v
    s.activation = int(activation)
    s.noise_level = float(noise_level)
    s.prior_scale = float(prior_scale)
    s.likelihood = int(likelihood)
    return s


def check(lib: ctypes.CDLL, rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.pbnn_last_error().decode("utf-8", "replace")
        exc = {-1: ValueError, -2: MemoryError}.get(rc, PbnnError)
        raise exc(f"{what}: {msg} (code {rc})")
