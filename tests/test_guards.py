"""Out-of-bounds WRITE detection without compute-sanitizer (the tool is closed on this GPU pool: the gpurun wrapper
refuses it -- profiles/r2_sanitizer_note.txt).  Every device buffer the operator layer hands to the C ABI (inputs,
outputs, traces, sampler state, the global workspace with the operand images and the noise ring) is carved out of a
larger allocation with 4 KiB guard bands of a byte pattern on both sides; after each call every band must be intact.
Races are covered by the bit-identical repeat tests (tests/test_real_shapes.py, tests/test_tc_engine.py): the mbarrier /
bulk-copy / TMEM hand-shakes of the tcgen05 kernels are deterministic by construction, so any race shows up as a bit flip."""
import numpy as np
import pytest
import torch

from pbnn_b200.ops import FlatSpec, Ops

pytestmark = pytest.mark.gpu
GUARD = 4096
PATTERN = 0xA5


class GuardedOps(Ops):
    def __init__(self, lib):
        super().__init__(lib)
        self.bases = []

    def _guarded(self, shape, dtype, fill=None):
        shape = tuple(int(s) for s in shape)
        n = int(np.prod(shape)) if shape else 1
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        pad = (-nbytes) % 1024
        base = torch.full((GUARD + nbytes + pad + GUARD,), PATTERN, dtype=torch.uint8, device=self._device())
        inner = base[GUARD:GUARD + nbytes].view(dtype).reshape(shape)
        if fill is not None:
            inner.fill_(fill)
        self.bases.append((base, nbytes))
        return inner

    def _empty(self, shape, dtype):
        return self._guarded(shape, dtype)

    def _zeros(self, shape, dtype):
        return self._guarded(shape, dtype, 0)

    def _as(self, a, dtype):
        t = super()._as(a, dtype)
        if t is None:
            return None
        g = self._guarded(t.shape, t.dtype)
        g.copy_(t)
        return g

    def check(self):
        torch.cuda.synchronize()
        for base, nbytes in self.bases:
            assert bool((base[:GUARD] == PATTERN).all().item()), "guard band BEFORE a buffer was overwritten"
            assert bool((base[GUARD + nbytes:] == PATTERN).all().item()), "guard band AFTER a buffer was overwritten"
        n = len(self.bases)
        self.bases = []
        return n


@pytest.fixture
def gops(cuda_ops):
    cuda_ops.reset_options()
    return GuardedOps(cuda_ops.lib)


def problem(gops, layers, N, C, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, layers[0])).astype(np.float32)
    y = rng.standard_normal((N, 1)).astype(np.float32)
    fs = FlatSpec(tuple(layers), 0, 0.1)
    keys = gops.split(np.array([[0, 11]], np.uint32), C)[0]
    th0 = 0.01 * gops.normal(keys, fs.num_params)
    return fs, X, y, keys, th0


@pytest.mark.parametrize("layers,B,N,C", [((9, 100, 100, 1), 128, 400, 3), ((90, 256, 256, 1), 128, 300, 2),
                                          ((20, 64, 64, 1), 64, 200, 150), ((70, 160, 160, 1), 77, 300, 2)])
def test_tc2_engine_stays_inside_its_buffers(gops, layers, B, N, C):
    fs, X, y, keys, th0 = problem(gops, layers, N, C)
    r = gops.sgld(fs, X, y, th0, keys, B, 1e-6, 3, thin=2)
    assert torch.isfinite(r["trace"]).all()
    r = gops.sgld(fs, X, y, th0, keys, B, 1e-6, 2, minibatch_mode="per_step", keep_trace=False)
    r = gops.psgld(fs, X, y, th0, keys, B, 1e-7, 2, 0.95)
    r = gops.sghmc(fs, X, y, th0, keys, B, 1e-7, 1, 2, momentum_resampling="sigma_sqrt_step")
    idx = np.random.default_rng(1).integers(0, N, (C, 2, B)).astype(np.int32)
    r = gops.adam_ensemble(fs, X, y, th0, idx, 1e-3)
    g, ll = gops.grad_estimator(fs, X, y, th0, idx[:, 0])
    m = gops.predict(fs, th0, X, want_preds=True, want_moments=True)  # tcgen05 predictive (image blocks in the workspace)
    assert torch.isfinite(m["mean"]).all()
    assert gops.check() > 10


@pytest.mark.parametrize("layers,N,C", [((13, 50, 1), 506, 5), ((13, 50, 1), 900, 3), ((5, 33, 1), 77, 300)])
def test_hmc_tc_engine_stays_inside_its_buffers(gops, layers, N, C):
    fs, X, y, keys, th0 = problem(gops, layers, N, C)
    r = gops.hmc(fs, X, y, th0, keys, 3, 1e-4, np.ones(fs.num_params, np.float32), 4, info=True)
    assert torch.isfinite(r["trace"]).all()
    assert gops.check() > 5


def test_fp32_engines_and_helpers_stay_inside_their_buffers(gops):
    fs, X, y, keys, th0 = problem(gops, (8, 50, 1), 1030, 4)
    gops.sgld(fs, X, y, th0, keys, 32, 1e-5, 5)
    gops.psgld(fs, X, y, th0, keys, 32, 1e-6, 3, 0.95)
    fs2, X2, y2, keys2, th2 = problem(gops, (8, 12, 12, 1), 333, 4)
    gops.sgld(fs2, X2, y2, th2, keys2, 32, 1e-5, 3)
    gops.hmc(fs2, X2, y2, th2, keys2, 2, 1e-4, np.ones(fs2.num_params, np.float32), 3)
    m = gops.predict(fs, th0, X, want_preds=True, want_moments=True)
    m = gops.predict(fs2, th2, X2, want_preds=True, want_moments=True)
    gops.permutation(keys, 1030)
    gops.minibatch_indices(keys, 2, 32, 1030)
    assert gops.check() > 10


@pytest.mark.parametrize("layers,B,C", [((8, 50, 1), 32, 300), ((13, 50, 1), 40, 5), ((15, 64, 1), 96, 3), ((3, 7, 1), 33, 4)])
def test_warp_per_chain_engine_stays_inside_its_buffers(gops, layers, B, C):
    """pbnn_warp.cuh: trace rows leave through a staging array in whole lines, the state arrays of pSGLD / Adam are read
    and written per slot -- every one of those global accesses must stay inside the caller's buffers."""
    gops.set_option("w1_min_chains", 1)
    fs, X, y, keys, th0 = problem(gops, layers, 400, C)
    r = gops.sgld(fs, X, y, th0, keys, B, 1e-6, 5, thin=2, burnin=1)
    assert torch.isfinite(r["trace"]).all()
    gops.sgld(fs, X, y, th0, keys, B, 1e-6, 3, minibatch_mode="per_step", keep_trace=False)
    gops.psgld(fs, X, y, th0, keys, B, 1e-7, 3, 0.95)
    gops.sghmc(fs, X, y, th0, keys, B, 1e-7, 2, 2, momentum_resampling="sigma_sqrt_step")
    idx = np.random.default_rng(1).integers(0, 400, (C, 2, B)).astype(np.int32)
    gops.adam_ensemble(fs, X, y, th0, idx, 1e-3)
    assert gops.check() > 10
    gops.reset_options()


@pytest.mark.parametrize("layers,B,C", [((8, 50, 50, 1), 32, 7), ((15, 64, 64, 1), 32, 3), ((3, 33, 33, 1), 17, 4),
                                        ((1, 1, 1, 1), 1, 2)])
def test_four_warp_adam_kernel_stays_inside_its_buffers(gops, layers, B, C):
    """pbnn_warp2.cuh: the moments are updated in place in global memory, pair by pair; lanes beyond the net's width and
    the second unit of an odd width must neither read nor write outside theta / m / v."""
    fs, X, y, keys, th0 = problem(gops, layers, 300, C)
    idx = np.random.default_rng(2).integers(0, 300, (C, 4, B)).astype(np.int32)
    r = gops.adam_ensemble(fs, X, y, th0, idx, 1e-2)
    assert torch.isfinite(r["theta"]).all() and torch.isfinite(r["m"]).all() and torch.isfinite(r["v"]).all()
    assert gops.check() > 5
