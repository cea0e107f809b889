"""GPU check of the EXPERIMENTAL tcgen05 building block for the wide-layer path.  Not on the
product path: it validates the operand layout, descriptors, TMEM epilogue and the 3xTF32 split on real hardware so that
the integration work starts from a tested primitive."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "pbnn_b200", "csrc", "experimental", "umma_tf32.cu")
LIB = os.path.join(ROOT, "pbnn_b200", "csrc", "experimental", "libpbnn_umma_test.so")


def _lib():
    if not os.path.exists(LIB) or os.path.getmtime(SRC) > os.path.getmtime(LIB):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                               "-shared", "-Xcompiler", "-fPIC", "-o", LIB, SRC])
    lib = ctypes.CDLL(LIB)
    lib.pbnn_umma_tf32_gemm.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 4 + [ctypes.c_void_p]
    return lib


def test_sass_has_tcgen05():
    _lib()
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass  # tcgen05.mma / tcgen05.ld (B200_PROFILING.md)


@pytest.mark.gpu
@pytest.mark.parametrize("K,N", [(8, 16), (32, 64), (64, 256), (48, 256)])
def test_tf32_umma_k_major_and_3x_split(K, N):
    import torch
    lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(K * 1000 + N)
    A = torch.randn(128, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g)
    trunc = lambda x: (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
    exact = A.double() @ B.double().t()
    for split3, ref, tol in ((0, trunc(A).double() @ trunc(B).double().t(), 1e-6), (1, exact, 5e-6)):
        D = torch.full((128, N), float("nan"), device="cuda")
        assert lib.pbnn_umma_tf32_gemm(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, N, 0, split3, None) == 0
        torch.cuda.synchronize()
        err = ((D.double() - ref).norm() / ref.norm()).item()
        assert err <= tol, (split3, err)
    # single-pass TF32 is ~8e-4 from the exact product: the 3-term split is what buys FP32 parity
    D = torch.empty((128, N), device="cuda")
    lib.pbnn_umma_tf32_gemm(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, N, 0, 0, None)
    torch.cuda.synchronize()
    assert ((D.double() - exact).norm() / exact.norm()).item() > 1e-4
