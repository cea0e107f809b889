"""GPU check of the experimental tcgen05 TF32 GEMM building block (pbnn_b200/csrc/experimental/umma_tf32.cu)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "pbnn_b200", "csrc", "experimental", "libpbnn_umma_test.so"))
lib.pbnn_umma_tf32_gemm.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 4 + [ctypes.c_void_p]

def trunc(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)

def run(K, N, b_mn, split3, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(128, K, device="cuda", generator=g)
    Bk = torch.randn(N, K, device="cuda", generator=g)        # logical B^T: D = A @ Bk^T
    Bin = Bk.t().contiguous() if (b_mn & 1) else Bk            # MN-major input is [K][N]
    D = torch.full((128, N), float("nan"), device="cuda")
    rc = lib.pbnn_umma_tf32_gemm(A.data_ptr(), Bin.data_ptr(), D.data_ptr(), K, N, b_mn, split3, None)
    torch.cuda.synchronize()
    exact = A.double() @ Bk.double().t()
    tf = trunc(A).double() @ trunc(Bk).double().t()
    e_exact = ((D.double() - exact).norm() / exact.norm()).item()
    e_tf = ((D.double() - tf).norm() / tf.norm()).item()
    print("D[0,:4]", D[0, :4].tolist(), "ref", tf[0, :4].tolist())
    print(f"K={K} N={N} b_mn={b_mn} split3={split3} rc={rc}: rel err vs exact {e_exact:.3e}, vs tf32-truncated {e_tf:.3e}, nan={bool(torch.isnan(D).any())}")
    return e_exact, e_tf

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("basic", "all"):
    run(8, 16, 0, 0); run(32, 64, 0, 0); run(64, 256, 0, 0)
if which in ("mn", "all"):
    run(8, 16, 1, 0); run(32, 64, 1, 0); run(64, 256, 1, 0)
if which == "mnswap":
    run(8, 16, 3, 0); run(16, 16, 3, 0); run(32, 64, 3, 0); run(16, 16, 1, 0)
if which in ("split", "all"):
    run(64, 256, 0, 1); run(64, 256, 1, 1); run(40, 128, 1, 1, seed=3)
