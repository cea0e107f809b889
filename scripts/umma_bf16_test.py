"""GPU check of the experimental bf16 tcgen05 building block (pbnn_b200/csrc/experimental/umma_bf16.cu):
SWIZZLE_128B images read K-major and MN-major, cp.async.bulk staging, bf16x3 split accuracy."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "pbnn_b200", "csrc", "experimental", "libpbnn_umma_bf16.so"))
lib.pbnn_umma_bf16_gemm.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 6 + [ctypes.c_void_p]


def split(x, terms):
    out, r = [], x.clone()
    for _ in range(terms):
        t = r.bfloat16().float()
        out.append(t)
        r = r - t
    return out


def image(S):
    """S [R][C] float (bf16-representable) -> SWIZZLE_128B image: [C/64 blocks][R rows][8 chunks ^ (r & 7)][8]."""
    R, C = S.shape
    ncb = (C + 63) // 64
    P = torch.zeros(R, ncb * 64, device=S.device)
    P[:, :C] = S
    bits = P.bfloat16().view(torch.int16).view(R, ncb, 8, 8).permute(1, 0, 2, 3).contiguous()  # [cb][r][chunk][e]
    r = torch.arange(R, device=S.device)
    j = torch.arange(8, device=S.device)
    phys = (j[None, :] ^ (r[:, None] & 7))  # [R][8]: logical chunk j of row r lives at phys[r][j]
    out = torch.zeros_like(bits)
    out.scatter_(2, phys[None, :, :, None].expand(ncb, R, 8, 8), bits)
    return out.contiguous()


def run(N, K, a_mn, b_mn, a_terms, b_terms, mask=False, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(128, K, device="cuda", generator=g)
    if mask:
        A = (A > 0).float()
    Bk = torch.randn(N, K, device="cuda", generator=g)
    As, Bs = split(A, a_terms), split(Bk, b_terms)
    Aimg = torch.stack([image(t.t().contiguous() if a_mn else t) for t in As]).contiguous()
    Bimg = torch.stack([image(t.t().contiguous() if b_mn else t) for t in Bs]).contiguous()
    D = torch.full((128, N), float("nan"), device="cuda")
    rc = lib.pbnn_umma_bf16_gemm(Aimg.data_ptr(), Bimg.data_ptr(), D.data_ptr(), N, K, a_mn, b_mn, a_terms, b_terms, None)
    torch.cuda.synchronize()
    exact = A.double() @ Bk.double().t()
    maxsum = max(a_terms, b_terms) + 1
    ref = sum(As[i].double() @ Bs[j].double().t() for i in range(a_terms) for j in range(b_terms) if i + j + 2 <= maxsum)
    e_exact = ((D.double() - exact).norm() / exact.norm()).item()
    e_ref = ((D.double() - ref).norm() / ref.norm()).item()
    print(f"N={N} K={K} a_mn={a_mn} b_mn={b_mn} terms={a_terms}/{b_terms} mask={mask} rc={rc}: "
          f"vs exact {e_exact:.3e}, vs term-sum {e_ref:.3e}, nan={bool(torch.isnan(D).any())}", flush=True)
    return e_exact, e_ref


if __name__ == "__main__":
    for (a_mn, b_mn) in ((0, 0), (1, 0), (0, 1), (1, 1)):
        run(64, 64, a_mn, b_mn, 1, 1)
        run(256, 128, a_mn, b_mn, 1, 1, seed=1)
        run(112, 48, a_mn, b_mn, 1, 1, seed=2)
    run(128, 64, 0, 0, 3, 3)
    run(128, 64, 1, 1, 3, 3)
    run(112, 64, 0, 1, 3, 3)
    run(128, 128, 1, 1, 1, 3, mask=True)
    run(256, 64, 0, 0, 1, 3, mask=True)
