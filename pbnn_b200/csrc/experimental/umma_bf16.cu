// EXPERIMENTAL building block of the two-hidden-layer tensor-core engine (pbnn_tc2.cuh) -- not on the product path.
//
// One CTA computes D[128 x N] (fp32, TMEM) = sum over term pairs of A_ta[128 x K] * B_tb[N x K]^T with
// tcgen05.mma.kind::f16 (bf16 operands, fp32 accumulate).  What it pins down on real hardware:
//   * the SWIZZLE_128B operand image: 64 bf16 (128 B) per storage row, 16-byte chunk j of row r stored at chunk
//     j ^ (r & 7), column blocks of 64 elements `rows * 128` bytes apart;
//   * ONE image read two ways: K-major (storage row = M/N index, 64 K-elements per row) and MN-major (storage row =
//     K index, 64 M/N-elements per row) -- descriptors differ only in (start, LBO, SBO) and the idesc major bits;
//   * operands brought in with cp.async.bulk (the TMA engine, SASS UBLKCP) completing on an mbarrier;
//   * the bf16x3 split: a = a1 + a2 + a3 (each bf16), products with i + j <= 4 -> 6 passes ~ fp32 accuracy;
//     a 0/1 mask operand is exact in one term -> 3 passes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o libpbnn_umma_bf16.so umma_bf16.cu
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__device__ __forceinline__ uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(mbar), "r"(parity)
        : "memory");
}

struct UmmaBf16Args {
  const uint16_t* Aimg;  // a_terms images of a_bytes each
  const uint16_t* Bimg;
  float* D;
  int N, K;
  int a_mn, b_mn, a_terms, b_terms;
  int a_rows, b_rows;    // storage rows of one column block
  int a_bytes, b_bytes;  // bytes of one term image
};

__global__ void __launch_bounds__(160) umma_bf16_kernel(const UmmaBf16Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* As = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);  // swizzle atoms are 1024-byte aligned
  uint8_t* Bs = As + ((a.a_terms * a.a_bytes + 1023) & ~1023);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_s;

  if (warp == 4 && lane == 0) {
    // ---- producer: bulk copies global -> shared, completion counted in bytes on mbar[0] ----
    const uint32_t total = (uint32_t)(a.a_terms * a.a_bytes + a.b_terms * a.b_bytes);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar[0])), "r"(total)
                 : "memory");
    for (int which = 0; which < 2; ++which) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(which ? a.Bimg : a.Aimg);
      uint8_t* dst = which ? Bs : As;
      const int bytes = which ? a.b_terms * a.b_bytes : a.a_terms * a.a_bytes;
      for (int off = 0; off < bytes; off += 16384) {
        const int n = bytes - off < 16384 ? bytes - off : 16384;
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_u32(dst + off)),
            "l"(src + off), "r"(n), "r"(smem_u32(&mbar[0]))
            : "memory");
      }
    }
    mbar_wait(smem_u32(&mbar[0]), 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    // ---- MMA issue ----
    const uint32_t idesc = idesc_bf16(128, a.N, a.a_mn, a.b_mn);
    const int maxsum = (a.a_terms > a.b_terms ? a.a_terms : a.b_terms) + 1;
    bool first = true;
    for (int tb = 0; tb < a.b_terms; ++tb)
      for (int ta = 0; ta < a.a_terms; ++ta) {
        if (ta + tb + 2 > maxsum) continue;
        const uint32_t ab = smem_u32(As + ta * a.a_bytes), bb = smem_u32(Bs + tb * a.b_bytes);
        for (int k0 = 0; k0 < a.K; k0 += 16) {
          uint64_t da, db;
          if (!a.a_mn) da = desc_sw128(ab + (k0 >> 6) * (a.a_rows * 128) + ((k0 >> 4) & 3) * 32, 16, 1024);
          else da = desc_sw128(ab + k0 * 128, a.a_rows * 128, 1024);
          if (!a.b_mn) db = desc_sw128(bb + (k0 >> 6) * (a.b_rows * 128) + ((k0 >> 4) & 3) * 32, 16, 1024);
          else db = desc_sw128(bb + k0 * 128, a.b_rows * 128, 1024);
          const uint32_t acc = first ? 0u : 1u;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
              "l"(da), "l"(db), "r"(idesc), "r"(acc)
              : "memory");
          first = false;
        }
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(&mbar[1]))
                 : "memory");
  }
  if (warp < 4) {
    mbar_wait(smem_u32(&mbar[1]), 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < a.N; c0 += 8) {
      uint32_t r[8];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
      for (int j = 0; j < 8; ++j) a.D[(size_t)row * a.N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

// images: see the file header.  N multiple of 16 (<= 256), K multiple of 16.
extern "C" int pbnn_umma_bf16_gemm(const uint16_t* Aimg, const uint16_t* Bimg, float* D, int N, int K, int a_mn,
                                   int b_mn, int a_terms, int b_terms, void* stream) {
  if (N % 16 || N > 256 || N < 16 || K % 16 || K < 16) return -1;
  UmmaBf16Args a;
  a.Aimg = Aimg; a.Bimg = Bimg; a.D = D; a.N = N; a.K = K; a.a_mn = a_mn; a.b_mn = b_mn;
  a.a_terms = a_terms; a.b_terms = b_terms;
  const int kcb = (K + 63) / 64, ncb = (N + 63) / 64;
  a.a_rows = a_mn ? K : 128;
  a.b_rows = b_mn ? K : ((N + 7) & ~7);
  a.a_bytes = (a_mn ? 2 : kcb) * a.a_rows * 128;
  a.b_bytes = (b_mn ? ncb : kcb) * a.b_rows * 128;
  if (a.a_rows % 8 || a.b_rows % 8) return -1;
  const size_t smem = (size_t)((a_terms * a.a_bytes + 1023) & ~1023) + (size_t)b_terms * a.b_bytes + 1024;
  if (smem > 220 * 1024) return -2;
  cudaFuncSetAttribute(umma_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_bf16_kernel<<<1, 160, smem, (cudaStream_t)stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
