// EXPERIMENTAL building block for the wide-layer path -- NOT on the product path yet.
//
// One CTA computes D[128 x N] (fp32, TMEM accumulator) = A[128 x K] * B^T with tcgen05.mma.kind::tf32, operands in
// shared memory in ONE storage layout: T[x][y] -> [y/4][x][y%4] (16-byte interleave, no swizzle):
//   A  : [k/4][128 rows][4 k]                K-major  (8 rows x 16 B core matrices, SBO = 128 B, LBO = 128*16 B)
//   B  : b_mn == 0 : given as Bk[N][K]  -> [k/4][N][4 k]     K-major
//        b_mn == 1 : given as Bn[K][N]  -> [n/4][K][4 n]     MN-major (same storage a weight matrix W[in][out] has when
//                                                            it is also used K-major by the dH GEMM)
// split3 == 1 : FP32-accurate 3xTF32:  hi*hi + hi*lo + lo*hi with hi = raw fp32 word (the tensor core drops the low
//               13 mantissa bits), lo = a - trunc19(a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o libpbnn_umma_test.so umma_tf32.cu
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, SWIZZLE_NONE ("interleave"), sm_100 version bits
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version for Blackwell
  return d;                // base_offset = 0, lbo_mode = 0, layout_type = 0 (no swizzle)
}

__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// K multiple of 8, N multiple of 16, N <= 256.  Dynamic smem: (128 + N) * K * 4 * (split3 ? 2 : 1) bytes.
__global__ void __launch_bounds__(128) umma_tf32_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                        float* __restrict__ D, int K, int N, int b_mn, int split3, int dbg_swap) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* As = smem;                   // [K/4][128][4]
  float* Bs = As + 128 * K;           // K-major: [K/4][N][4]   MN-major: [N/4][K][4]
  float* Al = Bs + N * K;             // lo parts (split3)
  float* Bl = Al + 128 * K;

  // ---- stage operands (generic proxy) ----
  for (int e = tid; e < 128 * K; e += 128) {
    const int r = e / K, k = e - r * K;
    const float v = A[e];
    const int o = (k >> 2) * (128 * 4) + r * 4 + (k & 3);
    As[o] = v;
    if (split3) Al[o] = v - trunc_tf32(v);
  }
  for (int e = tid; e < N * K; e += 128) {
    float v;
    int o;
    if (!b_mn) {  // Bk[N][K]
      const int n = e / K, k = e - n * K;
      v = B[e];
      o = (k >> 2) * (N * 4) + n * 4 + (k & 3);
    } else {      // Bn[K][N]
      const int k = e / N, n = e - k * N;
      v = B[e];
      o = (n >> 2) * (K * 4) + k * 4 + (n & 3);
    }
    Bs[o] = v;
    if (split3) Bl[o] = v - trunc_tf32(v);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {  // TMEM allocation by one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;

  // ---- issue: one thread ----
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N, 0, b_mn);
    const int npass = split3 ? 3 : 1;
    bool first = true;
    for (int pass = 0; pass < npass; ++pass) {
      const float* Ap = (pass == 2) ? Al : As;  // hi*hi, hi*lo, lo*hi
      const float* Bp = (pass == 1) ? Bl : Bs;
      for (int k0 = 0; k0 < K; k0 += 8) {
        // A K-major: core matrices 128 B; along M (8 rows): SBO = 128 B; along K (next 4 k): LBO = 128 rows * 16 B
        const uint64_t da = make_desc(smem_u32(Ap + (k0 >> 2) * (128 * 4)), 128 * 16, 128);
        uint64_t db;
        if (!b_mn) {
          db = make_desc(smem_u32(Bp + (k0 >> 2) * (N * 4)), (uint32_t)N * 16, 128);
        } else {
          // MN-major: core matrix = 8 k-rows x 16 B (4 n); next 8 k: LBO = 128 B; next 4 n: SBO = K * 16 B
          db = dbg_swap ? make_desc(smem_u32(Bp + k0 * 4), (uint32_t)K * 16, 128)
                        : make_desc(smem_u32(Bp + k0 * 4), 128, (uint32_t)K * 16);
        }
        const uint32_t acc = first ? 0u : 1u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
            "l"(da), "l"(db), "r"(idesc), "r"(acc));
        first = false;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
  }
  // ---- wait for the MMAs, then TMEM -> registers -> global (warp w owns TMEM lanes 32w..32w+31 = rows) ----
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(&mbar)), "r"(0));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
    for (int j = 0; j < 8; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

extern "C" int pbnn_umma_tf32_gemm(const float* A, const float* B, float* D, int K, int N, int b_mn, int split3,
                                   void* stream) {
  const int dbg_swap = b_mn >> 1;
  b_mn &= 1;
  if (K % 8 || N % 16 || N > 256 || K < 8) return -1;
  const size_t smem = (size_t)(128 + N) * K * 4 * (split3 ? 2 : 1);
  if (smem > 220 * 1024) return -2;
  cudaFuncSetAttribute(umma_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_tf32_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, K, N, b_mn, split3, dbg_swap);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
