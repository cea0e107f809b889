// pbnn_b200: CUDA (sm_100a) backend + C ABI.  Build: see __graft_entry__.build().
//
// Kernels
//   pb_sg_kernel<ALGO,ACT>   fused chain-step kernel: one CTA per chain / ensemble member, persistent over
//                            all T steps (forward, backward, drift, threefry noise, trace) -- SGLD, pSGLD,
//                            SGHMC, Adam, and the grad_estimator test hook
//   pb_hmc_kernel<ACT>       full-batch leapfrog + Metropolis accept, one CTA per chain (FP32 engines)
//   pb_hmc_tc_kernel<ENG>    the same chain on the tcgen05 engine (pbnn_tc.cuh): split-TF32 GEMMs with TMEM
//                            accumulators, two small CTAs per SM, balanced persistent schedule for C > #SM
//   pb_predict_kernel<ACT>   posterior predictive over (test-row tile) x (sample chunk)
//   pb_rng_kernel / pb_mbidx_kernel / pb_perm_kernel   jax.random primitives
#include <cuda_runtime.h>
#include <stdio.h>

#include <atomic>
#include <map>
#include <mutex>
#include <utility>

#include "pbnn_engine.cuh"
#include "pbnn_warp.cuh"
#include "pbnn_warp2.cuh"
#include "pbnn_tc2_pred.cuh"

#if defined(PB_PROFILE)
#define PB_PROF_START                                                  \
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {        \
    pb_prof_last = clock64();                                          \
    pb_prof_tag = 0;                                                   \
  }
#else
#define PB_PROF_START
#endif

// ------------------------------------------------------------------------------------------------
template <int ALGO, int ACT, bool GS, int ENG>
__global__ void __launch_bounds__(PB_MAX_NT)
pb_sg_kernel(const __grid_constant__ PbPlan pl, const __grid_constant__ PbSgArgs a) {
  extern __shared__ __align__(16) float pb_smem[];
  PB_PROF_START
  pb_sg_chain<ALGO, ACT, GS, ENG>(pl, a, pb_smem, (int)blockIdx.x, (int)blockDim.x);
}

// two-hidden-layer tcgen05 engine (pbnn_tc2.cuh): 8 worker warps + TMA producer warp + MMA-issue warp, one persistent
// CTA per SM looping over its chains; state arrays in the global workspace, operand images in the CTA's image block
template <int ALGO>
__global__ void __launch_bounds__(PB_T2_NT, 1)
pb_sg_tc2_kernel(const __grid_constant__ PbPlan pl, const __grid_constant__ PbSgArgs a) {
  extern __shared__ __align__(16) float pb_smem[];
  PB_PROF_START
  const int nt = (int)blockDim.x;
  PbTc2 tc;
  pb_tc2_setup(pl, pb_smem, a.ws_img + (size_t)blockIdx.x * (size_t)pl.t2_img_bytes, nt, tc);
  for (int c = (int)blockIdx.x; c < a.n_chains; c += (int)gridDim.x) {
    pb_sg_chain<ALGO, PBNN_ACT_RELU, true, 2>(pl, a, pb_smem, c, nt, &tc);
    __syncthreads();
  }
  pb_tc2_teardown(tc);
}

template <int ACT, bool GS, int ENG>
__global__ void __launch_bounds__(PB_MAX_NT)
pb_hmc_kernel(const __grid_constant__ PbPlan pl, const __grid_constant__ PbHmcArgs a) {
  extern __shared__ __align__(16) float pb_smem[];
  PB_PROF_START
  pb_hmc_chain<ACT, GS, ENG>(pl, a, pb_smem, (int)blockIdx.x, (int)blockDim.x);
}

// tcgen05 engine: 4 epilogue warps + 1 MMA-issue warp, two CTAs per SM (~93 KB of shared memory, 256 TMEM columns).
// a.sched_done == NULL: CTA c advances chain c.  Else (more chains than SMs) the CTAs are persistent and balanced:
// item n = segment * C + chain; a CTA takes the next item from a global counter and starts it when the CTA that ran
// the chain's previous segment has published its state (release store / acquire load of sched_done[chain]).  Items are
// handed out in increasing order, so a predecessor (item n - C) is always owned by a CTA that is already running:
// no co-residency assumption, no deadlock.
template <int ENG>  // 2: N <= 512 (X resident in TMEM); 3: larger N (X re-staged per group of 4 tiles)
__global__ void __launch_bounds__(160, 2)
pb_hmc_tc_kernel(const __grid_constant__ PbPlan pl, const __grid_constant__ PbHmcArgs a) {
  extern __shared__ __align__(16) float pb_smem[];
  PB_PROF_START
  const int nt = (int)blockDim.x;
  pb_init_acts(pl, pb_smem, nt);
  PbTc tc;
  pb_tc_setup(pl, pb_smem, a.X, a.y, a.N, nt, tc);
  if (!a.sched_done) {
    pb_hmc_chain<PBNN_ACT_RELU, false, ENG>(pl, a, pb_smem, (int)blockIdx.x, nt, 0, -1, false, &tc);
  } else {
    const int nseg = (a.S + a.seg_len - 1) / a.seg_len;
    const int nitems = nseg * a.n_chains;
    int* next = a.sched_done + a.n_chains;
    volatile int* slot = reinterpret_cast<volatile int*>(pb_scu(pl, pb_smem) + 24);
    for (;;) {
      if (threadIdx.x == 0) *slot = atomicAdd(next, 1);
      __syncthreads();
      const int item = *slot;
      __syncthreads();
      if (item >= nitems) break;
      const int seg = item / a.n_chains, c = item - seg * a.n_chains;
      const int s0 = seg * a.seg_len, s1 = s0 + a.seg_len < a.S ? s0 + a.seg_len : a.S;
      volatile int* ok = slot + 1;
      if (threadIdx.x == 0) {
        int done = seg;  // segment 0 has no predecessor
        if (seg > 0) {
          // bounded in TIME (30 s): a scheduling bug must not hang the GPU.  On time-out the segment is SKIPPED (the
          // chain keeps the last published state), bit 1 of the chain's flag word is set and carried to the end --
          // the host turns it into an error -- and the hand-over is still published so later segments do not stall.
          unsigned long long t_start, t_now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
          for (;;) {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(done) : "l"(a.sched_done + c) : "memory");
            if (done >= seg) break;
            __nanosleep(200);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
            if (t_now - t_start > 30000000000ull) break;
          }
        }
        *ok = done >= seg ? 1 : 0;
      }
      __syncthreads();
      const bool run = *ok != 0;
      if (run) pb_hmc_chain<PBNN_ACT_RELU, false, ENG>(pl, a, pb_smem, c, nt, s0, s1, seg > 0, &tc);
      __syncthreads();
      if (threadIdx.x == 0) {
        if (!run) {
          atomicOr(a.sched_flags + c, 2u);
          if (s1 >= a.S && a.flags) a.flags[c] = a.sched_flags[c];
        }
        if (s1 < a.S) {
          __threadfence();
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.sched_done + c), "r"(seg + 1) : "memory");
        }
      }
    }
  }
  pb_tc_teardown(pl, pb_smem, tc);
}

// ------------------------------------------------------------------------------------------------
// Posterior predictive on the tcgen05 engine (one-hidden-layer relu nets, scalar output): CTA (g, ch) evaluates the
// samples of chunk ch on `tpc` 128-row tiles of test rows.  X_test (hi, lo) of the CTA's tiles is resident in TMEM
// (forward A operand, as in pb_hmc_tc_kernel); per sample the first-layer weights are staged as the B operand (hi, lo)
// and 6 MMAs per tile put Hpre into one of two TMEM accumulators; one thread per row forms f = w2 . relu(Hpre) + b2,
// writes it and accumulates sum f, sum f^2 in registers (predict_fn, langevin.py:75-76; moments:
// benchmark/pipelines/metrics_prediction_intervals.py:188).
// ------------------------------------------------------------------------------------------------
struct PbPredTcArgs {
  const float* samples;
  int S;
  const float* Xtest;
  int M;
  float* preds;
  float* part;
  int nchunk, chunk, tpc;
  int d, H, P, fb0, fw0, fb1, fw1;
};

__global__ void __launch_bounds__(160, 2) pb_predict_tc_kernel(const __grid_constant__ PbPredTcArgs a) {
  __shared__ __align__(16) float Wb[2 * 16 * 64];  // forward B operand: raw | lo
  __shared__ __align__(16) float w2s[64];
  __shared__ __align__(8) uint64_t mbar[4];        // 0,1: forward MMAs into accumulator 0/1 done; 2,3: accumulator read
  __shared__ uint32_t tmem_s;
  __shared__ float b2s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int row0 = (int)blockIdx.x * a.tpc * 128, ch = (int)blockIdx.y;
  const int left = (a.M - row0 + 127) >> 7;
  const int ntl = left < a.tpc ? left : a.tpc;
  const int d = a.d, H = a.H, nj8 = (H + 7) >> 3;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pb_smem_u32(mbar + i)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(pb_smem_u32(mbar + 2 + i)));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(pb_smem_u32(&tmem_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  pb_tc_fence_before();
  __syncthreads();
  pb_tc_fence_after();
  const uint32_t tmem = tmem_s;
  if (tid < 128) {  // X_test tiles -> TMEM columns 128 + 32 t (raw | lo), lane = row of the tile
    for (int t = 0; t < ntl; ++t) {
      const int row = row0 + t * 128 + tid;
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float v = 0.f;
        if (row < a.M) v = k < d ? a.Xtest[(size_t)row * d + k] : (k == d ? 1.f : 0.f);
        hi[k] = __float_as_uint(v);
        lo[k] = __float_as_uint(v - pb_trunc_tf32(v));
      }
      const uint32_t taddr = tmem + ((uint32_t)(tid & ~31) << 16) + 128u + 32u * (uint32_t)t;
      pb_tmem_st16(taddr, hi);
      pb_tmem_st16(taddr + 16u, lo);
    }
    pb_tmem_wait_st();
  }
  pb_tc_fence_before();
  __syncthreads();
  pb_tc_fence_after();

  const uint32_t idesc = pb_umma_idesc_tf32(128, 64);
  const uint64_t dwb = pb_umma_desc(pb_smem_u32(Wb), 1024u, 128u);
  const bool k2 = d + 1 > 8;
  uint32_t pf0 = 0u, pf1 = 0u, ph0 = 0u, ph1 = 0u;  // mbarrier parities (epilogue: pf, issue warp: ph)
  float sum[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
  const int s0 = ch * a.chunk, s1 = s0 + a.chunk < a.S ? s0 + a.chunk : a.S;
  // this thread's 7 first-layer slots (constant over the samples) and a register prefetch of the next sample's
  // parameters: the global-memory latency of a sample is hidden behind the MMAs + epilogue of the previous one
  int wsrc[7], wdst[7];
#pragma unroll
  for (int u = 0; u < 7; ++u) {
    const int e = tid + u * 160;
    const int i = (e & 3) + 4 * (e >> 8), j = (e >> 2) & 63;
    wdst[u] = e < 16 * 64 ? (i >> 2) * 256 + j * 4 + (i & 3) : -1;
    wsrc[u] = (e < 16 * 64 && j < H) ? (i < d ? a.fw0 + i * H + j : (i == d ? a.fb0 + j : -1)) : -1;
  }
  const int w2src = tid < 64 ? (tid < H ? a.fw1 + tid : -1) : (tid == 64 ? a.fb1 : -2);
  float wv[7], w2v = 0.f;
  auto fetch = [&](int smp) {
    const float* th = a.samples + (size_t)smp * a.P;
#pragma unroll
    for (int u = 0; u < 7; ++u) wv[u] = wsrc[u] >= 0 ? __ldg(th + wsrc[u]) : 0.f;
    w2v = w2src >= 0 ? __ldg(th + w2src) : 0.f;
  };
  if (s0 < s1) fetch(s0);
  for (int smp = s0; smp < s1; ++smp) {
    // first-layer weights -> B operand (raw + lo); kernel [d][H] at fw0, bias [H] at fb0 (ravel_pytree order)
#pragma unroll
    for (int u = 0; u < 7; ++u)
      if (wdst[u] >= 0) {
        Wb[wdst[u]] = wv[u];
        Wb[1024 + wdst[u]] = wv[u] - pb_trunc_tf32(wv[u]);
      }
    if (tid < 64) w2s[tid] = w2v;
    if (tid == 64) b2s = w2v;
    if (smp + 1 < s1) fetch(smp + 1);
    pb_fence_async_smem();
    pb_tc_fence_before();
    __syncthreads();
    if (warp_u == 4) {
      pb_tc_fence_after();
      for (int t = 0; t < ntl; ++t) {
        const int hb = t & 1;
        if (t >= 2) {  // the accumulator's previous tile has been read
          if (hb) { pb_mbar_wait(pb_smem_u32(mbar + 3), ph1); ph1 ^= 1u; }
          else { pb_mbar_wait(pb_smem_u32(mbar + 2), ph0); ph0 ^= 1u; }
          pb_tc_fence_after();
        }
        if (pb_elect_one()) {
          const uint32_t dcol = tmem + 64u * (uint32_t)hb, ta = tmem + 128u + 32u * (uint32_t)t;
          const uint64_t w_lo = 256u, w_k = 128u;
          pb_umma_tf32_ts(dcol, ta, dwb, idesc, 0u);
          if (k2) pb_umma_tf32_ts(dcol, ta + 8u, dwb + w_k, idesc, 1u);
          pb_umma_tf32_ts(dcol, ta, dwb + w_lo, idesc, 1u);
          if (k2) pb_umma_tf32_ts(dcol, ta + 8u, dwb + w_lo + w_k, idesc, 1u);
          pb_umma_tf32_ts(dcol, ta + 16u, dwb, idesc, 1u);
          if (k2) pb_umma_tf32_ts(dcol, ta + 24u, dwb + w_k, idesc, 1u);
          pb_umma_commit(pb_smem_u32(mbar + hb));
        }
        __syncwarp();
      }
      // every "accumulator read" phase is waited for exactly once (the parities stay in step)
      for (int t = ntl >= 2 ? ntl - 2 : 0; t < ntl; ++t) {
        if (t & 1) { pb_mbar_wait(pb_smem_u32(mbar + 3), ph1); ph1 ^= 1u; }
        else { pb_mbar_wait(pb_smem_u32(mbar + 2), ph0); ph0 ^= 1u; }
      }
    } else {
      const float b2 = b2s;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (t < ntl) {
          const int hb = t & 1;
          if (hb) { pb_mbar_wait(pb_smem_u32(mbar + 1), pf1); pf1 ^= 1u; }
          else { pb_mbar_wait(pb_smem_u32(mbar + 0), pf0); pf0 ^= 1u; }
          pb_tc_fence_after();
          uint32_t h[64];
          const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16) + 64u * (uint32_t)hb;
          pb_tmem_ld32(lane_base, h);
          pb_tmem_ld32(lane_base + 32u, h + 32);
          pb_tmem_wait_ld();
          pb_tc_fence_before();
          asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(
                           pb_smem_u32(mbar + 2 + hb))
                       : "memory");
          float f0 = 0.f, f1 = 0.f;
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8)
            if (c8 < nj8) {
#pragma unroll
              for (int j4 = 2 * c8; j4 < 2 * c8 + 2; ++j4) {
                const float4 w = pb_ld4(w2s + 4 * j4);
                f0 = fmaf(fmaxf(__uint_as_float(h[4 * j4]), 0.f), w.x, f0);
                f1 = fmaf(fmaxf(__uint_as_float(h[4 * j4 + 1]), 0.f), w.y, f1);
                f0 = fmaf(fmaxf(__uint_as_float(h[4 * j4 + 2]), 0.f), w.z, f0);
                f1 = fmaf(fmaxf(__uint_as_float(h[4 * j4 + 3]), 0.f), w.w, f1);
              }
            }
          const float f = (f0 + f1) + b2;
          const int row = row0 + t * 128 + tid;
          if (row < a.M) {
            sum[t] += f;
            sq[t] += f * f;
            if (a.preds) a.preds[(size_t)smp * a.M + row] = f;
          }
        }
    }
    pb_tc_fence_before();
    __syncthreads();  // every MMA of this sample has completed (each tile was waited for): Wb may be rewritten
  }
  if (a.part && tid < 128) {
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (t < ntl) {
        const int row = row0 + t * 128 + tid;
        if (row < a.M) {
          a.part[((size_t)ch * 2) * a.M + row] = sum[t];
          a.part[((size_t)ch * 2 + 1) * a.M + row] = sq[t];
        }
      }
  }
  pb_tc_fence_before();
  __syncthreads();
  if (tid < 32) {
    pb_tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
  }
}

template <int ACT, bool GS>
__global__ void __launch_bounds__(PB_MAX_NT)
pb_predict_kernel(const __grid_constant__ PbPlan pl, const __grid_constant__ PbPredArgs a) {
  extern __shared__ __align__(16) float pb_smem[];
  pb_predict_block<ACT, GS>(pl, a, pb_smem, (int)blockIdx.x, (int)blockIdx.y, (int)blockDim.x);
}

__global__ void pb_predict_reduce_kernel(const float* __restrict__ part, int nchunk, int n, float* __restrict__ sum,
                                         float* __restrict__ sumsq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = 0.f, q = 0.f;
  for (int ch = 0; ch < nchunk; ++ch) {
    a += part[(size_t)ch * 2 * n + i];
    q += part[(size_t)ch * 2 * n + n + i];
  }
  sum[i] = a;
  sumsq[i] = q;
}

// kind: 0 split, 1 bits, 2 uniform, 3 normal, 4 fold_in (n carries the 32-bit datum)
__global__ void pb_rng_kernel(int kind, const uint32_t* __restrict__ keys, int C, int n, void* __restrict__ out) {
  if (kind == 4) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
      PbKey k;
      k.k0 = keys[2 * c];
      k.k1 = keys[2 * c + 1];
      const PbKey o = pb_split_at(k, (uint32_t)n);
      reinterpret_cast<uint32_t*>(out)[2 * c] = o.k0;
      reinterpret_cast<uint32_t*>(out)[2 * c + 1] = o.k1;
    }
    return;
  }
  const size_t total = (size_t)C * n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e / n);
    const uint32_t i = (uint32_t)(e - (size_t)c * n);
    PbKey k;
    k.k0 = keys[2 * c];
    k.k1 = keys[2 * c + 1];
    if (kind == 0) {
      const PbKey o = pb_split_at(k, i);
      reinterpret_cast<uint32_t*>(out)[2 * e] = o.k0;
      reinterpret_cast<uint32_t*>(out)[2 * e + 1] = o.k1;
    } else if (kind == 1) {
      reinterpret_cast<uint32_t*>(out)[e] = pb_bits_at(k, i);
    } else if (kind == 2) {
      reinterpret_cast<float*>(out)[e] = fmaxf(0.f, pb_bits_to_uniform(pb_bits_at(k, i)));
    } else {
      reinterpret_cast<float*>(out)[e] = pb_normal_at(k, i);
    }
  }
}

__global__ void pb_mbidx_kernel(const uint32_t* __restrict__ keys, int draw, int B, int N, uint32_t mult,
                                int32_t* __restrict__ out) {
  const int c = blockIdx.x;
  PbKey k;
  k.k0 = keys[2 * c];
  k.k1 = keys[2 * c + 1];
  for (int i = 0; i < draw; ++i) k = pb_split_at(k, 1u);
  const PbKey k1 = pb_split_at(k, 0u), k2 = pb_split_at(k, 1u);
  for (int b = threadIdx.x; b < B; b += blockDim.x)
    out[(size_t)c * B + b] = pb_randint_at(k1, k2, (uint32_t)b, (uint32_t)N, mult);
}

// jax.random.permutation (jax/_src/random.py::_shuffle): `rounds` x { key, sub = split(key); stable sort by
// bits(sub) }.  The stable sort is a bitonic sort of the composite (bits << 32 | position), which is a strict
// total order, so any sorting network reproduces the stable result.  One CTA per key.
__global__ void __launch_bounds__(1024, 1)
pb_perm_kernel(const uint32_t* __restrict__ keys, int N, int Np2, int rounds, int use_smem, int32_t* __restrict__ out,
               unsigned long long* __restrict__ ws) {
  extern __shared__ __align__(16) unsigned long long pb_sort[];
  const int c = blockIdx.x;
  unsigned long long* wsc = ws + (size_t)c * 3 * N;
  unsigned long long* buf = use_smem ? pb_sort : wsc;
  int32_t* xa = reinterpret_cast<int32_t*>(wsc + 2 * (size_t)N);
  int32_t* xb = xa + N;
  int32_t* res = out + (size_t)c * N;
  PbKey key;
  key.k0 = keys[2 * c];
  key.k1 = keys[2 * c + 1];
  const int32_t* xold = nullptr;  // nullptr == identity
  for (int r = 0; r < rounds; ++r) {
    const PbKey sub = pb_split_at(key, 1u);
    key = pb_split_at(key, 0u);
    for (int i = threadIdx.x; i < Np2; i += blockDim.x)
      buf[i] = i < N ? (((unsigned long long)pb_bits_at(sub, (uint32_t)i) << 32) | (unsigned)i) : ~0ull;
    __syncthreads();
    for (int k = 2; k <= Np2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < Np2; i += blockDim.x) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const unsigned long long a = buf[i], b = buf[ixj];
            const bool up = (i & k) == 0;
            if ((a > b) == up) {
              buf[i] = b;
              buf[ixj] = a;
            }
          }
        }
        __syncthreads();
      }
    int32_t* xnew = (r == rounds - 1) ? res : ((r & 1) ? xb : xa);
    for (int jn = threadIdx.x; jn < N; jn += blockDim.x) {
      const int src = (int)(buf[jn] & 0xFFFFFFFFull);
      xnew[jn] = xold ? xold[src] : src;
    }
    __syncthreads();
    xold = xnew;
  }
  if (rounds == 0)
    for (int jn = threadIdx.x; jn < N; jn += blockDim.x) res[jn] = jn;
}

// ------------------------------------------------------------------------------------------------
// thinning_fn (src/pbnn/utils/misc.py:58-125): greedy energy-distance thinning of a chain.
//   k(x, y) = |x| + |y| - sqrt(max(x.x + y.y - 2 x.y, 0));  k0_mean[i] = mean_j k(x_i, x_j)
//   obj = 2|x_i| - 2 k0_mean;  idx_0 = argmin obj;  then size-1 times: obj += 2 k(x_idx, .) - 2 k0_mean, argmin.
// ------------------------------------------------------------------------------------------------
__global__ void pb_thin_norm_kernel(const float* __restrict__ X, int n, int D, float* __restrict__ xx) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  float a = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float v = X[(size_t)row * D + k];
    a += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) xx[row] = a;
}

// row sums of the kernel matrix without materialising it: CTA = 64 rows, loops over all column tiles (deterministic
// order), 4x4 register tiles, operands staged K-major in shared memory
__global__ void __launch_bounds__(256) pb_thin_rowsum_kernel(const float* __restrict__ X, const float* __restrict__ xx,
                                                             int n, int D, float* __restrict__ k0sum) {
  __shared__ __align__(16) float As[16][68], Bs[16][68];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.x * 64;
  float rowsum[4] = {0.f, 0.f, 0.f, 0.f};
  float ni[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) ni[r] = (i0 + ty * 4 + r < n) ? xx[i0 + ty * 4 + r] : 0.f;
  for (int j0 = 0; j0 < n; j0 += 64) {
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    for (int k0 = 0; k0 < D; k0 += 16) {
      for (int e = tid; e < 64 * 16; e += 256) {
        const int r = e >> 4, k = e & 15;
        As[k][r] = (i0 + r < n && k0 + k < D) ? X[(size_t)(i0 + r) * D + k0 + k] : 0.f;
        Bs[k][r] = (j0 + r < n && k0 + k < D) ? X[(size_t)(j0 + r) * D + k0 + k] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] += av[r] * bv[c];
      }
      __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + tx * 4 + c;
        if (j < n) {
          const float nj = xx[j];
          rowsum[r] += sqrtf(ni[r]) + sqrtf(nj) - sqrtf(fmaxf(ni[r] + nj - 2.0f * acc[r][c], 0.f));
        }
      }
  }
  // reduce the 16 column-threads of each row (lanes 0..15 / 16..31 of a warp share ty)
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float v = rowsum[r];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (tx == 0 && i0 + ty * 4 + r < n) k0sum[i0 + ty * 4 + r] = v;
  }
}

// greedy selection, one step = two launches (size-1 dependent argmins: the dependency is carried through `cur`):
//   update: every warp of the grid takes rows i, i+W, ...: obj[i] += 2 k(x_cur, x_i) - 2 k0_mean[i] (step > 0) and keeps the
//           CTA's (min, first index); argmin: one CTA reduces the per-CTA partials, records the pick.
__global__ void __launch_bounds__(256) pb_thin_update_kernel(const float* __restrict__ X, const float* __restrict__ xx,
                                                             const float* __restrict__ k0sum, int n, int D, int step,
                                                             const int32_t* __restrict__ cur, float* __restrict__ obj,
                                                             float* __restrict__ pval, int32_t* __restrict__ pidx) {
  __shared__ float sval[8];
  __shared__ int sidx[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * 8 + warp, nwarps = gridDim.x * 8;
  const float inv_n = 1.0f / (float)n;
  float best = INFINITY;
  int bi = 0x7fffffff;
  const int c = step > 0 ? cur[0] : 0;
  const float nc = xx[c];
  const float* xc = X + (size_t)c * D;
  for (int i = gw; i < n; i += nwarps) {
    float v;
    if (step == 0) {
      v = 2.0f * sqrtf(xx[i]) - 2.0f * (k0sum[i] * inv_n);
    } else {
      const float* xi = X + (size_t)i * D;
      float a = 0.f;
      for (int k = lane; k < D; k += 32) a += xc[k] * xi[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      const float ki = sqrtf(nc) + sqrtf(xx[i]) - sqrtf(fmaxf(nc + xx[i] - 2.0f * a, 0.f));
      v = obj[i] + 2.0f * ki - 2.0f * (k0sum[i] * inv_n);
    }
    if (lane == 0) obj[i] = v;
    if (v < best) {  // rows ascend within a warp: strict < keeps the first index on ties (jnp.argmin)
      best = v;
      bi = i;
    }
  }
  if (lane == 0) {
    sval[warp] = best;
    sidx[warp] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (sval[w] < best || (sval[w] == best && sidx[w] < bi)) {
        best = sval[w];
        bi = sidx[w];
      }
    pval[blockIdx.x] = best;
    pidx[blockIdx.x] = bi;
  }
}

__global__ void __launch_bounds__(256) pb_thin_argmin_kernel(const float* __restrict__ pval,
                                                             const int32_t* __restrict__ pidx, int nblocks, int step,
                                                             int size, int32_t* __restrict__ cur,
                                                             int32_t* __restrict__ out) {
  __shared__ float sval[256];
  __shared__ int sidx[256];
  float best = INFINITY;
  int bi = 0x7fffffff;
  for (int b = threadIdx.x; b < nblocks; b += 256)
    if (pval[b] < best || (pval[b] == best && pidx[b] < bi)) {
      best = pval[b];
      bi = pidx[b];
    }
  sval[threadIdx.x] = best;
  sidx[threadIdx.x] = bi;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int t = 1; t < 256; ++t)
      if (sval[t] < best || (sval[t] == best && sidx[t] < bi)) {
        best = sval[t];
        bi = sidx[t];
      }
    cur[0] = bi;
    // reference order: [picks of steps 1..size-1], then the initial pick appended last (misc.py:121-123)
    out[step == 0 ? size - 1 : step - 1] = bi;
  }
}

// FP32 FMA throughput probe (roofline denominator measured in-run): 16 independent accumulator pairs per thread,
// `iters` x 32 FMAs each, as scalar FFMA (mode 0) or packed FFMA2 / fma.rn.f32x2 (mode 1).
__global__ void __launch_bounds__(256, 4) pb_fma_probe_kernel(int mode, int iters, float seed, float* sink) {
  float2 a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
  const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-7f, -1e-7f);
  if (mode == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        a[i].x = fmaf(a[i].x, m.x, c.x);
        a[i].y = fmaf(a[i].y, m.y, c.y);
      }
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = __ffma2_rn(a[i], m, c);
    }
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += a[i].x + a[i].y;
  if (r == 123.456f) sink[0] = r;
}

// ------------------------------------------------------------------------------------------------
// backends
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
#define PB_ERRBUF_DEFINED 1

#define PB_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      snprintf(g_err, sizeof(g_err), "%s failed: %s", #call, cudaGetErrorString(e_)); \
      return PBNN_ECUDA;                                                                      \
    }                                                                                         \
  } while (0)

// SM count of the current device (148 on a B200; also the answer when no device is visible: host-side queries only)
// (per-device caches of immutable device / kernel properties -- not behavioural state)
static int pb_backend_num_sms() {
  static std::atomic<int> cached[64];
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return 148;
  }
  if (dev >= 0 && dev < 64 && cached[dev].load(std::memory_order_relaxed) > 0) return cached[dev].load();
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) {
    (void)cudaGetLastError();
    return 148;
  }
  if (dev >= 0 && dev < 64) cached[dev].store(n);
  return n;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is raised once per (kernel, device), not on every launch
static int pb_ensure_smem(const void* kern, size_t smem) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> have;
  int dev = 0;
  PB_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  size_t& cur = have[std::make_pair(kern, dev)];
  if (smem > cur) {
    PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cur = smem;
  }
  return PBNN_OK;
}

template <class K, class A>
static int pb_launch(K kern, const PbPlan& pl, const A& a, dim3 grid, void* stream) {
  const size_t smem = (size_t)pl.smem_floats * 4;
  const int rc = pb_ensure_smem(reinterpret_cast<const void*>(kern), smem);
  if (rc) return rc;
  kern<<<grid, pl.NT, smem, (cudaStream_t)stream>>>(pl, a);
  PB_CUDA(cudaGetLastError());
  return PBNN_OK;
}

template <int ALGO>
static int pb_sg_dispatch(const PbPlan& pl, const PbSgArgs& a, int C, void* stream) {
  if (pl.tc2) {
    int nsm = pb_backend_num_sms();
    const int cap = pb_opt(PB_OPT_TC2_CTAS, 0);  // experiments: fewer persistent CTAs (smaller L2 working set)
    if (cap > 0 && cap < nsm) nsm = cap;
    return pb_launch(pb_sg_tc2_kernel<ALGO>, pl, a, dim3(C < nsm ? C : nsm), stream);
  }
  if constexpr (ALGO == PB_ALGO_ADAM) {
    if (pl.w2) {  // Adam for small two-hidden-layer relu nets: one four-warp CTA per member
#define PB_W2_LAUNCH(DP_) return pb_launch(pb_adam_w2_kernel<DP_>, pl, a, dim3(C), stream);
      switch (pl.w2_dp) {
        case 1: PB_W2_LAUNCH(1)
        case 2: PB_W2_LAUNCH(2)
        case 3: PB_W2_LAUNCH(3)
        default: PB_W2_LAUNCH(4)
      }
#undef PB_W2_LAUNCH
    }
  }
  if constexpr (ALGO == PB_ALGO_SGLD) {
    if (pl.w1 == 2) {  // few chains: a CTA of four warps per chain (pbnn_warp.cuh, NWC = 4)
      switch (pl.h1_dp) {
        case 1: return pb_launch(pb_sgc_kernel<1>, pl, a, dim3(C), stream);
        case 2: return pb_launch(pb_sgc_kernel<2>, pl, a, dim3(C), stream);
        case 3: return pb_launch(pb_sgc_kernel<3>, pl, a, dim3(C), stream);
        default: return pb_launch(pb_sgc_kernel<4>, pl, a, dim3(C), stream);
      }
    }
  }
  if constexpr (ALGO != PB_ALGO_GRAD) {
    if (pl.w1 == 1) {  // warp-per-chain engine: PB_W1_WPC chains per CTA, DP = float4 chunks of an augmented input row
      const dim3 g((C + PB_W1_WPC - 1) / PB_W1_WPC);
#define PB_W1_LAUNCH(DP_)                                                                                         \
  return pl.act == PBNN_ACT_RELU ? pb_launch(pb_sgw_kernel<ALGO, PBNN_ACT_RELU, DP_>, pl, a, g, stream)           \
                                 : pb_launch(pb_sgw_kernel<ALGO, PBNN_ACT_TANH, DP_>, pl, a, g, stream);
      switch (pl.h1_dp) {
        case 1: PB_W1_LAUNCH(1)
        case 2: PB_W1_LAUNCH(2)
        case 3: PB_W1_LAUNCH(3)
        default: PB_W1_LAUNCH(4)
      }
#undef PB_W1_LAUNCH
    }
  }
  if (pl.gstate) {
    if (pl.act == PBNN_ACT_RELU) return pb_launch(pb_sg_kernel<ALGO, PBNN_ACT_RELU, true, 0>, pl, a, dim3(C), stream);
    return pb_launch(pb_sg_kernel<ALGO, PBNN_ACT_TANH, true, 0>, pl, a, dim3(C), stream);
  }
  if (pl.h1) {
    if (pl.act == PBNN_ACT_RELU) return pb_launch(pb_sg_kernel<ALGO, PBNN_ACT_RELU, false, 1>, pl, a, dim3(C), stream);
    return pb_launch(pb_sg_kernel<ALGO, PBNN_ACT_TANH, false, 1>, pl, a, dim3(C), stream);
  }
  if (pl.act == PBNN_ACT_RELU) return pb_launch(pb_sg_kernel<ALGO, PBNN_ACT_RELU, false, 0>, pl, a, dim3(C), stream);
  return pb_launch(pb_sg_kernel<ALGO, PBNN_ACT_TANH, false, 0>, pl, a, dim3(C), stream);
}

static int pb_backend_sg(int algo, const PbPlan& pl, const PbSgArgs& a, int C, void* stream) {
  switch (algo) {
    case PB_ALGO_SGLD: return pb_sg_dispatch<PB_ALGO_SGLD>(pl, a, C, stream);
    case PB_ALGO_PSGLD: return pb_sg_dispatch<PB_ALGO_PSGLD>(pl, a, C, stream);
    case PB_ALGO_SGHMC: return pb_sg_dispatch<PB_ALGO_SGHMC>(pl, a, C, stream);
    case PB_ALGO_ADAM: return pb_sg_dispatch<PB_ALGO_ADAM>(pl, a, C, stream);
    default: return pb_sg_dispatch<PB_ALGO_GRAD>(pl, a, C, stream);
  }
}

static int pb_backend_hmc(const PbPlan& pl, const PbHmcArgs& a, int C, void* stream) {
  if (pl.gstate) {
    if (pl.act == PBNN_ACT_RELU) return pb_launch(pb_hmc_kernel<PBNN_ACT_RELU, true, 0>, pl, a, dim3(C), stream);
    return pb_launch(pb_hmc_kernel<PBNN_ACT_TANH, true, 0>, pl, a, dim3(C), stream);
  }
  if (pl.tc) {
    if (a.sched_done) {  // balanced persistent CTAs, two per SM (by construction of the plan: <= 113 KB, 160 threads)
      const long long nitems = (long long)((a.S + a.seg_len - 1) / a.seg_len) * C;
      const int slots = 2 * pb_backend_num_sms();
      const int grid = slots < nitems ? slots : (int)nitems;
      if (pb_opt(PB_OPT_DEBUG, 0))
        fprintf(stderr, "[pbnn] hmc tc sched: grid %d, threads %d, smem %d, seg_len %d, items %lld\n", grid, pl.NT,
                pl.smem_floats * 4, a.seg_len, nitems);
      PB_CUDA(cudaMemsetAsync(a.sched_done, 0, sizeof(int) * (size_t)(C + 1), (cudaStream_t)stream));
      return pl.tc_npad > 512 ? pb_launch(pb_hmc_tc_kernel<3>, pl, a, dim3(grid), stream)
                              : pb_launch(pb_hmc_tc_kernel<2>, pl, a, dim3(grid), stream);
    }
    return pl.tc_npad > 512 ? pb_launch(pb_hmc_tc_kernel<3>, pl, a, dim3(C), stream)
                            : pb_launch(pb_hmc_tc_kernel<2>, pl, a, dim3(C), stream);
  }
  if (pl.h1) {
    if (pl.act == PBNN_ACT_RELU) return pb_launch(pb_hmc_kernel<PBNN_ACT_RELU, false, 1>, pl, a, dim3(C), stream);
    return pb_launch(pb_hmc_kernel<PBNN_ACT_TANH, false, 1>, pl, a, dim3(C), stream);
  }
  if (pl.act == PBNN_ACT_RELU) return pb_launch(pb_hmc_kernel<PBNN_ACT_RELU, false, 0>, pl, a, dim3(C), stream);
  return pb_launch(pb_hmc_kernel<PBNN_ACT_TANH, false, 0>, pl, a, dim3(C), stream);
}

static int pb_backend_predict(const PbPlan& pl, const PbPredArgs& a, int mtiles, float* sum, float* sumsq,
                              void* stream) {
  const PbLayer& L0 = pl.L[0];
  if (pl.nl == 2 && pl.s == 1 && pl.act == PBNN_ACT_RELU && L0.out <= 64 && pl.d + 1 <= 16 && !pl.gstate &&
      !pb_opt(PB_OPT_NO_TC, 0)) {
    // tcgen05 path; tiles per CTA: as many as still leave every CTA slot (2 per SM) with work
    PbPredTcArgs t;
    t.samples = a.samples; t.S = a.S; t.Xtest = a.Xtest; t.M = a.M; t.preds = a.preds; t.part = a.part;
    t.nchunk = a.nchunk; t.chunk = a.chunk;
    t.d = pl.d; t.H = L0.out; t.P = pl.P; t.fb0 = L0.fb; t.fw0 = L0.fw; t.fb1 = pl.L[1].fb; t.fw1 = pl.L[1].fw;
    const int tiles = (a.M + 127) / 128, slots = 2 * pb_backend_num_sms();
    t.tpc = pb_opt(PB_OPT_PRED_TPC, 4);
    t.tpc = t.tpc >= 4 ? 4 : (t.tpc >= 2 ? 2 : 1);  // the kernel holds 4 X tiles in its 256 TMEM columns
    while (t.tpc > 1 && (long long)((tiles + t.tpc - 1) / t.tpc) * a.nchunk < slots) t.tpc >>= 1;
    const dim3 g((tiles + t.tpc - 1) / t.tpc, a.nchunk);
    pb_predict_tc_kernel<<<g, 160, 0, (cudaStream_t)stream>>>(t);
    PB_CUDA(cudaGetLastError());
    if (sum) {
      const int n = a.M;
      pb_predict_reduce_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a.part, a.nchunk, n, sum, sumsq);
      PB_CUDA(cudaGetLastError());
    }
    return PBNN_OK;
  }
  const dim3 grid(mtiles, a.nchunk);
  int rc;
  if (pl.gstate) {
    if (pl.act == PBNN_ACT_RELU) rc = pb_launch(pb_predict_kernel<PBNN_ACT_RELU, true>, pl, a, grid, stream);
    else rc = pb_launch(pb_predict_kernel<PBNN_ACT_TANH, true>, pl, a, grid, stream);
  } else {
    if (pl.act == PBNN_ACT_RELU) rc = pb_launch(pb_predict_kernel<PBNN_ACT_RELU, false>, pl, a, grid, stream);
    else rc = pb_launch(pb_predict_kernel<PBNN_ACT_TANH, false>, pl, a, grid, stream);
  }
  if (rc) return rc;
  if (sum) {
    const int n = a.M * pl.s;
    pb_predict_reduce_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a.part, a.nchunk, n, sum, sumsq);
    PB_CUDA(cudaGetLastError());
  }
  return PBNN_OK;
}

static int pb_backend_predict_tc2(const PbPlan& pl, const PbPredTc2Args& a, int groups, float* sum, float* sumsq,
                                  void* stream) {
  const size_t smem = (size_t)pl.smem_floats * 4;
  const int rc = pb_ensure_smem(reinterpret_cast<const void*>(pb_predict_tc2_kernel), smem);
  if (rc) return rc;
  pb_predict_tc2_kernel<<<dim3(groups, a.nchunk), PB_T2P_NT, smem, (cudaStream_t)stream>>>(pl, a);
  PB_CUDA(cudaGetLastError());
  if (sum) {
    pb_predict_reduce_kernel<<<(a.M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a.part, a.nchunk, a.M, sum, sumsq);
    PB_CUDA(cudaGetLastError());
  }
  return PBNN_OK;
}

static int pb_backend_rng(int kind, const uint32_t* keys, int C, int n, void* out, void* stream) {
  const size_t total = kind == 4 ? (size_t)C : (size_t)C * n;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pb_rng_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(kind, keys, C, n, out);
  PB_CUDA(cudaGetLastError());
  return PBNN_OK;
}

static int pb_backend_mbidx(const uint32_t* keys, int C, int draw, int B, int N, uint32_t mult, int32_t* out,
                            void* stream) {
  pb_mbidx_kernel<<<C, 128, 0, (cudaStream_t)stream>>>(keys, draw, B, N, mult, out);
  PB_CUDA(cudaGetLastError());
  return PBNN_OK;
}

static int pb_perm_rounds(int N) {
  return (int)ceil(3.0 * log((double)(N > 1 ? N : 1)) / log(4294967295.0));
}

static int pb_backend_perm(const uint32_t* keys, int C, int N, int32_t* out, void* ws, void* stream) {
  int Np2 = 1;
  while (Np2 < N) Np2 <<= 1;
  const size_t bytes = (size_t)Np2 * 8;
  const int use_smem = bytes <= 200 * 1024;
  if (use_smem) {
    const int rc = pb_ensure_smem(reinterpret_cast<const void*>(pb_perm_kernel), bytes);
    if (rc) return rc;
  }
  pb_perm_kernel<<<C, 1024, use_smem ? bytes : 0, (cudaStream_t)stream>>>(keys, N, Np2, pb_perm_rounds(N), use_smem,
                                                                          out, (unsigned long long*)ws);
  PB_CUDA(cudaGetLastError());
  return PBNN_OK;
}

#include "pbnn_api.inc"

extern "C" int pbnn_thinning(const float* positions, int n, int D, int size, int32_t* idx_out, float* workspace,
                             void* stream) {
  if (!positions || !idx_out || !workspace || n < 1 || D < 1 || size < 1 || size > n) {
    snprintf(g_err, sizeof(g_err), "thinning: need positions, idx_out, workspace, 1 <= size <= n");
    return PBNN_EINVAL;
  }
  float* xx = workspace;
  float* k0sum = workspace + n;
  float* obj = workspace + 2 * (size_t)n;
  float* pval = workspace + 3 * (size_t)n;
  int32_t* pidx = reinterpret_cast<int32_t*>(workspace + 3 * (size_t)n + 1024);
  int32_t* cur = reinterpret_cast<int32_t*>(workspace + 3 * (size_t)n + 2048);
  cudaStream_t st = (cudaStream_t)stream;
  int nblocks = (n + 7) / 8;
  if (nblocks > 148 * 4) nblocks = 148 * 4;
  pb_thin_norm_kernel<<<(n + 7) / 8, 256, 0, st>>>(positions, n, D, xx);
  pb_thin_rowsum_kernel<<<(n + 63) / 64, 256, 0, st>>>(positions, xx, n, D, k0sum);
  for (int step = 0; step < size; ++step) {
    pb_thin_update_kernel<<<nblocks, 256, 0, st>>>(positions, xx, k0sum, n, D, step, cur, obj, pval, pidx);
    pb_thin_argmin_kernel<<<1, 256, 0, st>>>(pval, pidx, nblocks, step, size, cur, idx_out);
  }
  PB_CUDA(cudaGetLastError());
  return PBNN_OK;
}

// measurement utility (not part of the sampler ABI): launches the FP32 FMA probe; flops = grid*256*iters*64
extern "C" int pbnn_probe_fp32(int mode, int iters, int ctas, float* sink, void* stream) {
  pb_fma_probe_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(mode, iters, 1.0f, sink);
  PB_CUDA(cudaGetLastError());
  return PBNN_OK;
}

#if defined(PB_PROFILE)
// profiling build only: per-phase cycle totals [0..31] and phase counts [32..63] of CTA 0; resets the counters
extern "C" int pbnn_debug_profile(unsigned long long out[64]) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, pb_prof_cycles, sizeof(unsigned long long) * 64);
  unsigned long long zero[64] = {0};
  cudaMemcpyToSymbol(pb_prof_cycles, zero, sizeof(zero));
  return 0;
}
#endif
