// pbnn_b200: posterior predictive for nets d-H-H-1 (two equal hidden relu layers) on the tcgen05 engine of
// pbnn_tc2.cuh -- predict_fn (13 copies in the reference, e.g. src/pbnn/mcmc/langevin.py:75-76: vmap over the stored
// positions of network.apply) and its mean / variance over the samples (benchmark/pipelines/
// metrics_prediction_intervals.py:188) for the nets pbnn's own MLP produces (src/pbnn/models.py:19-35).
//
// Shape of the work: S samples x M test rows.  A CTA owns `tpc` 128-row tiles of X_test (their bf16x3 operand images
// sit in the CTA's global image block for the whole kernel) and a chunk of the samples.  Per sample: the flat parameter
// vector is split into the operand images of W1aug and W2^T once (the restage of pbnn_tc2.cuh, reading the FLAT
// ravel_pytree layout), then every tile runs the forward half of the chain-step engine -- G1 Z1 = Xaug W1aug, relu
// chunks written back as bf16x3 operands, G2 Z2 = relu(Z1) W2 (6 passes each, fp32 accumulation in TMEM), and
// f = w3 . relu(Z2 + b2) + b3 on the CUDA cores.  sum f and sum f^2 per test row accumulate in shared memory over the
// CTA's samples and leave as one partial per (sample chunk, row), reduced by pb_predict_reduce_kernel.
// Roles (320 threads): warps 0-7 workers, warp 8 bulk-copy producer, warp 9 MMA issue.
#pragma once
#if !defined(PB_EMU)

struct PbPredTc2Args {
  const float* samples;  // [S][P] flat
  int S;
  const float* Xtest;  // [M][d]
  int M;
  float* preds;        // [S][M] or NULL
  float* part;         // [nchunk][2][M] or NULL
  int nchunk, chunk;   // sample chunks (grid.y) / samples per chunk
  int tpc;             // tiles per CTA (<= 4)
  unsigned char* img;  // [grid.y * grid.x] image blocks
  long long img_bytes; // bytes of one block: tpc X images | W1aug image | W2^T image (1024-byte multiple)
};

#define PB_T2P_NT 320

// 128 test rows from row m0 on -> bf16x3 image of Xaug = (x, 1, 0..) [128 rows][Kx] (rows >= M are zero)
PB_DEV void pb_tc2p_stage_x(const PbPlan& pl, uint8_t* ximg, const float* X, int m0, int M, int nt) {
  const int tid = threadIdx.x, d = pl.d, Kx = pl.t2_Kx;
  const int tbytes = pl.t2_ncbx * PB_T2_TILE;
  for (int e = tid; e < 128 * (Kx >> 1); e += nt) {
    const int b = e / (Kx >> 1), k0 = 2 * (e - b * (Kx >> 1));
    float v0 = 0.f, v1 = 0.f;
    if (m0 + b < M) {
      const size_t row = (size_t)(m0 + b);
      v0 = k0 < d ? X[row * d + k0] : (k0 == d ? 1.f : 0.f);
      v1 = k0 + 1 < d ? X[row * d + k0 + 1] : (k0 + 1 == d ? 1.f : 0.f);
    }
    uint32_t t1, t2, t3;
    pb_t2_split3(v0, v1, t1, t2, t3);
    uint8_t* p = ximg + pb_t2_off(128, b, k0);
    *reinterpret_cast<uint32_t*>(p) = t1;
    *reinterpret_cast<uint32_t*>(p + tbytes) = t2;
    *reinterpret_cast<uint32_t*>(p + 2 * tbytes) = t3;
  }
}

// one sample (FLAT layout: bias before kernel, kernel [in][out]) -> operand images of W1aug and W2^T, small vectors
PB_DEV void pb_tc2p_restage(const PbPlan& pl, float* sm, uint8_t* img1, uint8_t* img2, const float* __restrict__ th,
                            int nt) {
  const int tid = threadIdx.x;
  const PbLayer& L0 = pl.L[0];
  const PbLayer& L1 = pl.L[1];
  const PbLayer& L2 = pl.L[2];
  const int H = pl.t2_H, Hp = pl.t2_Hp, Kx = pl.t2_Kx, d = pl.d;
  {
    const int tb = pl.t2_nch * Kx * 128, hp2 = Hp >> 1;
    const int total = (d + 1) * hp2;
    for (int e = tid; e < total; e += nt) {
      const int k = e / hp2, n0 = 2 * (e - k * hp2);
      const float* row = k < d ? th + L0.fw + k * H : th + L0.fb;
      // (loads from clamped addresses, selects AFTER: a zero written to the register a load is still filling makes the
      //  loads of a thread wait for each other -- measured in pbnn_warp2.cuh)
      const float l0 = row[n0 < H ? n0 : 0], l1 = row[n0 + 1 < H ? n0 + 1 : 0];
      const float v0 = n0 < H ? l0 : 0.f, v1 = n0 + 1 < H ? l1 : 0.f;
      uint32_t t1, t2, t3;
      pb_t2_split3(v0, v1, t1, t2, t3);
      uint8_t* p = img1 + pb_t2_off(Kx, k, n0);
      *reinterpret_cast<uint32_t*>(p) = t1;
      *reinterpret_cast<uint32_t*>(p + tb) = t2;
      *reinterpret_cast<uint32_t*>(p + 2 * tb) = t3;
    }
  }
  {
    // W2^T image: row n, 16-byte chunk = 8 consecutive input indices i.  One item = (chunk of 8 i, output unit n):
    // consecutive threads take consecutive n (coalesced reads of the flat [i][n] kernel), each writes whole chunks
    const int tb = pl.t2_nch * Hp * 128, nck = Hp >> 3;
    const int total = H * nck;
    for (int e = tid; e < total; e += nt) {
      const int cb = e / H, n = e - cb * H, i0 = 8 * cb;
      float v[8], l[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) l[j] = th[L1.fw + (i0 + j < H ? i0 + j : 0) * H + n];  // (all eight loads in flight)
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = i0 + j < H ? l[j] : 0.f;
      uint32_t t1[4], t2[4], t3[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) pb_t2_split3(v[2 * m], v[2 * m + 1], t1[m], t2[m], t3[m]);
      uint8_t* p = img2 + pb_t2_off(Hp, n, i0);
      *reinterpret_cast<uint4*>(p) = make_uint4(t1[0], t1[1], t1[2], t1[3]);
      *reinterpret_cast<uint4*>(p + tb) = make_uint4(t2[0], t2[1], t2[2], t2[3]);
      *reinterpret_cast<uint4*>(p + 2 * tb) = make_uint4(t3[0], t3[1], t3[2], t3[3]);
    }
  }
  float* b2v = sm + pl.off_t2_vec;
  float* w3v = b2v + 256;
  for (int n = tid; n < 256; n += nt) {
    const float lb = th[L1.fb + (n < H ? n : 0)], lw = th[L2.fw + (n < H ? n : 0)];
    b2v[n] = n < H ? lb : 0.f;
    w3v[n] = n < H ? lw : 0.f;
  }
  if (tid == 0) sm[pl.off_t2_vec + 1408] = th[L2.fb];
  pb_t2_fence_async_global();
}

// forward of one 128-row tile: returns f for row b in the threads of column group 0 (other threads: undefined)
PB_DEV float pb_tc2p_forward(const PbPlan& pl, float* sm, const PbTc2& tc, const uint8_t* ximg, const uint8_t* img1,
                             const uint8_t* img2) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int H = pl.t2_H, Hp = pl.t2_Hp, Kx = pl.t2_Kx;
  const int nch = pl.t2_nch, ncbx = pl.t2_ncbx;
  const uint32_t acta = tc.big, ring = tc.big + 3u * PB_T2_TILE + (uint32_t)nch * PB_T2_TILE;
  const uint32_t wstage = (uint32_t)pl.t2_wstage;
  const uint32_t S0 = tc.tmem, S1 = tc.tmem + 256u;
  float* b2v = sm + pl.off_t2_vec;
  float* w3v = b2v + 256;
  float* fp = b2v + 640;
  float fout = 0.f;
  (void)H;

  if (tid == 0) {
    for (int i = 0; i <= PB_T2_B_ACC; ++i) {
      const uint32_t b = pb_t2_bar(pl, sm, i);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(i == PB_T2_B_AFULL ? PB_T2_WORKERS : 1u));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  pb_tc_fence_before();
  __syncthreads();
  pb_tc_fence_after();
  const uint32_t bar0 = pb_t2_bar(pl, sm, 0);
#define PB_T2_BAR(i) (bar0 + 8u * (uint32_t)(i))

  if (warp_u == 8) {
    if (pb_elect_one()) {
      int fill = 0;
      auto stage_begin = [&](uint32_t bytes) -> uint32_t {
        const int s = fill % PB_T2_NS;
        if (fill >= PB_T2_NS) pb_mbar_wait(PB_T2_BAR(PB_T2_B_WEMPTY + s), (uint32_t)((fill / PB_T2_NS - 1) & 1));
        pb_t2_expect(PB_T2_BAR(PB_T2_B_WFULL + s), bytes);
        return ring + (uint32_t)s * wstage;
      };
      // G1: X chunk kc (3 terms) -> activation area; W1aug rows of chunk kc (one term per ring stage)
      for (int kc = 0; kc < ncbx; ++kc) {
        const int rows = (Kx - 64 * kc) < 64 ? (Kx - 64 * kc) : 64;
        if (kc > 0) pb_mbar_wait(PB_T2_BAR(PB_T2_B_AEMPTY), (uint32_t)((kc - 1) & 1));
        pb_t2_expect(PB_T2_BAR(PB_T2_B_AXFULL), 3u * PB_T2_TILE);
        for (int t = 0; t < 3; ++t)
          pb_t2_bulk(acta + (uint32_t)t * PB_T2_TILE, ximg, (uint32_t)(t * ncbx + kc) * PB_T2_TILE, PB_T2_TILE,
                     PB_T2_BAR(PB_T2_B_AXFULL));
        for (int t = 0; t < 3; ++t) {
          const uint32_t dst = stage_begin((uint32_t)(nch * rows * 128));
          const int s = fill % PB_T2_NS;
          for (int cb = 0; cb < nch; ++cb)
            pb_t2_bulk(dst + (uint32_t)cb * 8192u, img1, (uint32_t)((t * nch + cb) * Kx + 64 * kc) * 128u,
                       (uint32_t)(rows * 128), PB_T2_BAR(PB_T2_B_WFULL + s));
          ++fill;
        }
      }
      // G2: column block kc of the transposed W2 image, all Hp rows (read K-major)
      for (int kc = 0; kc < nch; ++kc)
        for (int t = 0; t < 3; ++t) {
          const uint32_t bytes = (uint32_t)(Hp * 128);
          const uint32_t dst = stage_begin(bytes);
          const int s = fill % PB_T2_NS;
          const uint32_t so = (uint32_t)((t * nch + kc) * Hp) * 128u;
          for (uint32_t o = 0; o < bytes; o += 16384u)
            pb_t2_bulk(dst + o, img2, so + o, bytes - o < 16384u ? bytes - o : 16384u, PB_T2_BAR(PB_T2_B_WFULL + s));
          ++fill;
        }
    }
    __syncwarp();
  } else if (warp_u == 9) {
    int fill = 0;
    uint32_t axph = 0u, afph = 0u;
    auto kgemm = [&](uint32_t dcol, int nchunks, int ktotal, bool a_from_tma, bool b_kmajor) {
      const uint32_t idesc = pb_t2_idesc(128, Hp, 0, b_kmajor ? 0 : 1);
      for (int c = 0; c < nchunks; ++c) {
        const int ksteps = ((ktotal - 64 * c) < 64 ? (ktotal - 64 * c) : 64) >> 4;
        if (a_from_tma) {
          pb_mbar_wait(PB_T2_BAR(PB_T2_B_AXFULL), axph);
          axph ^= 1u;
        } else {
          pb_mbar_wait(PB_T2_BAR(PB_T2_B_AFULL), afph);
          afph ^= 1u;
        }
        pb_tc_fence_after();
        for (int wt = 0; wt < 3; ++wt) {
          const int s = fill % PB_T2_NS;
          pb_mbar_wait(PB_T2_BAR(PB_T2_B_WFULL + s), (uint32_t)((fill / PB_T2_NS) & 1));
          pb_tc_fence_after();
          const uint32_t st = ring + (uint32_t)s * wstage;
          if (pb_elect_one()) {
            for (int xa = 0; xa + wt < 3; ++xa)
              for (int ks = 0; ks < ksteps; ++ks)
                pb_t2_mma(dcol, pb_t2_desc(acta + (uint32_t)xa * PB_T2_TILE + (uint32_t)ks * 32u, 16u, 1024u),
                          b_kmajor ? pb_t2_desc(st + (uint32_t)ks * 32u, 16u, 1024u)
                                   : pb_t2_desc(st + (uint32_t)ks * 2048u, 8192u, 1024u),
                          idesc, (c | wt | xa | ks) ? 1u : 0u);
          }
          __syncwarp();
          // (the commit sits in its OWN elect block, as in pbnn_tc2.cuh: issued from inside the block that issues the MMAs,
          // behind the same predicate, the kernel died with an illegal-address fault on this toolchain -- bisected on a B200)
          if (pb_elect_one()) pb_umma_commit(PB_T2_BAR(PB_T2_B_WEMPTY + s));
          __syncwarp();
          ++fill;
        }
        if (pb_elect_one()) pb_umma_commit(PB_T2_BAR(PB_T2_B_AEMPTY));
        __syncwarp();
      }
      if (pb_elect_one()) pb_umma_commit(PB_T2_BAR(PB_T2_B_ACC));
      __syncwarp();
    };
    kgemm(S0, ncbx, Kx, true, false);  // G1: Z1
    kgemm(S1, nch, Hp, false, true);   // G2: Z2
    // Tail: every tcgen05.commit of this evaluation must have ARRIVED before the barriers are re-initialised by the next
    // evaluation or the CTA exits.  The workers only wait for the accumulator barrier; the commits that free the last
    // ring stages and the last activation chunk are separate asynchronous arrivals on other mbarriers, and nothing
    // orders them before the accumulator's.
    for (int i = fill > PB_T2_NS ? fill - PB_T2_NS : 0; i < fill; ++i)
      pb_mbar_wait(PB_T2_BAR(PB_T2_B_WEMPTY + i % PB_T2_NS), (uint32_t)((i / PB_T2_NS) & 1));
    pb_mbar_wait(PB_T2_BAR(PB_T2_B_AEMPTY), (uint32_t)((ncbx + nch - 1) & 1));
  } else {
    const int q = warp & 3, g = warp >> 2;
    const int b = 32 * q + lane;
    const uint32_t lane_off = (uint32_t)(32 * q) << 16;
    const uint32_t rowaddr = (uint32_t)b * 128u;
    const uint32_t bx = (uint32_t)(b & 7);
    int ause = ncbx;
    pb_mbar_wait(PB_T2_BAR(PB_T2_B_ACC), 0u);  // Z1
    pb_tc_fence_after();
    for (int ic = 0; ic < nch; ++ic) {
      uint32_t z[32];
      pb_tmem_ld32(S0 + lane_off + (uint32_t)(64 * ic + 32 * g), z);
      pb_tmem_wait_ld();
      pb_mbar_wait(PB_T2_BAR(PB_T2_B_AEMPTY), (uint32_t)((ause - 1) & 1));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t t1[4], t2[4], t3[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          pb_t2_split3(fmaxf(__uint_as_float(z[8 * u + 2 * j]), 0.f), fmaxf(__uint_as_float(z[8 * u + 2 * j + 1]), 0.f),
                       t1[j], t2[j], t3[j]);
        const uint32_t a = acta + rowaddr + ((((uint32_t)(4 * g + u)) ^ bx) << 4);
        pb_t2_st4(a, t1[0], t1[1], t1[2], t1[3]);
        pb_t2_st4(a + PB_T2_TILE, t2[0], t2[1], t2[2], t2[3]);
        pb_t2_st4(a + 2u * PB_T2_TILE, t3[0], t3[1], t3[2], t3[3]);
      }
      pb_fence_async_smem();
      pb_tc_fence_before();
      pb_t2_arrive(PB_T2_BAR(PB_T2_B_AFULL));
      ++ause;
    }
    pb_mbar_wait(PB_T2_BAR(PB_T2_B_ACC), 1u);  // Z2
    pb_tc_fence_after();
    float fpart = 0.f;
    for (int ic = 0; ic < nch; ++ic) {
      uint32_t z[32];
      pb_tmem_ld32(S1 + lane_off + (uint32_t)(64 * ic + 32 * g), z);
      pb_tmem_wait_ld();
      const int c0 = 64 * ic + 32 * g;
#pragma unroll
      for (int c = 0; c < 32; ++c) fpart = fmaf(fmaxf(__uint_as_float(z[c]) + b2v[c0 + c], 0.f), w3v[c0 + c], fpart);
    }
    pb_tc_fence_before();
    fp[g * 128 + b] = fpart;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    fout = (fp[b] + fp[128 + b]) + sm[pl.off_t2_vec + 1408];
  }
#undef PB_T2_BAR
  pb_tc_fence_before();
  __syncthreads();
  pb_tc_fence_after();
  return fout;
}

__global__ void __launch_bounds__(PB_T2P_NT, 1)
pb_predict_tc2_kernel(const __grid_constant__ PbPlan pl, const __grid_constant__ PbPredTc2Args a) {
  extern __shared__ __align__(16) float pb_smem[];
  float* sm = pb_smem;
  const int tid = threadIdx.x, nt = PB_T2P_NT;
  const int grp = blockIdx.x, ch = blockIdx.y;
  const int ncbx = pl.t2_ncbx, nch = pl.t2_nch, Kx = pl.t2_Kx;
  const int xi = 3 * ncbx * PB_T2_TILE;
  uint8_t* img = a.img + (size_t)(ch * gridDim.x + grp) * (size_t)a.img_bytes;
  uint8_t* img1 = img + (size_t)a.tpc * xi;
  uint8_t* img2 = img1 + 3 * nch * Kx * 128;
  // TMEM, zeroed image block (pad rows / columns of the images must read as zeros)
  PbTc2 tc;
  {
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + pl.off_t2_bar) + 2 * PB_T2_NBAR;
    if (tid < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(pb_smem_u32(slot)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    uint4* z = reinterpret_cast<uint4*>(img);
    for (long long i = tid; i < a.img_bytes / 16; i += nt) z[i] = make_uint4(0u, 0u, 0u, 0u);
    pb_tc_fence_before();
    __syncthreads();
    pb_tc_fence_after();
    tc.tmem = *slot;
    tc.img = img;
    tc.big = (pb_smem_u32(sm + pl.off_t2_big) + 1023u) & ~1023u;
  }
  float* acc = sm + pl.off_scr;  // [2][tpc][128] running sum f, sum f^2 of this CTA's rows
  for (int i = tid; i < 2 * a.tpc * 128; i += nt) acc[i] = 0.f;
  for (int t = 0; t < a.tpc; ++t) {
    const int m0 = (grp * a.tpc + t) * 128;
    if (m0 < a.M) pb_tc2p_stage_x(pl, img + (size_t)t * xi, a.Xtest, m0, a.M, nt);
  }
  pb_t2_fence_async_global();
  __syncthreads();
  const int s0 = ch * a.chunk, s1 = s0 + a.chunk < a.S ? s0 + a.chunk : a.S;
  const bool lead = tid < 128;  // workers of column group 0: thread = row
  for (int s = s0; s < s1; ++s) {
    pb_tc2p_restage(pl, sm, img1, img2, a.samples + (size_t)s * pl.P, nt);
    __syncthreads();
    for (int t = 0; t < a.tpc; ++t) {
      const int m0 = (grp * a.tpc + t) * 128;
      if (m0 >= a.M) break;
      const float f = pb_tc2p_forward(pl, sm, tc, img + (size_t)t * xi, img1, img2);
      if (lead && m0 + tid < a.M) {
        acc[t * 128 + tid] += f;
        acc[(a.tpc + t) * 128 + tid] += f * f;
        if (a.preds) a.preds[(size_t)s * a.M + m0 + tid] = f;
      }
    }
  }
  __syncthreads();
  if (a.part && lead)
    for (int t = 0; t < a.tpc; ++t) {
      const int row = (grp * a.tpc + t) * 128 + tid;
      if (row < a.M) {
        a.part[((size_t)ch * 2) * a.M + row] = acc[t * 128 + tid];
        a.part[((size_t)ch * 2 + 1) * a.M + row] = acc[(a.tpc + t) * 128 + tid];
      }
    }
  pb_tc_fence_before();
  __syncthreads();
  if (tid < 32) {
    pb_tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tc.tmem));
  }
}
#endif
