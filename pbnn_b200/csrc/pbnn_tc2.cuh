// pbnn_b200: tcgen05 gradient engine for nets d-H-H-1 (two equal hidden relu layers, scalar output) -- the minibatch
// log-likelihood gradient behind blackjax's grad_estimator (src/pbnn/utils/misc.py:34-55 with
// src/pbnn/models.py:19-35 and benchmark/pipelines/logprobs.py:12-17) for BASELINE configs 3-5 (9-100-100-1,
// 8-50-50-1, 90-256-256-1; B <= 128 rows = ONE 128-row MMA tile).
//
// FP32 accuracy on bf16 tensor cores.  Every fp32 operand a is fed as a = a1 + a2 + a3 (three bf16 terms, exact to 24
// bits); a product keeps the term pairs with i + j <= 4 (6 MMA passes, fp32 accumulation in TMEM; measured 7e-7 relative
// to the exact product, pbnn_b200/csrc/experimental/umma_bf16.cu).  A 0/1 relu mask is exact in ONE term, so the two
// backward GEMMs that can be written against the mask take 3 passes:
//
//   G1 fwd1   Z1 = Xaug W1aug                      A = X tile   (K-major)   B = W1aug (MN-major)            6 passes
//   G2 fwd2   Z2 = relu(Z1) W2                     A = H1 chunk (K-major)   B = W2^T image read K-major     6 passes
//   G3        f = w3 . relu(Z2 + b2) + b3,  r = (y - f) scale,  M2 = [Z2 + b2 > 0]           (CUDA cores)
//   G5 dH1    E[b][i]  = sum_n (w3[n] M2[b][n]) W2[i][n]    A = w3 * M2 chunk (K-major)  B = W2^T image read MN-major  6 passes
//             (the chunk terms are SELECTED from a per-step 3-term split of w3: no arithmetic per element)
//   G6 dW1    D1[k][i] = sum_b Xaug[b][k] dZ1[b][i],  dZ1 = r * E * [Z1 > 0]   A = X^T (MN-major)  B = dZ1 chunk (MN-major) 6 passes
//   G4 dW2    D2[n][i] = sum_b M2[b][n] Y[b][i],   Y = r * [relu(Z1) | 1]   A = M2^T (MN-major)  B = Y chunk (MN-major)   3 passes
//             (last: its drains rewrite the W2 image)  dW2[i][n] = w3[n] D2[n][i],  db2[n] = w3[n] D2[n][H],
//             dw3[n] = sum_i W2[i][n] D2[n][i] + b2[n] D2[n][H]      (relu(Z2 + b2) = M2 * (H1 W2 + b2): no extra GEMM)
//
// Operand images are SWIZZLE_128B tiles: 64 bf16 (128 B) per storage row, 16-byte chunk j of row r at chunk j ^ (r & 7),
// column blocks of 64 elements one after the other.  The SAME image is read K-major (storage row = M/N index) or
// MN-major (storage row = K index): X, M2 and the activation chunks are written once and used both ways, and W2 is kept
// TRANSPOSED (storage row = output unit n) so that the thread draining TMEM lane n owns one image row.
//
// Where things live.  Per CTA a global IMAGE block (L2) holds the bf16x3 operand images of X (fixed minibatch), W1aug
// and W2^T plus the noise ring.  The images are the MASTER copy of W1, b1, W2 (three bf16 terms = the fp32 value,
// exactly): no fp32 copy of them is read or written per step; b2, w3, b3 stay fp32 in the chain's workspace slice.
// The images are streamed through a 3-stage shared-memory ring with cp.async.bulk (TMA engine, UBLKCP) + mbarriers by
// a dedicated producer warp.
//
// Fused drains.  The weight-gradient accumulators are never written out: the thread that reads eight consecutive
// dW values of its image row from TMEM also loads the eight weights (one 16-byte chunk per term), takes the eight
// normals of the block from the noise ring, applies the caller's update functor (SGLD / SGHMC / Adam / pSGLD step),
// splits the new values and stores the three chunks back -- and, on kept steps, the fp32 values to the trace row /
// final state.  One pass over the parameters per step: no gradient array, no separate update, re-staging or trace pass.
// Shared memory: one activation chunk [128 rows][64] x 3 terms (48 KB), the layer-2 mask [128][Hp] bf16, the weight
// ring.  TMEM (512 columns): S0 = Z1 (kept for G4);  S1 = Z2 -> E -> D1 -> D2 chunk slots.
//
// Roles (576 threads): warps 0-7 workers (TMEM -> registers -> operand chunks / drains; warp w owns TMEM lanes
// 32 (w % 4) .., column half w / 4), warp 8 producer (bulk copies), warp 9 MMA issue, warps 10-17 noise: thread t of
// them mirrors worker thread t, walks the drains' block sequence and writes each block's eight jax.random.normal
// draws into a ring (mbarrier full / empty per slot) up to t2_nz blocks ahead; when a whole step's noise fits the ring
// (small nets) the same warps then drain every other D2 step.  All mbarriers are re-initialised at the start of every
// gradient evaluation (the evaluation is bracketed by CTA barriers), so every parity starts at 0.
#pragma once
#if !defined(PB_EMU)

#define PB_T2_WORKERS 256

// 4 normals behind a call boundary: inside the drains the register allocator otherwise re-serialises the four
// threefry chains (measured: LG2 instructions ~100 apart in the SASS, 2-3x slower drains); as a separate function the
// chains are scheduled on their own, interleaved, and cost the caller one call + a float4 return
__device__ __noinline__ float4 pb_t2_normal4(uint32_t k0, uint32_t k1, uint32_t i0, uint32_t i1, uint32_t i2,
                                             uint32_t i3) {
  PbKey k;
  k.k0 = k0;
  k.k1 = k1;
  const uint32_t ix[4] = {i0, i1, i2, i3};
  float z[4];
  pb_normal_n<4, true>(k, ix, z);
  return make_float4(z[0], z[1], z[2], z[3]);
}

PB_DEV uint64_t pb_t2_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {  // SWIZZLE_128B
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// instruction descriptor: D = f32, A = B = bf16, dense; a_mn / b_mn: operand is MN-major
PB_DEV uint32_t pb_t2_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

PB_DEV void pb_t2_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

// (a0, a1) -> three packed bf16x2 words (terms 1..3); element 0 in the low half
PB_DEV void pb_t2_split3(float a0, float a1, uint32_t& t1, uint32_t& t2, uint32_t& t3) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t1) : "f"(a1), "f"(a0));
  float r0 = a0 - __uint_as_float(t1 << 16);
  float r1 = a1 - __uint_as_float(t1 & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t2) : "f"(r1), "f"(r0));
  r0 -= __uint_as_float(t2 << 16);
  r1 -= __uint_as_float(t2 & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t3) : "f"(r1), "f"(r0));
}

PB_DEV void pb_t2_arrive(uint32_t mbar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(mbar) : "memory");
}
PB_DEV void pb_t2_expect(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
// Bulk copy global -> shared.  The source is (base, 32-bit byte offset) and the 64-bit sum is formed INSIDE the asm
// block: with the offset arithmetic done in 64 bits in C++ (((size_t)t * ncbx + kc) * 16384 and the like), ptxas
// 12.9 turned one of three unrolled copies of pb_tc2p_forward into a uniform-datapath ULEA that drops the upper
// word of the address (SASS: UMOV UR7, URZ; ULEA UR6, UR6, UR32, 0xf; UBLKCP.S.G [..], [UR6], ..) -- an
// illegal-address fault that came and went with unrelated changes to the translation unit (found with cuda-gdb).
// A 32-bit product has defined wrap-around, so it cannot be folded into a 64-bit shift-add.
PB_DEV void pb_t2_bulk(uint32_t dst, const void* base, uint32_t off, uint32_t bytes, uint32_t mbar) {
  asm volatile(
      "{\n\t.reg .u64 a;\n\tcvt.u64.u32 a, %2;\n\tadd.u64 a, a, %1;\n\t"
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [a], %3, [%4];\n\t}\n" ::"r"(dst),
      "l"(base), "r"(off), "r"(bytes), "r"(mbar)
      : "memory");
}
PB_DEV void pb_t2_fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
PB_DEV void pb_t2_st4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// mbarrier slots (8 bytes each) at off_t2_bar
enum {
  PB_T2_B_WFULL = 0,                    // [NS] ring stage filled (tx bytes)
  PB_T2_B_WEMPTY = PB_T2_B_WFULL + 3,   // [NS] ring stage consumed (tcgen05.commit)
  PB_T2_B_AXFULL = PB_T2_B_WEMPTY + 3,  // activation area filled by bulk copies (X chunk of fwd1)
  PB_T2_B_AFULL,                        // activation chunk written by the 256 workers
  PB_T2_B_AEMPTY,                       // activation chunk consumed (commit)
  PB_T2_B_ACC,                          // accumulator complete: Z1, Z2, E, D1 in this order (commit)
  PB_T2_B_D2FULL,                       // [2] D2 chunk slot complete (commit)
  PB_T2_B_D2EMPTY = PB_T2_B_D2FULL + 2, // [2] D2 chunk slot drained by the 256 workers
  PB_T2_B_XFULL = PB_T2_B_D2EMPTY + 2,  // X image resident in the ring area (dW1)
  PB_T2_B_S1FREE,                       // D1 drained by the 256 workers: TMEM S1 may take the D2 chunk slots
  PB_T2_B_ZFULL,                        // [NZ] noise-ring slot written by the 256 noise threads
  PB_T2_B_ZEMPTY = PB_T2_B_ZFULL + PB_T2_NZ,  // [NZ] noise-ring slot consumed by the 256 workers
  PB_T2_NBAR = PB_T2_B_ZEMPTY + PB_T2_NZ
};

struct PbTc2 {
  uint32_t tmem;   // TMEM base (512 columns)
  uint32_t big;    // shared address of the 1024-byte aligned operand area: activation | mask | ring
  uint8_t* img;    // this CTA's global image block
};

PB_DEV uint32_t pb_t2_bar(const PbPlan& pl, float* sm, int i) {
  return pb_smem_u32(reinterpret_cast<uint64_t*>(sm + pl.off_t2_bar) + i);
}

// byte offset of element (row r, column c) inside an image whose column blocks hold `rows` rows
PB_DEV uint32_t pb_t2_off(int rows, int r, int c) {
  return (uint32_t)((c >> 6) * rows * 128 + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + ((c & 7) << 1));
}

// once per CTA: TMEM, image block zeroed (pad rows / columns of the images must read as zeros)
PB_DEV void pb_tc2_setup(const PbPlan& pl, float* sm, uint8_t* img, int nt, PbTc2& tc) {
  const int tid = threadIdx.x;
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + pl.off_t2_bar) + 2 * PB_T2_NBAR;
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(pb_smem_u32(slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  uint4* z = reinterpret_cast<uint4*>(img);
  for (int i = tid; i < pl.t2_img_bytes / 16; i += nt) z[i] = make_uint4(0u, 0u, 0u, 0u);
  pb_tc_fence_before();
  __syncthreads();
  pb_tc_fence_after();
  tc.tmem = *slot;
  tc.img = img;
  const uint32_t raw = pb_smem_u32(sm + pl.off_t2_big);
  tc.big = (raw + 1023u) & ~1023u;
  __syncthreads();
}

PB_DEV void pb_tc2_teardown(const PbTc2& tc) {
  pb_tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    pb_tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tc.tmem));
  }
}

// minibatch rows -> bf16x3 image of Xaug = (x, 1, 0..) [128 rows][Kx] (rows >= nrows are zero) + y in shared memory
PB_DEV void pb_tc2_stage_x(const PbPlan& pl, float* sm, const PbTc2& tc, const float* X, const float* y,
                           const int* idxs, int nrows, int nt) {
  const int tid = threadIdx.x, d = pl.d, Kx = pl.t2_Kx;
  const int tbytes = pl.t2_ncbx * PB_T2_TILE;
  float* yv = sm + pl.off_t2_vec + 512;
  for (int e = tid; e < 128 * (Kx >> 1); e += nt) {
    const int b = e / (Kx >> 1), k0 = 2 * (e - b * (Kx >> 1));
    float v0 = 0.f, v1 = 0.f;
    if (b < nrows) {
      const size_t row = (size_t)(idxs ? idxs[b] : b);
      v0 = k0 < d ? X[row * d + k0] : (k0 == d ? 1.f : 0.f);
      v1 = k0 + 1 < d ? X[row * d + k0 + 1] : (k0 + 1 == d ? 1.f : 0.f);
    }
    uint32_t t1, t2, t3;
    pb_t2_split3(v0, v1, t1, t2, t3);
    uint8_t* p = tc.img + pb_t2_off(128, b, k0);
    *reinterpret_cast<uint32_t*>(p) = t1;
    *reinterpret_cast<uint32_t*>(p + tbytes) = t2;
    *reinterpret_cast<uint32_t*>(p + 2 * tbytes) = t3;
  }
  for (int b = tid; b < 128; b += nt) yv[b] = b < nrows ? y[(size_t)(idxs ? idxs[b] : b)] : 0.f;
  pb_t2_fence_async_global();
  PB_TAG(10)
  PB_ENDPHASE_RAW
}

// small vectors of one evaluation: b2, w3 (fp32) and the packed 3-term split of w3 (pairs), b3 -> shared memory
PB_DEV void pb_tc2_vectors(const PbPlan& pl, float* sm, const float* __restrict__ TH, int nt) {
  const int tid = threadIdx.x;
  const PbLayer& L1 = pl.L[1];
  const PbLayer& L2 = pl.L[2];
  const int H = pl.t2_H;
  float* b2v = sm + pl.off_t2_vec;
  float* w3v = b2v + 256;
  uint32_t* w3t = reinterpret_cast<uint32_t*>(b2v + 1412);  // [3][128] packed pairs
  for (int n = tid; n < 256; n += nt) {
    b2v[n] = n < H ? TH[L1.soff + H * L1.wp + n] : 0.f;
    w3v[n] = n < H ? TH[L2.soff + n * L2.wp] : 0.f;
  }
  for (int p = tid; p < 128; p += nt) {
    const float a0 = 2 * p < H ? TH[L2.soff + (2 * p) * L2.wp] : 0.f;
    const float a1 = 2 * p + 1 < H ? TH[L2.soff + (2 * p + 1) * L2.wp] : 0.f;
    uint32_t t1, t2, t3;
    pb_t2_split3(a0, a1, t1, t2, t3);
    w3t[p] = t1;
    w3t[128 + p] = t2;
    w3t[256 + p] = t3;
  }
  if (tid == 0) sm[pl.off_t2_vec + 1408] = TH[L2.soff + H * L2.wp];  // b3
}

// theta (padded resident layout, global) -> operand images of W1aug and W2 (start of a chain; afterwards the fused
// drains keep the images current)
PB_DEV void pb_tc2_restage(const PbPlan& pl, float* sm, const PbTc2& tc, const float* __restrict__ TH, int nt) {
  const int tid = threadIdx.x;
  const PbLayer& L0 = pl.L[0];
  const PbLayer& L1 = pl.L[1];
  const int H = pl.t2_H, Hp = pl.t2_Hp, Kx = pl.t2_Kx, d = pl.d;
  PB_TAG(1)
  {  // W1aug image: storage row = input index k (k == d: bias), columns = hidden unit
    const int tb = pl.t2_nch * Kx * 128, hp2 = Hp >> 1;
    uint8_t* base = tc.img + pl.t2_img_w1;
    const int total = (d + 1) * hp2;
    for (int e0 = tid; e0 < total; e0 += 4 * nt) {
      float v0[4], v1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * nt;
        const int k = e / hp2, n0 = 2 * (e - k * hp2);
        v0[u] = (e < total && n0 < H) ? TH[L0.soff + k * L0.wp + n0] : 0.f;
        v1[u] = (e < total && n0 + 1 < H) ? TH[L0.soff + k * L0.wp + n0 + 1] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * nt;
        if (e < total) {
          const int k = e / hp2, n0 = 2 * (e - k * hp2);
          uint32_t t1, t2, t3;
          pb_t2_split3(v0[u], v1[u], t1, t2, t3);
          uint8_t* p = base + pb_t2_off(Kx, k, n0);
          *reinterpret_cast<uint32_t*>(p) = t1;
          *reinterpret_cast<uint32_t*>(p + tb) = t2;
          *reinterpret_cast<uint32_t*>(p + 2 * tb) = t3;
        }
      }
    }
  }
  {  // W2 image, TRANSPOSED: storage row = output unit n, columns = input index i (pairs i, i + 1 packed): the drain
     // thread of output unit n (its TMEM lane) then owns one image row and moves 8 weights per 16-byte access
    const int tb = pl.t2_nch * Hp * 128, hp2 = Hp >> 1;
    uint8_t* base = tc.img + pl.t2_img_w2;
    const int total = H * hp2;  // (n, pair of i)
    for (int e0 = tid; e0 < total; e0 += 4 * nt) {
      float v0[4], v1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * nt;
        const int ip = e / H, n = e - ip * H, i0 = 2 * ip;  // consecutive threads: consecutive n (coalesced TH reads)
        v0[u] = (e < total && i0 < H) ? TH[L1.soff + i0 * L1.wp + n] : 0.f;
        v1[u] = (e < total && i0 + 1 < H) ? TH[L1.soff + (i0 + 1) * L1.wp + n] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * nt;
        if (e < total) {
          const int ip = e / H, n = e - ip * H, i0 = 2 * ip;
          uint32_t t1, t2, t3;
          pb_t2_split3(v0[u], v1[u], t1, t2, t3);
          uint8_t* p = base + pb_t2_off(Hp, n, i0);
          *reinterpret_cast<uint32_t*>(p) = t1;
          *reinterpret_cast<uint32_t*>(p + tb) = t2;
          *reinterpret_cast<uint32_t*>(p + 2 * tb) = t3;
        }
      }
    }
  }
  pb_t2_fence_async_global();
  PB_ENDPHASE_RAW
}

// new value of one weight -> its three bf16 terms in an operand image (term images `tb` bytes apart)
PB_DEV void pb_t2_put(uint8_t* p, int tb, float v) {
  uint32_t t1, t2, t3;
  pb_t2_split3(v, 0.f, t1, t2, t3);
  *reinterpret_cast<uint16_t*>(p) = (uint16_t)t1;
  *reinterpret_cast<uint16_t*>(p + tb) = (uint16_t)t2;
  *reinterpret_cast<uint16_t*>(p + 2 * tb) = (uint16_t)t3;
}

// the weight itself from its three bf16 terms (exact: the split is error-free, and so are the two additions); .cg:
// the image is written with plain stores by this CTA and read through L2 by the bulk copies -- keep L1 out of it
PB_DEV float pb_t2_get(const uint8_t* p, int tb) {
  const uint32_t t1 = __ldcg(reinterpret_cast<const unsigned short*>(p));
  const uint32_t t2 = __ldcg(reinterpret_cast<const unsigned short*>(p + tb));
  const uint32_t t3 = __ldcg(reinterpret_cast<const unsigned short*>(p + 2 * tb));
  return (__uint_as_float(t1 << 16) + __uint_as_float(t2 << 16)) + __uint_as_float(t3 << 16);
}

// ---- one gradient evaluation with the parameter update fused into the drains ------------------------------------
//   pre = pf(resident_index, flat_index)          loads what the update needs BESIDES theta (sampler state, control
//                                                 variates); issued for a group of elements before any of them is
//                                                 stored.  The engine fills pre.a = theta itself: the operand IMAGES are
//                                                 the master copy of W1, b1 and W2 (three bf16 terms = the fp32 value,
//                                                 exactly) -- no fp32 copy of them is read or written per step, which
//                                                 keeps a 90-256-256-1 chain's working set at ~640 KB (148 chains fit the
//                                                 L2); b2, w3, b3 stay fp32 in TH
//   out0 / out1                                   optional flat [P] destinations of the NEW theta (trace row / final
//                                                 state; out1 also raises bit 0 of *flagword on a non-finite value)
//   theta_new = uf(resident_index, flat_index, pre, gll, z)   gll = d loglik / d theta of the element, z =
//                                                 jax.random.normal(nkey, (P,))[flat_index] when `noise` (else 0); the
//                                                 functor owns the sampler arithmetic (prior term, state arrays) and
//                                                 returns the value theta is set to (pre.a leaves it unchanged)
// The drains work on groups of 8 elements: 8 x pf, then the 8 threefry + erf_inv chains (branch-free, interleaved by the
// compiler), then 8 x uf, then the stores.  llbuf[0..127] <- per-row log-likelihood terms.  TH is updated in place.
template <class PF, class UF>
PB_DEV void pb_grad_tc2(const PbPlan& pl, float* sm, float* TH, int nrows, float scale, int nt, const PbTc2& tc,
                        PbKey nkey, bool noise, float* out0, float* out1, uint32_t* flagword, PF pf, UF uf) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const PbLayer& L0 = pl.L[0];
  const PbLayer& L1 = pl.L[1];
  const PbLayer& L2 = pl.L[2];
  const int H = pl.t2_H, Hp = pl.t2_Hp, Hy = pl.t2_Hy, Kx = pl.t2_Kx;
  const int nch = pl.t2_nch, nchy = pl.t2_nchy, ncbx = pl.t2_ncbx, nmt = pl.t2_nmt;
  const uint32_t acta = tc.big, m2 = tc.big + 3u * PB_T2_TILE, ring = m2 + (uint32_t)nch * PB_T2_TILE;
  const uint32_t wstage = (uint32_t)pl.t2_wstage;
  const int NZr = pl.t2_nz;
  const uint32_t S0 = tc.tmem, S1 = tc.tmem + 256u;
  float* b2v = sm + pl.off_t2_vec;
  float* w3v = b2v + 256;
  float* yv = b2v + 512;
  float* fp = b2v + 640;    // [2][128] partial outputs of the two column halves
  float* dw3p = b2v + 896;  // [2][256] partial output-weight gradients
  const uint32_t* w3t = reinterpret_cast<const uint32_t*>(b2v + 1412);
  float* llbuf = pb_llbuf(pl, sm);

  PB_TAG(1)
  pb_tc2_vectors(pl, sm, TH, nt);
  if (tid == 0) {
    for (int i = 0; i < PB_T2_NBAR; ++i) {
      const uint32_t b = pb_t2_bar(pl, sm, i);
      uint32_t cnt = (i == PB_T2_B_AFULL || i == PB_T2_B_D2EMPTY || i == PB_T2_B_D2EMPTY + 1 ||
                      i == PB_T2_B_S1FREE || i >= PB_T2_B_ZFULL)
                         ? PB_T2_WORKERS
                         : 1u;
      if (pl.t2_help && (i == PB_T2_B_D2EMPTY || i == PB_T2_B_D2EMPTY + 1)) cnt = 2u * PB_T2_WORKERS;  // + helpers
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(cnt));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  pb_t2_fence_async_global();  // image stores of the previous evaluation's drains -> visible to the bulk copies
  pb_tc_fence_before();
  PB_ENDPHASE_RAW
  pb_tc_fence_after();
  const uint32_t bar0 = pb_t2_bar(pl, sm, 0);
#define PB_T2_BAR(i) (bar0 + 8u * (uint32_t)(i))

  bool bad = false;
  auto emit = [&](int fi, float v) {
    if (out0) out0[fi] = v;
    if (out1) {
      out1[fi] = v;
      bad |= !pb_finite(v);
    }
  };
  // ---- one block of 8 consecutive weights of ONE image row (a 16-byte chunk per bf16 term) ----------------------
  //   p: address of the chunk in the first term image, terms tb bytes apart;  element j: resident index si0 + j sis,
  //   flat index fi0 + j fis;  nval (1..8) valid elements (the rest are pad columns and stay zero);  dv: the eight
  //   accumulator values, scaled by gs into d loglik / d theta;  acc += sum_j theta_j dv_j (output-layer gradient)
  auto drain8 = [&](uint8_t* p, int tb, int si0, int sis, int fi0, int fis, int nval, const float* dv, float gs,
                    float& acc, const float* zp) {
    const uint4 q1 = __ldcg(reinterpret_cast<const uint4*>(p));
    const uint4 q2 = __ldcg(reinterpret_cast<const uint4*>(p + tb));
    const uint4 q3 = __ldcg(reinterpret_cast<const uint4*>(p + 2 * tb));
    const uint32_t w1[4] = {q1.x, q1.y, q1.z, q1.w}, w2[4] = {q2.x, q2.y, q2.z, q2.w}, w3[4] = {q3.x, q3.y, q3.z, q3.w};
    float nv[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // two groups of four: the four threefry + erf_inv chains of a group are interleaved
      PbPre pre[4];
      float zz[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = 4 * h + j, m = e >> 1;
        if (e < nval) {
          pre[j] = pf(si0 + e * sis, fi0 + e * fis);
          pre[j].a = (e & 1) ? (__uint_as_float(w1[m] & 0xFFFF0000u) + __uint_as_float(w2[m] & 0xFFFF0000u)) +
                                   __uint_as_float(w3[m] & 0xFFFF0000u)
                             : (__uint_as_float(w1[m] << 16) + __uint_as_float(w2[m] << 16)) +
                                   __uint_as_float(w3[m] << 16);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)  // the block's normals: written by the mirror noise thread (slot layout [j][thread])
        zz[j] = (zp && 4 * h + j < nval) ? __ldcg(zp + (4 * h + j) * PB_T2_WORKERS) : 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = 4 * h + j;
        nv[e] = 0.f;
        if (e < nval) {
          acc = fmaf(pre[j].a, dv[e], acc);
          nv[e] = uf(si0 + e * sis, fi0 + e * fis, pre[j], gs * dv[e], zz[j]);
          emit(fi0 + e * fis, nv[e]);
        }
      }
    }
    uint32_t t1[4], t2[4], t3[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) pb_t2_split3(nv[2 * m], nv[2 * m + 1], t1[m], t2[m], t3[m]);
    *reinterpret_cast<uint4*>(p) = make_uint4(t1[0], t1[1], t1[2], t1[3]);
    *reinterpret_cast<uint4*>(p + tb) = make_uint4(t2[0], t2[1], t2[2], t2[3]);
    *reinterpret_cast<uint4*>(p + 2 * tb) = make_uint4(t3[0], t3[1], t3[2], t3[3]);
  };
  // ---- one 8-column step of a D2 chunk slot: TMEM lane n_own (= row n_own of the transposed W2 image), columns c0 .. of
  //      chunk ic: fused update of W2[i0 .. i0 + 7][n_own] and, in the block that holds column H, of the bias b2[n_own] ----
  const bool two_tiles = nmt > 1;
  const bool help = pl.t2_help != 0;  // the whole evaluation's noise fits the ring: the noise warps then drain every other
                                      // D2 step (the drains of small nets are latency-bound, not issue-bound)
  const int tb2 = nch * Hp * 128;
  uint8_t* img2 = tc.img + pl.t2_img_w2;
  auto d2_step = [&](int ic, int slot, int c0, int gq, uint32_t lane_off, int n_own, float w3n, const float* zp,
                     float& dw3) {
    uint32_t dr[8];
    pb_tmem_ld8(S1 + lane_off + (uint32_t)(slot * 128 + (two_tiles ? 64 * gq : 0) + c0), dr);
    pb_tmem_wait_ld();
    const int i0 = 64 * ic + c0;  // input index of W2 (i == H: the bias row b2)
    if (n_own < H && i0 <= H) {
      float dv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) dv[j] = __uint_as_float(dr[j]);
      if (i0 < H)  // kernel rows i0 .. : columns i0 .. of this thread's row of the transposed image
        drain8(img2 + pb_t2_off(Hp, n_own, i0), tb2, L1.soff + i0 * L1.wp + n_own, L1.wp, L1.fw + i0 * H + n_own, H,
               H - i0 < 8 ? H - i0 : 8, dv, w3n, dw3, zp);
      if (H - i0 < 8) {  // this block holds the bias column (i == H): b2 stays fp32 in TH
        float dd = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) dd = (i0 + j == H) ? dv[j] : dd;
        const int si = L1.soff + H * L1.wp + n_own, fi = L1.fb + n_own;
        PbPre pre = pf(si, fi);
        pre.a = TH[si];
        dw3 = fmaf(pre.a, dd, dw3);
        const float nvb = uf(si, fi, pre, w3n * dd, noise ? pb_normal_at(nkey, (uint32_t)fi) : 0.f);
        TH[si] = nvb;
        emit(fi, nvb);
      }
    }
  };
  float* dw3h = b2v + 1412;  // [256] the helpers' share of the output-layer gradient (the w3 split there is dead by then)

  if (warp_u == 8) {
    // =============================== producer: bulk copies ===============================
    if (pb_elect_one()) {
      int fill = 0, ause = 0;  // ring fills / activation-area uses so far
      auto stage_begin = [&](uint32_t bytes) -> uint32_t {
        const int s = fill % PB_T2_NS;
        if (fill >= PB_T2_NS) pb_mbar_wait(PB_T2_BAR(PB_T2_B_WEMPTY + s), (uint32_t)((fill / PB_T2_NS - 1) & 1));
        pb_t2_expect(PB_T2_BAR(PB_T2_B_WFULL + s), bytes);
        return ring + (uint32_t)s * wstage;
      };
      // row chunks of an MN-major B image: rows [64 c, 64 c + rows) of every column block, one term per stage
      auto row_chunk = [&](const uint8_t* img, int img_rows, int c, int rows) {
        for (int t = 0; t < 3; ++t) {
          const uint32_t dst = stage_begin((uint32_t)(nch * rows * 128));
          const int s = fill % PB_T2_NS;
          for (int cb = 0; cb < nch; ++cb)
            pb_t2_bulk(dst + (uint32_t)cb * 8192u, img, (uint32_t)((t * nch + cb) * img_rows + 64 * c) * 128u,
                       (uint32_t)(rows * 128), PB_T2_BAR(PB_T2_B_WFULL + s));
          ++fill;
        }
      };
      // G1: X chunk kc (3 terms) -> activation area; W1aug rows of chunk kc
      for (int kc = 0; kc < ncbx; ++kc) {
        const int rows = (Kx - 64 * kc) < 64 ? (Kx - 64 * kc) : 64;
        if (ause > 0) pb_mbar_wait(PB_T2_BAR(PB_T2_B_AEMPTY), (uint32_t)((ause - 1) & 1));
        pb_t2_expect(PB_T2_BAR(PB_T2_B_AXFULL), 3u * PB_T2_TILE);
        for (int t = 0; t < 3; ++t)
          pb_t2_bulk(acta + (uint32_t)t * PB_T2_TILE, tc.img, (uint32_t)(t * ncbx + kc) * PB_T2_TILE, PB_T2_TILE,
                     PB_T2_BAR(PB_T2_B_AXFULL));
        ++ause;
        row_chunk(tc.img + pl.t2_img_w1, Kx, kc, rows);
      }
      // G2 (K = input index i = image COLUMNS): column block ic of the transposed W2 image, all Hp rows (K-major B)
      for (int kc = 0; kc < nch; ++kc)
        for (int t = 0; t < 3; ++t) {
          const uint32_t bytes = (uint32_t)(Hp * 128);
          const uint32_t dst = stage_begin(bytes);
          const int s = fill % PB_T2_NS;
          const uint32_t so = (uint32_t)pl.t2_img_w2 + (uint32_t)((t * nch + kc) * Hp) * 128u;
          for (uint32_t o = 0; o < bytes; o += 16384u)
            pb_t2_bulk(dst + o, tc.img, so + o, bytes - o < 16384u ? bytes - o : 16384u, PB_T2_BAR(PB_T2_B_WFULL + s));
          ++fill;
        }
      // G5 (K = output unit n = image ROWS): rows of chunk kc, every column block (MN-major B)
      for (int ic = 0; ic < nch; ++ic) row_chunk(tc.img + pl.t2_img_w2, Hp, ic, (Hp - 64 * ic) < 64 ? (Hp - 64 * ic) : 64);
      // G6: the whole X image into the ring area once every stage has been consumed
      for (int j = fill > PB_T2_NS ? fill - PB_T2_NS : 0; j < fill; ++j)
        pb_mbar_wait(PB_T2_BAR(PB_T2_B_WEMPTY + j % PB_T2_NS), (uint32_t)((j / PB_T2_NS) & 1));
      const uint32_t xbytes = 3u * (uint32_t)ncbx * PB_T2_TILE;
      pb_t2_expect(PB_T2_BAR(PB_T2_B_XFULL), xbytes);
      for (uint32_t o = 0; o < xbytes; o += PB_T2_TILE)
        pb_t2_bulk(ring + o, tc.img, o, PB_T2_TILE, PB_T2_BAR(PB_T2_B_XFULL));
    }
    __syncwarp();
  } else if (warp_u == 9) {
    // =============================== MMA issue ===============================
    int fill = 0;
    uint32_t axph = 0u, afph = 0u;
    auto stage_wait = [&]() -> uint32_t {
      const int s = fill % PB_T2_NS;
      pb_mbar_wait(PB_T2_BAR(PB_T2_B_WFULL + s), (uint32_t)((fill / PB_T2_NS) & 1));
      pb_tc_fence_after();
      return ring + (uint32_t)s * wstage;
    };
    auto stage_done = [&]() {
      if (pb_elect_one()) pb_umma_commit(PB_T2_BAR(PB_T2_B_WEMPTY + fill % PB_T2_NS));
      __syncwarp();
      ++fill;
    };
    auto commit = [&](int b) {
      if (pb_elect_one()) pb_umma_commit(PB_T2_BAR(b));
      __syncwarp();
    };
    // GEMM over 64-wide K chunks: A chunk (3 terms, K-major) in the activation area, one term of B per ring stage
    // (b_kmajor: B rows = output index, else B rows = K index of the chunk); 6 passes (term pairs i + j <= 4)
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
          const uint32_t st = stage_wait();
          if (pb_elect_one()) {
            for (int xa = 0; xa + wt < 3; ++xa)
              for (int ks = 0; ks < ksteps; ++ks)
                pb_t2_mma(dcol, pb_t2_desc(acta + (uint32_t)xa * PB_T2_TILE + (uint32_t)ks * 32u, 16u, 1024u),
                          b_kmajor ? pb_t2_desc(st + (uint32_t)ks * 32u, 16u, 1024u)
                                   : pb_t2_desc(st + (uint32_t)ks * 2048u, 8192u, 1024u),
                          idesc, (c | wt | xa | ks) ? 1u : 0u);
          }
          __syncwarp();
          stage_done();
        }
        commit(PB_T2_B_AEMPTY);
      }
      commit(PB_T2_B_ACC);
    };
    kgemm(S0, ncbx, Kx, true, false);   // G1: Z1
    kgemm(S1, nch, Hp, false, true);    // G2: Z2 (B = transposed W2 image read K-major)
    kgemm(S1, nch, Hp, false, false);   // G5: E = (w3 * M2) W2^T  (the same image read MN-major; S1: Z2 has been
                                        // consumed; Z1 stays in S0 for G4)
    // G6: D1 chunks.  A = X^T (MN-major, lanes = input index k), B = dZ1 chunk (MN-major), K = 128 batch rows
    pb_mbar_wait(PB_T2_BAR(PB_T2_B_XFULL), 0u);
    pb_tc_fence_after();
    for (int ic = 0; ic < nch; ++ic) {
      const int ncols = (Hp - 64 * ic) < 64 ? (Hp - 64 * ic) : 64;
      pb_mbar_wait(PB_T2_BAR(PB_T2_B_AFULL), afph);
      afph ^= 1u;
      pb_tc_fence_after();
      if (pb_elect_one()) {
        const uint32_t idesc = pb_t2_idesc(128, ncols, 1, 1);
        for (int zt = 0; zt < 3; ++zt)
          for (int xa = 0; xa + zt < 3; ++xa)
            for (int ks = 0; ks < 8; ++ks)
              pb_t2_mma(S1 + (uint32_t)ic * 64u,
                        pb_t2_desc(ring + (uint32_t)(xa * ncbx) * PB_T2_TILE + (uint32_t)ks * 2048u,
                                   ncbx > 1 ? PB_T2_TILE : 0u, 1024u),
                        pb_t2_desc(acta + (uint32_t)zt * PB_T2_TILE + (uint32_t)ks * 2048u, 0u, 1024u), idesc,
                        (zt | xa | ks) ? 1u : 0u);
      }
      __syncwarp();
      commit(PB_T2_B_AEMPTY);
    }
    commit(PB_T2_B_ACC);
    // G4 (last: the fused drains overwrite the W2 image, which fwd2 and dH1 must have read): D2 chunks.  A = M2^T tile (MN-major, 128 n), B = Y chunk (MN-major), K = 128 batch rows
    pb_mbar_wait(PB_T2_BAR(PB_T2_B_S1FREE), 0u);
    pb_tc_fence_after();
    for (int ic = 0; ic < nchy; ++ic) {
      const int ncols = (Hy - 64 * ic) < 64 ? (Hy - 64 * ic) : 64;
      const int slot = ic & 1;
      pb_mbar_wait(PB_T2_BAR(PB_T2_B_AFULL), afph);
      afph ^= 1u;
      if (ic >= 2) pb_mbar_wait(PB_T2_BAR(PB_T2_B_D2EMPTY + slot), (uint32_t)(((ic >> 1) - 1) & 1));
      pb_tc_fence_after();
      if (pb_elect_one()) {
        const uint32_t idesc = pb_t2_idesc(128, ncols, 1, 1);
        for (int mt = 0; mt < nmt; ++mt)
          for (int t = 0; t < 3; ++t)
            for (int ks = 0; ks < 8; ++ks)
              pb_t2_mma(S1 + (uint32_t)slot * 128u + (uint32_t)mt * 64u,
                        pb_t2_desc(m2 + (uint32_t)(2 * mt) * PB_T2_TILE + (uint32_t)ks * 2048u,
                                   nch > 2 * mt + 1 ? PB_T2_TILE : 0u, 1024u),
                        pb_t2_desc(acta + (uint32_t)t * PB_T2_TILE + (uint32_t)ks * 2048u, 0u, 1024u), idesc,
                        (t | ks) ? 1u : 0u);
      }
      __syncwarp();
      commit(PB_T2_B_AEMPTY);
      commit(PB_T2_B_D2FULL + slot);
    }
    // Tail: the commits nobody else waits for (last ring stages, last activation chunk) must have ARRIVED before the next
    // evaluation re-initialises the barriers or the CTA exits.
    for (int i = fill > PB_T2_NS ? fill - PB_T2_NS : 0; i < fill; ++i)
      pb_mbar_wait(PB_T2_BAR(PB_T2_B_WEMPTY + i % PB_T2_NS), (uint32_t)((i / PB_T2_NS) & 1));
    pb_mbar_wait(PB_T2_BAR(PB_T2_B_AEMPTY), (uint32_t)((ncbx + 3 * nch + nchy - 1) & 1));
  } else if (warp_u >= 10) {
    // =============================== noise warps ===============================
    // thread wt mirrors worker thread wt: it walks the SAME sequence of 8-weight blocks (first-layer items, then the
    // D2 chunks) and writes the block's eight normals to ring slot (block number) % NZ.  The draws depend on nothing
    // but (key, flat index), so these warps start with the evaluation and run up to NZ blocks ahead of the drains.
    // (a warp reaches the TMEM lanes 32 (warp % 4) ..: the mirror of noise warp pw is the worker with the same quadrant)
    const int q = warp & 3, g = (warp - 10) >> 2, b = 32 * q + lane, wt = 128 * g + b;
    const int n_own = two_tiles ? wt : b;
    float* zr = reinterpret_cast<float*>(tc.img + pl.t2_img_z) + wt;
    if (noise) {
      int cnt = 0;
      auto put = [&](int fi0, int fis, bool any) {
        const int slot = cnt % NZr;
        if (cnt >= NZr) pb_mbar_wait(PB_T2_BAR(PB_T2_B_ZEMPTY + slot), (uint32_t)((cnt / NZr - 1) & 1));
        if (any) {
          uint32_t ix[8];
          float z[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) ix[j] = (uint32_t)(fi0 + j * fis);
          pb_normal_n<8, true>(nkey, ix, z);
#pragma unroll
          for (int j = 0; j < 8; ++j) zr[slot * PB_T2_ZSLOT + j * PB_T2_WORKERS] = z[j];
        }
        pb_t2_arrive(PB_T2_BAR(PB_T2_B_ZFULL + slot));
        ++cnt;
      };
      {
        const int nblk = (H + 7) >> 3, items = (pl.d + 1) * nblk;
        for (int it = wt; it - wt < items; it += PB_T2_WORKERS) {
          const int k = it / nblk, i0 = 8 * (it - k * nblk);
          put((k < pl.d ? L0.fw + k * H : L0.fb) + i0, 1, it < items);
        }
      }
      for (int ic = 0; ic < nchy; ++ic) {
        const int ncols = (Hy - 64 * ic) < 64 ? (Hy - 64 * ic) : 64;
        const int nsteps = two_tiles ? (ncols >> 3) : ((ncols >> 3) < 4 ? (ncols >> 3) : 4);
        for (int s = 0; s < nsteps; ++s) {
          const int c0 = two_tiles ? 8 * s : 32 * g + 8 * s, i0 = 64 * ic + c0;
          put(L1.fw + i0 * H + n_own, H, n_own < H && c0 < ncols && i0 < H);
        }
      }
    }
    if (help) {
      // ---- helper drains: the odd 8-column steps of every D2 chunk (the workers take the even ones) ----
      const uint32_t lane_off = (uint32_t)(32 * q) << 16;
      const float w3n = n_own < H ? w3v[n_own] : 0.f;
      float dw3 = 0.f;
      int zc = ((pl.d + 1) * ((H + 7) >> 3) + PB_T2_WORKERS - 1) / PB_T2_WORKERS;  // ring slots of the first-layer items
      for (int ic = 0; ic < nchy; ++ic) {
        const int slot = ic & 1;
        const int ncols = (Hy - 64 * ic) < 64 ? (Hy - 64 * ic) : 64;
        const int nsteps = two_tiles ? (ncols >> 3) : ((ncols >> 3) < 4 ? (ncols >> 3) : 4);
        pb_mbar_wait(PB_T2_BAR(PB_T2_B_D2FULL + slot), (uint32_t)((ic >> 1) & 1));
        pb_tc_fence_after();
        for (int s = 0; s < nsteps; ++s, ++zc) {
          const int c0 = two_tiles ? 8 * s : 32 * g + 8 * s;
          if ((s & 1) && c0 < ncols)  // (this thread wrote slot zc itself; the ring never wraps in this mode)
            d2_step(ic, slot, c0, g, lane_off, n_own, w3n, noise ? zr + (zc % NZr) * PB_T2_ZSLOT : nullptr, dw3);
        }
        pb_tc_fence_before();
        pb_t2_arrive(PB_T2_BAR(PB_T2_B_D2EMPTY + slot));
      }
      dw3h[wt] = dw3;
      if (bad && flagword) atomicOr(flagword, 1u);
      asm volatile("bar.sync 2, 512;" ::: "memory");
    }
    __syncwarp();
  } else {
    // =============================== workers ===============================
    const int q = warp & 3, g = warp >> 2;
    const int b = 32 * q + lane;                       // batch row == TMEM lane in the forward accumulators
    const uint32_t lane_off = (uint32_t)(32 * q) << 16;
    const uint32_t rowaddr = (uint32_t)b * 128u;
    const uint32_t bx = (uint32_t)(b & 7);
    int ause = ncbx;                                   // activation-area uses before the first worker chunk
    uint32_t accph = 0u;
    uint32_t m1[4] = {0u, 0u, 0u, 0u};                 // relu masks: this thread's 32 columns of each chunk
    uint32_t m2b[4] = {0u, 0u, 0u, 0u};
    // three packed terms of 32 values -> this thread's 4 chunks (16 B each) of the activation rows
    auto store_chunk = [&](const float* v) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t t1[4], t2[4], t3[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pb_t2_split3(v[8 * u + 2 * j], v[8 * u + 2 * j + 1], t1[j], t2[j], t3[j]);
        const uint32_t a = acta + rowaddr + ((((uint32_t)(4 * g + u)) ^ bx) << 4);
        pb_t2_st4(a, t1[0], t1[1], t1[2], t1[3]);
        pb_t2_st4(a + PB_T2_TILE, t2[0], t2[1], t2[2], t2[3]);
        pb_t2_st4(a + 2u * PB_T2_TILE, t3[0], t3[1], t3[2], t3[3]);
      }
    };
    auto chunk_begin = [&]() {  // the previous user of the activation area has been consumed by the tensor core
      pb_mbar_wait(PB_T2_BAR(PB_T2_B_AEMPTY), (uint32_t)((ause - 1) & 1));
    };
    auto chunk_end = [&]() {
      pb_fence_async_smem();
      pb_tc_fence_before();
      pb_t2_arrive(PB_T2_BAR(PB_T2_B_AFULL));
      ++ause;
    };
    auto acc_wait = [&]() {
      pb_mbar_wait(PB_T2_BAR(PB_T2_B_ACC), accph);
      accph ^= 1u;
      pb_tc_fence_after();
    };

    // ---- G2: H1 chunks = relu(Z1) ----
    acc_wait();  // Z1
    PB_MARK(2)
    for (int ic = 0; ic < nch; ++ic) {
      uint32_t z[32];
      float v[32];
      pb_tmem_ld32(S0 + lane_off + (uint32_t)(64 * ic + 32 * g), z);
      pb_tmem_wait_ld();
      uint32_t mb = 0u;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float zz = __uint_as_float(z[j]);
        mb |= zz > 0.f ? (1u << j) : 0u;
        v[j] = fmaxf(zz, 0.f);
      }
      if (ic == 0) m1[0] = mb; else if (ic == 1) m1[1] = mb; else if (ic == 2) m1[2] = mb; else m1[3] = mb;
      chunk_begin();
      store_chunk(v);
      chunk_end();
    }
    PB_MARK(3)
    // ---- G3: output, residual, layer-2 mask ----
    acc_wait();  // Z2
    PB_MARK(4)
    float fpart = 0.f;
    for (int ic = 0; ic < nch; ++ic) {
      uint32_t z[32];
      pb_tmem_ld32(S1 + lane_off + (uint32_t)(64 * ic + 32 * g), z);
      pb_tmem_wait_ld();
      const int c0 = 64 * ic + 32 * g;
      uint32_t mb = 0u;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t mk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = 8 * u + 2 * j;
          const float h0 = __uint_as_float(z[c]) + b2v[c0 + c], h1 = __uint_as_float(z[c + 1]) + b2v[c0 + c + 1];
          fpart = fmaf(fmaxf(h0, 0.f), w3v[c0 + c], fpart);
          fpart = fmaf(fmaxf(h1, 0.f), w3v[c0 + c + 1], fpart);
          mk[j] = (h0 > 0.f ? 0x3F80u : 0u) | (h1 > 0.f ? 0x3F800000u : 0u);
          mb |= (h0 > 0.f ? (1u << c) : 0u) | (h1 > 0.f ? (2u << c) : 0u);
        }
        pb_t2_st4(m2 + (uint32_t)ic * PB_T2_TILE + rowaddr + ((((uint32_t)(4 * g + u)) ^ bx) << 4), mk[0], mk[1], mk[2],
                  mk[3]);
      }
      if (ic == 0) m2b[0] = mb; else if (ic == 1) m2b[1] = mb; else if (ic == 2) m2b[2] = mb; else m2b[3] = mb;
    }
    pb_tc_fence_before();
    fp[g * 128 + b] = fpart;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    float r;
    {
      const float f = (fp[b] + fp[128 + b]) + sm[pl.off_t2_vec + 1408];
      const bool valid = b < nrows;
      const float res = yv[b] - f;
      r = valid ? res * scale : 0.f;
      if (g == 0) llbuf[b] = valid ? -0.5f * res * res * pl.inv_s2 : 0.f;
    }
    PB_MARK(5)
    // ---- G5: chunks of w3 * M2 (terms selected from the split of w3) ----
    for (int kc = 0; kc < nch; ++kc) {
      const uint32_t mb = kc == 0 ? m2b[0] : kc == 1 ? m2b[1] : kc == 2 ? m2b[2] : m2b[3];
      const int p0 = (64 * kc + 32 * g) >> 1;
      chunk_begin();
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t t1[4], t2[4], t3[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = 8 * u + 2 * j;
          const uint32_t mw = ((mb >> c) & 1u ? 0xFFFFu : 0u) | ((mb >> (c + 1)) & 1u ? 0xFFFF0000u : 0u);
          t1[j] = w3t[p0 + 4 * u + j] & mw;
          t2[j] = w3t[128 + p0 + 4 * u + j] & mw;
          t3[j] = w3t[256 + p0 + 4 * u + j] & mw;
        }
        const uint32_t a = acta + rowaddr + ((((uint32_t)(4 * g + u)) ^ bx) << 4);
        pb_t2_st4(a, t1[0], t1[1], t1[2], t1[3]);
        pb_t2_st4(a + PB_T2_TILE, t2[0], t2[1], t2[2], t2[3]);
        pb_t2_st4(a + 2u * PB_T2_TILE, t3[0], t3[1], t3[2], t3[3]);
      }
      chunk_end();
    }
    // ---- G6: dZ1 chunks = r * E * [Z1 > 0] ----
    acc_wait();  // E
    PB_MARK(6)
    for (int ic = 0; ic < nch; ++ic) {
      uint32_t z[32];
      float v[32];
      pb_tmem_ld32(S1 + lane_off + (uint32_t)(64 * ic + 32 * g), z);
      pb_tmem_wait_ld();
      const uint32_t mb = ic == 0 ? m1[0] : ic == 1 ? m1[1] : ic == 2 ? m1[2] : m1[3];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (mb >> j) & 1u ? r * __uint_as_float(z[j]) : 0.f;
      chunk_begin();
      store_chunk(v);
      chunk_end();
    }
    PB_MARK(7)
    // ---- G7: D1 drain with the fused update of W1, b1: TMEM lane = input index k (k == d: bias row), group g takes
    //      its half of every chunk ----
    acc_wait();  // D1
    PB_MARK(8)
    // D1 (TMEM lane = input index k, k == d: bias row) -> shared memory [k][Hp + 4] in the ring area (the X image
    // there has been consumed): the update below then runs with one COLUMN per thread, coalesced and balanced
    float* d1buf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sm + pl.off_t2_big) +
                                            ((ring - pb_smem_u32(sm + pl.off_t2_big))));
    const int dp = Hp + 4;
    {
      const int k = b;
      for (int ic = 0; ic < nch; ++ic) {
        const int ncols = (Hp - 64 * ic) < 64 ? (Hp - 64 * ic) : 64;
        for (int c0 = 32 * g; c0 < ncols && c0 < 32 * g + 32; c0 += 8) {
          uint32_t dv[8];
          pb_tmem_ld8(S1 + lane_off + (uint32_t)(64 * ic + c0), dv);
          pb_tmem_wait_ld();
          if (k <= pl.d) {
            float4* dst = reinterpret_cast<float4*>(d1buf + k * dp + 64 * ic + c0);
            dst[0] = make_float4(__uint_as_float(dv[0]), __uint_as_float(dv[1]), __uint_as_float(dv[2]),
                                 __uint_as_float(dv[3]));
            dst[1] = make_float4(__uint_as_float(dv[4]), __uint_as_float(dv[5]), __uint_as_float(dv[6]),
                                 __uint_as_float(dv[7]));
          }
        }
      }
    }
    pb_tc_fence_before();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // noise ring: block number zcnt of this evaluation sits in slot zcnt % NZ (the noise warps walk the same sequence)
    const int wt = warp * 32 + lane;
    const float* zr = reinterpret_cast<const float*>(tc.img + pl.t2_img_z) + wt;
    int zcnt = 0;
    auto zwait = [&]() -> const float* {
      if (!noise) return nullptr;
      const int slot = zcnt % NZr;
      pb_mbar_wait(PB_T2_BAR(PB_T2_B_ZFULL + slot), (uint32_t)((zcnt / NZr) & 1));
      return zr + slot * PB_T2_ZSLOT;
    };
    auto zdone = [&]() {
      if (noise && !help) pb_t2_arrive(PB_T2_BAR(PB_T2_B_ZEMPTY + zcnt % NZr));  // (help: the ring never wraps)
      ++zcnt;
    };
    // first layer: (row k, 8-column block) items of the dumped D1, dealt round-robin to the 256 workers
    auto update_w1 = [&]() {
      const int d = pl.d;
      const int tb1 = nch * Kx * 128;
      uint8_t* img1 = tc.img + pl.t2_img_w1;
      const int nblk = (H + 7) >> 3, items = (d + 1) * nblk;
      float unused = 0.f;
      for (int it = wt; it - wt < items; it += PB_T2_WORKERS) {
        const float* zp = zwait();
        if (it < items) {
          const int k = it / nblk, i0 = 8 * (it - k * nblk);
          const float4 a0 = *reinterpret_cast<const float4*>(d1buf + k * dp + i0);
          const float4 a1 = *reinterpret_cast<const float4*>(d1buf + k * dp + i0 + 4);
          const float dv[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          drain8(img1 + pb_t2_off(Kx, k, i0), tb1, L0.soff + k * L0.wp + i0, 1, (k < d ? L0.fw + k * H : L0.fb) + i0, 1,
                 H - i0 < 8 ? H - i0 : 8, dv, 1.0f, unused, zp);
        }
        zdone();
      }
    };
    pb_tc_fence_before();
    pb_t2_arrive(PB_T2_BAR(PB_T2_B_S1FREE));
    PB_MARK(9)
    // ---- G4: Y chunks = r * [relu(Z1) | 1 | 0..]; the D2 drains (fused update of W2, b2) run one chunk behind ----
    const int n_own = two_tiles ? 128 * g + b : b;  // output unit (TMEM lane of D2) this thread drains = its image row
    const float w3n = n_own < H ? w3v[n_own] : 0.f;
    float dw3 = 0.f;
    auto drain_d2 = [&](int ic) {
      const int slot = ic & 1;
      const int ncols = (Hy - 64 * ic) < 64 ? (Hy - 64 * ic) : 64;
      pb_mbar_wait(PB_T2_BAR(PB_T2_B_D2FULL + slot), (uint32_t)((ic >> 1) & 1));
      pb_tc_fence_after();
      // two tiles: group g drains tile g, all columns; one tile: group g drains columns [32 g, 32 g + 32).  Every
      // worker walks the same number of 8-column steps (the noise ring is consumed in lock step)
      const int nsteps = two_tiles ? (ncols >> 3) : ((ncols >> 3) < 4 ? (ncols >> 3) : 4);
      for (int s = 0; s < nsteps; ++s) {
        const int c0 = two_tiles ? 8 * s : 32 * g + 8 * s;
        if (help && (s & 1)) {  // a helper warp drains this step
          ++zcnt;
          continue;
        }
        const float* zp = zwait();
        if (c0 < ncols) d2_step(ic, slot, c0, g, lane_off, n_own, w3n, zp, dw3);  // (uniform per warp)
        zdone();
      }
      pb_tc_fence_before();
      pb_t2_arrive(PB_T2_BAR(PB_T2_B_D2EMPTY + slot));
    };
    for (int ic = 0; ic < nchy; ++ic) {
      float v[32];
      const int c0 = 64 * ic + 32 * g;
      if (c0 < Hp) {
        uint32_t z[32];
        pb_tmem_ld32(S0 + lane_off + (uint32_t)c0, z);
        pb_tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = r * fmaxf(__uint_as_float(z[j]), 0.f);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (H >= c0 && H < c0 + 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (c0 + j == H) ? r : v[j];
      }
      chunk_begin();
      store_chunk(v);
      chunk_end();
      if (ic >= 1) drain_d2(ic - 1);
      else update_w1();  // the tensor core works on the first D2 chunk meanwhile
    }
    drain_d2(nchy - 1);
    PB_MARK(10)
    // output layer: dw3[n] (one tile: the two column halves are summed), db3 = sum_b r_b; fused update of w3, b3
    dw3p[g * 256 + (n_own & 255)] = dw3;
    fp[g * 128 + b] = r;
    if (help) asm volatile("bar.sync 2, 512;" ::: "memory");  // workers + helpers
    else asm volatile("bar.sync 1, 256;" ::: "memory");
    if ((two_tiles || g == 0) && n_own < H) {
      float gl = two_tiles ? dw3 : dw3p[n_own] + dw3p[256 + n_own];
      if (help) gl += two_tiles ? dw3h[n_own] : dw3h[n_own] + dw3h[128 + n_own];
      const int si = L2.soff + n_own * L2.wp, fi = L2.fw + n_own;
      PbPre pre = pf(si, fi);
      pre.a = TH[si];
      const float nv = uf(si, fi, pre, gl, noise ? pb_normal_at(nkey, (uint32_t)fi) : 0.f);
      TH[si] = nv;
      emit(fi, nv);
    }
    if (warp == 0) {
      float a = fp[lane] + fp[32 + lane] + fp[64 + lane] + fp[96 + lane];
      a = pb_warp_sum(a);
      if (lane == 0) {
        const int si = L2.soff + H * L2.wp, fi = L2.fb;
        PbPre pre = pf(si, fi);
        pre.a = TH[si];
        const float nv = uf(si, fi, pre, a, noise ? pb_normal_at(nkey, (uint32_t)fi) : 0.f);
        TH[si] = nv;
        emit(fi, nv);
      }
    }
    if (bad && flagword) atomicOr(flagword, 1u);
    pb_tc_fence_before();
  }
#undef PB_T2_BAR
  pb_tc_fence_before();
  PB_TAG(11)
  PB_ENDPHASE_RAW
  pb_tc_fence_after();
  PB_TAG(0)
}
#endif
