// pbnn_b200: tcgen05 (5th-gen tensor core) gradient engine for one-hidden-layer relu nets with a scalar output
// (BASELINE configs 0/1: 8-50-1, 13-50-1) -- the full-batch log-likelihood gradient of the HMC leapfrog
// (reference: jax.value_and_grad(logprob_fn) inside blackjax.hmc, src/pbnn/mcmc/hamiltonian.py:18-77 with
// benchmark/pipelines/logprobs.py:12-17).
//
// FP32 accuracy on TF32 tensor cores.  tcgen05.mma.kind::tf32 reads the top 19 bits of every fp32 operand word, so an
// operand a is fed as a = hi + lo with hi = the raw word (truncated by the hardware) and lo = a - trunc19(a) (exact in
// fp32); hi*hi + hi*lo + lo*hi accumulated in the fp32 TMEM accumulator reproduces the fp32 product to ~1e-6 relative
// (measured by tests/test_umma_experimental.py and tests/test_tc_engine.py).  Both GEMMs of the gradient are arranged so
// that one operand is (nearly) free:
//
//   forward   Hpre[b][j]  = sum_i Xaug[b][i] W1aug[i][j]        A = X tile   [128 rows][K=16]   (hi, lo: in TMEM, ONCE)
//                                                               B = W1aug^T  [64 units][K=16]   (hi, lo: per gradient)
//             3 passes x 2 k-steps per 128-row tile, D = TMEM columns 64..127 (one tile in flight)
//   epilogue  one thread per row (= TMEM lane): h = relu(Hpre), f = w2.h + b2, r = (y - f)/sigma^2, and the relu
//             MASK m[b][j] = [Hpre > 0] plus Z[b][i] = r_b Xaug[b][i] go back to shared memory as MMA operands
//   backward  Gm[j][i]    = sum_b m[b][j] Z[b][i]               mask is 0/1: EXACT in tf32, no lo part; Z = hi + lo.
//             The two 64-row halves of a tile share one MMA per k-step (M = 128 fully used, K = 64 rows):
//               A = [mask^T of rows 0..63 ; mask^T of rows 64..127]            [2 x 64 units][K = row within the half]
//               B = [Z hi ; Z lo of rows 0..63 ; Z hi ; Z lo of rows 64..127]  [4 x 16 inputs][K]
//             D (TMEM columns 0..63): lanes 0..63 x columns 0..31 = first halves (hi | lo), lanes 64..127 x columns
//             32..63 = second halves; the two off-diagonal blocks are cross terms nobody reads.  8 k-steps per tile, every
//             operand byte read from shared memory once (MMAs of this shape are shared-memory-bandwidth bound).
//             dW1[i][j] = w2_j Gm[j][i]       dw2_j = sum_b r_b relu(Hpre)[b][j] = sum_i W1aug[i][j] Gm[j][i]
//             (the second identity: relu(Hpre) = m * (Xaug W1aug), so the output-layer gradient needs no extra GEMM)
//
// Occupancy.  One gradient is a dependent chain (forward MMA -> epilogue -> backward MMA -> readout -> update); a single
// chain keeps an SM ~30 % busy.  The CTA is therefore kept small -- 4 epilogue warps + 1 MMA-issue warp, ~93 KB of
// shared memory, 256 TMEM columns -- so that TWO chains are co-resident per SM and fill each other's bubbles.
//
// Operand storage (K-major, SWIZZLE_NONE "interleaved" core matrices of 8 rows x 16 B):
//   element (row x, k) at (k/4)*LBO + x*16 B + (k%4)*4 B,   SBO (next 8 rows) = 128 B.
// The mask and Z k-groups are padded by one 16-byte row (LBO = 129 / 65 rows) so that the 32 lanes of a warp (32
// consecutive batch rows) store to 32 different banks.
// TMEM columns: 0..63 backward accumulator, 64..127 forward accumulator, 128 + 32 t .. : X tile t (16 raw | 16 lo) of
// the current group of 4 tiles (N <= 512: staged once per CTA; larger N: re-staged from L2 per group and evaluation).
#pragma once
#if !defined(PB_EMU)


PB_DEV uint32_t pb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: SWIZZLE_NONE, sm_100 descriptor version; byte offsets in 16-byte units
PB_DEV uint64_t pb_umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}

// instruction descriptor: D = f32, A = B = tf32, both K-major, dense
PB_DEV uint32_t pb_umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

PB_DEV float pb_trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

PB_DEV void pb_umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

// one lane of a converged warp (the compiler keeps tcgen05 issue on the uniform datapath only behind elect.sync;
// a plain `tid == 0` test wraps every MMA in a divergence loop: measured 90-130 cycles per MMA issued)
PB_DEV bool pb_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0u;
}

PB_DEV void pb_umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

PB_DEV void pb_mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(mbar), "r"(parity)
        : "memory");
  }
}

PB_DEV void pb_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
PB_DEV void pb_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
PB_DEV void pb_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

PB_DEV void pb_tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

PB_DEV void pb_tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

PB_DEV void pb_tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

PB_DEV void pb_tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

PB_DEV void pb_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// A operand in TMEM (lane = row, one tf32 per 32-bit column), B from shared memory
PB_DEV void pb_umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

PB_DEV void pb_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

#define PB_TC_RU 7  // first-layer weights per thread in the update phase: ceil(16 * 64 / 160)

struct PbTc {
  uint32_t tmem;  // TMEM base address (256 columns allocated)
  uint32_t ph_f;  // parity of the forward-MMA mbarrier   (epilogue warps)
  uint32_t ph_d;  // parity of the backward-MMA mbarrier  (epilogue warps)
  uint32_t ph_h;  // parity of the "forward accumulator has been read" mbarrier (issue warp)
  uint32_t wsel;  // which half of the double-buffered output-weight vector the next evaluation reads
  const float* X;  // data rows (global): re-staged into TMEM per group of 4 tiles when N > 512
  int xd, xn;      // columns / rows of X
  // update phase: this thread's first-layer weights -- constant over the chain, kept in registers
  int ridx[PB_TC_RU];  // index in the padded resident arrays (a pad slot when the element does not exist)
  int rwo[PB_TC_RU];   // index in the forward B operand Wb / Wl
  int rgx[PB_TC_RU];   // i * 72 + j in the readout scratch
  int rw2[PB_TC_RU];   // j (index of the output weight)
  uint32_t rok;        // bit u: element u exists
  // ... and, during a trajectory, their position / momentum / inverse mass (registers: no shared-memory round trip
  // per leapfrog step; written back by the last step)
  float rth[PB_TC_RU], rpm[PB_TC_RU], rmi[PB_TC_RU];
};

PB_DEV uint64_t* pb_tc_mbar(const PbPlan& pl, float* sm, int i) {
  return reinterpret_cast<uint64_t*>(sm + pl.off_tcm) + i;
}

// X rows of tiles 4 g .. 4 g + 3 (raw + lo) -> TMEM slots 0..3 as the forward A operand.  Lane = row of the tile; written
// by the thread that owns the row in the epilogue (warp w -> lanes 32 w ..).  Threads 0..127.
PB_DEV void pb_tc_stage_x(const PbTc& tc, int group, int ntiles) {
  const int tid = threadIdx.x, d = tc.xd;
  for (int t = 0; t < 4 && 4 * group + t < ntiles; ++t) {
    const int row = (4 * group + t) * 128 + tid;
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float v = 0.f;
      if (row < tc.xn) v = k < d ? tc.X[(size_t)row * d + k] : (k == d ? 1.f : 0.f);
      hi[k] = __float_as_uint(v);
      lo[k] = __float_as_uint(v - pb_trunc_tf32(v));
    }
    const uint32_t taddr = tc.tmem + ((uint32_t)(tid & ~31) << 16) + 128u + 32u * (uint32_t)t;
    pb_tmem_st16(taddr, hi);
    pb_tmem_st16(taddr + 16u, lo);
  }
  pb_tmem_wait_st();
}

// once per CTA: y, mbarriers, TMEM, X (raw + lo) into TMEM as the forward A operand
PB_DEV void pb_tc_setup(const PbPlan& pl, float* sm, const float* X, const float* y, int nrows, int nt, PbTc& tc) {
  PB_TAG(10)
  const int tid = threadIdx.x;
  const int Npad = pl.tc_npad, d = pl.d;
  float* yr = sm + pl.off_yr;
  for (int b = tid; b < Npad; b += nt) yr[b] = b < nrows ? y[b] : 0.f;
  if (tid == 0) {
    // 0: forward MMAs done; 1: forward accumulator read by all 128 epilogue threads; 4: backward MMAs done
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pb_smem_u32(pb_tc_mbar(pl, sm, 0))));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(pb_smem_u32(pb_tc_mbar(pl, sm, 1))));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pb_smem_u32(pb_tc_mbar(pl, sm, 4))));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (tid < 32) {  // TMEM allocation by one full warp: half of the SM's columns (two CTAs per SM)
    // (immediate column count: with a register operand the occupancy calculator assumes the whole TMEM -> 1 CTA/SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(
        pb_smem_u32(sm + pl.off_tcm + 12)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  pb_tc_fence_before();
  PB_ENDPHASE_RAW
  pb_tc_fence_after();
  tc.tmem = *reinterpret_cast<const uint32_t*>(sm + pl.off_tcm + 12);
  tc.X = X;
  tc.xd = d;
  tc.xn = nrows;
  if (tid < 128) pb_tc_stage_x(tc, 0, Npad >> 7);
  pb_tc_fence_before();
  PB_ENDPHASE_RAW
  pb_tc_fence_after();
  tc.ph_f = tc.ph_d = tc.ph_h = tc.wsel = 0u;
  tc.rok = 0u;
  for (int u = 0; u < PB_TC_RU; ++u) {
    // element e -> (i, j) with i % 4 fastest: 4 consecutive lanes fill one 16-byte slot of the forward B operand
    // (conflict-free Wb / Wl stores)
    const int e = pb_min(tid + u * nt, 16 * 64 - 1);
    const int i = (e & 3) + 4 * (e >> 8), j = (e >> 2) & 63;
    const bool ok = tid + u * nt < 16 * 64 && i < pl.L[0].inA && j < pl.L[0].out;
    tc.rok |= ok ? (1u << u) : 0u;
    // missing elements use a pad slot of the output layer (row 0, column 1: always zero and stays zero)
    tc.ridx[u] = ok ? pl.L[0].soff + i * pl.L[0].wp + j : pl.L[1].soff + 1;
    tc.rwo[u] = (i >> 2) * 256 + j * 4 + (i & 3);
    tc.rw2[u] = j;
    tc.rgx[u] = i * 72 + j;  // pitch 72: the 4 x 8 (i, j) pairs of a warp hit 32 different banks
  }
}

PB_DEV void pb_tc_teardown(const PbPlan& pl, float* sm, const PbTc& tc) {
  pb_tc_fence_before();
  PB_ENDPHASE_RAW
  if (threadIdx.x < 32) {
    pb_tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tc.tmem));
  }
}

// leapfrog quantities for the fused readout (MODE 1 / 2 of pb_tc_grad_core)
struct PbTcLeap {
  float* PM;        // momentum (padded resident layout)
  const float* MI;  // inverse mass diagonal
  float eps, heps, inv_pv;
};

// W1aug -> forward B operand (raw + lo), w2 -> contiguous vector; with FIRST_PRE the first half-kick + drift of a
// trajectory (p += eps/2 g; theta += eps Minv p) is applied on the way (same arithmetic as the scalar engines).
template <bool FIRST_PRE>
PB_DEV void pb_tc_stage_w(const PbPlan& pl, float* sm, float* TH, const float* G, const PbTcLeap& lf, int nt,
                          PbTc& tc) {
  PB_TAG(1)
  const int tid = threadIdx.x;
  const PbLayer& L0 = pl.L[0];
  const PbLayer& L1 = pl.L[1];
  const int H = L0.out, wp = L0.wp, inA = L0.inA;
  float* Wb = sm + pl.off_wb;
  float* Wl = Wb + 16 * 64;
  float* w2s = sm + pl.off_w2s + 64 * (int)tc.wsel;
  auto upd = [&](int idx) -> float {
    float th = TH[idx];
    if (FIRST_PRE) {
      const float p = lf.PM[idx] + lf.heps * G[idx];
      lf.PM[idx] = p;
      th = th + lf.eps * (lf.MI[idx] * p);
      TH[idx] = th;
    }
    return th;
  };
  if (FIRST_PRE) {
    // start of a trajectory: this thread's first-layer weights move into registers (same enumeration as the update
    // phase of pb_tc_grad_core); the shared-memory copies of theta / p are stale until the last step writes them back
#pragma unroll
    for (int u = 0; u < PB_TC_RU; ++u) {
      const int idx = tc.ridx[u];
      const float mi = lf.MI[idx];
      const float p = lf.PM[idx] + lf.heps * G[idx];
      const float th = TH[idx] + lf.eps * (mi * p);
      tc.rth[u] = th;
      tc.rpm[u] = p;
      tc.rmi[u] = mi;
      const float v = (tc.rok >> u) & 1u ? th : 0.f;
      Wb[tc.rwo[u]] = v;
      Wl[tc.rwo[u]] = v - pb_trunc_tf32(v);
    }
  } else {
#pragma unroll 4
    for (int e = tid; e < 16 * 64; e += nt) {
      const int i = (e & 3) + 4 * (e >> 8), j = (e >> 2) & 63;
      const float v = (i < inA && j < H) ? upd(L0.soff + i * wp + j) : 0.f;
      const int o = (i >> 2) * 256 + j * 4 + (i & 3);
      Wb[o] = v;
      Wl[o] = v - pb_trunc_tf32(v);
    }
  }
  if (tid < 64) w2s[tid] = tid < H ? upd(L1.soff + tid * L1.wp) : 0.f;
  if (FIRST_PRE && tid == 64) upd(L1.soff + H * L1.wp);
  pb_fence_async_smem();
  pb_tc_fence_before();
  PB_ENDPHASE_RAW
}

// 6 forward MMAs of one tile (3xTF32 x 2 k-steps): A = X tile in TMEM (raw at +0, lo at +16; second k-step + 8
// columns), B = W1aug^T in shared memory (raw, lo at + 16*64 floats; second k-step + 2 k-groups)
PB_DEV void pb_tc_issue_fwd(const PbTc& tc, uint64_t dwb, uint32_t idesc, int tile, bool k2, uint32_t mb_f) {
  const uint32_t dcol = tc.tmem + 64u, ta = tc.tmem + 128u + 32u * (uint32_t)(tile & 3);
  const uint64_t w_lo = 256u, w_k = 128u;
  pb_umma_tf32_ts(dcol, ta, dwb, idesc, 0u);  // hi * hi
  if (k2) pb_umma_tf32_ts(dcol, ta + 8u, dwb + w_k, idesc, 1u);
  pb_umma_tf32_ts(dcol, ta, dwb + w_lo, idesc, 1u);  // hi * lo
  if (k2) pb_umma_tf32_ts(dcol, ta + 8u, dwb + w_lo + w_k, idesc, 1u);
  pb_umma_tf32_ts(dcol, ta + 16u, dwb, idesc, 1u);  // lo * hi
  if (k2) pb_umma_tf32_ts(dcol, ta + 24u, dwb + w_k, idesc, 1u);
  pb_umma_commit(mb_f);
}

// The gradient proper; the forward B operand (Wb, Wl), w2s and TH must be current (pb_tc_stage_w or a MODE 1 call).
// 160 threads: warps 0-3 are the epilogue / readout warps (warp w owns TMEM lanes 32 w .. = rows of the tile), warp 4
// issues the MMAs.  One 128-row tile per round.
//   MODE 0: G <- d loglik / d theta (padded resident layout, pads untouched = 0)
//   MODE 1: leapfrog interior, fused into the readout: G <- grad log posterior, p += eps/2 G (end of this step),
//           p += eps/2 G, theta += eps Minv p (start of the next step), next W operands staged
//   MODE 2: last step of a trajectory: G <- grad log posterior, p += eps/2 G
// BIG: N > 512 (a separate kernel, so that the X re-staging code stays out of the headline configuration's round loop).
// always: llbuf[0..127] <- per-warp log-likelihood partials.
template <int MODE, bool BIG>
PB_DEV void pb_tc_grad_core(const PbPlan& pl, float* sm, float* TH, float* G, int nrows, float scale, int nt,
                            PbTc& tc, const PbTcLeap& lf) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const PbLayer& L0 = pl.L[0];
  const PbLayer& L1 = pl.L[1];
  const int H = L0.out, inA = L0.inA;
  const int ntiles = pl.tc_npad >> 7, nj8 = (H + 7) >> 3;
  const float* yr = sm + pl.off_yr;
  float* Wb = sm + pl.off_wb;
  float* Wl = Wb + 16 * 64;
  const float* w2s = sm + pl.off_w2s + 64 * (int)tc.wsel;  // output weights of this evaluation (double-buffered)
  float* w2nxt = sm + pl.off_w2s + 64 * (int)(tc.wsel ^ 1u);
  float* rs = sm + pl.off_tcm + 16;
  float* gx = sm + pl.off_z;  // readout scratch (half 0 | half 1 partial sums [2][16][72], dots [2][64]): the Z buffer
                              // is free once the last backward MMAs are done
  float* llbuf = pb_llbuf(pl, sm);
  const uint32_t mb_f = pb_smem_u32(pb_tc_mbar(pl, sm, 0)), mb_h = pb_smem_u32(pb_tc_mbar(pl, sm, 1));
  const uint32_t mb_d = pb_smem_u32(pb_tc_mbar(pl, sm, 4));
  const float b2 = TH[L1.soff + H * L1.wp];
  PB_TAG(2)
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);  // warp-uniform for the compiler
  const uint32_t idesc = pb_umma_idesc_tf32(128, 64);
  const uint64_t dwb = pb_umma_desc(pb_smem_u32(Wb), 1024u, 128u);
  const bool k2 = pl.h1_dp > 2;
  if (BIG) {  // more than one group of 4 tiles (N > 512): the previous evaluation left the last group in TMEM
    if (warp_u < 4) pb_tc_stage_x(tc, 0, ntiles);
    pb_tc_fence_before();
    PB_ENDPHASE_RAW
  }
  if (warp_u == 4) {
    pb_tc_fence_after();
    if (pb_elect_one()) pb_tc_issue_fwd(tc, dwb, idesc, 0, k2, mb_f);
    __syncwarp();
  }

  // ---- one tile per round ------------------------------------------------------------------------------------
  const int half = tid >> 6, kk = tid & 63;  // backward operands: row kk of half `half` (epilogue threads)
  float ll = 0.f, rsum = 0.f;
  for (int round = 0; round < ntiles; ++round) {
    PB_TAG(round == 0 ? 2 : 4)
    if (warp_u < 4) {
      pb_mbar_wait(mb_f, tc.ph_f);
      tc.ph_f ^= 1u;
      PB_MARK(21)
      pb_tc_fence_after();
      uint32_t h[64], xr[16];
      const uint32_t lane_base = tc.tmem + ((uint32_t)(warp * 32) << 16);
      pb_tmem_ld32(lane_base + 64u, h);
      pb_tmem_ld32(lane_base + 96u, h + 32);
      pb_tmem_ld16(lane_base + 128u + 32u * (uint32_t)(BIG ? (round & 3) : round), xr);
      pb_tmem_wait_ld();
      // group boundary: every forward MMA of this group is done, bring the next 4 tiles of X into TMEM
      if (BIG && (round & 3) == 3 && round + 1 < ntiles) pb_tc_stage_x(tc, (round + 1) >> 2, ntiles);
      pb_tc_fence_before();
      if (round + 1 < ntiles)  // the accumulator is free: the issue warp may start the next tile's forward MMAs
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(mb_h) : "memory");
      float2 f01 = make_float2(0.f, 0.f);  // (.x: even units, .y: odd units -- packed FMA, same sums as two scalar chains)
      // blocks of 8 hidden units behind a (uniform) branch: units >= H are skipped at that granularity
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8)
        if (c8 < nj8) {
#pragma unroll
          for (int j4 = 2 * c8; j4 < 2 * c8 + 2; ++j4) {
            const float4 w = pb_ld4(w2s + 4 * j4);
            f01 = __ffma2_rn(make_float2(fmaxf(__uint_as_float(h[4 * j4]), 0.f), fmaxf(__uint_as_float(h[4 * j4 + 1]), 0.f)),
                             make_float2(w.x, w.y), f01);
            f01 = __ffma2_rn(make_float2(fmaxf(__uint_as_float(h[4 * j4 + 2]), 0.f), fmaxf(__uint_as_float(h[4 * j4 + 3]), 0.f)),
                             make_float2(w.z, w.w), f01);
          }
        }
      const float f0 = f01.x, f1 = f01.y;
      const int row = round * 128 + tid;
      const bool valid = row < nrows;
      const float res = yr[row] - ((f0 + f1) + b2);
      const float r = valid ? res * scale : 0.f;
      ll += valid ? -0.5f * res * res * pl.inv_s2 : 0.f;
      rsum += r;
      if (round > 0) {  // the previous tile's backward MMAs have finished reading the mask / Z buffers
        pb_mbar_wait(mb_d, tc.ph_d);
        tc.ph_d ^= 1u;
        PB_MARK(24)
      }
      float* mk = sm + pl.off_mk + (kk >> 2) * PB_TC_MK_LBO + (half * 64) * 4 + (kk & 3);
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8)
        if (c8 < nj8) {
#pragma unroll
          for (int j = 8 * c8; j < 8 * c8 + 8; ++j) mk[j * 4] = __uint_as_float(h[j]) > 0.f ? 1.f : 0.f;
        }
      float* zr = sm + pl.off_z + (kk >> 2) * PB_TC_Z_LBO + (half * 32) * 4 + (kk & 3);
#pragma unroll
      for (int i = 0; i < 16; i += 2) {  // (packed multiply / subtract: per component the same arithmetic)
        const float2 z = __fmul2_rn(make_float2(r, r), make_float2(__uint_as_float(xr[i]), __uint_as_float(xr[i + 1])));
        const float2 lo = __ffma2_rn(make_float2(pb_trunc_tf32(z.x), pb_trunc_tf32(z.y)), make_float2(-1.f, -1.f), z);
        zr[i * 4] = z.x;
        zr[(i + 1) * 4] = z.y;
        zr[(16 + i) * 4] = lo.x;
        zr[(17 + i) * 4] = lo.y;
      }
      PB_MARK(22)
      pb_fence_async_smem();
      PB_MARK(25)
      if (round == ntiles - 1) {
        ll = pb_warp_sum(ll);
        rsum = pb_warp_sum(rsum);
        if (lane == 0) {
          llbuf[warp] = ll;
          rs[warp] = rsum;
        }
        if (tid >= 4) llbuf[tid] = 0.f;
      }
    } else if (round + 1 < ntiles) {
      // issue warp: the next tile's forward MMAs start as soon as the accumulator has been read
      pb_mbar_wait(mb_h, tc.ph_h);
      tc.ph_h ^= 1u;
      pb_tc_fence_after();
      if (pb_elect_one()) pb_tc_issue_fwd(tc, dwb, idesc, round + 1, k2, mb_f);
      __syncwarp();
    }
    pb_tc_fence_before();
    PB_ENDPHASE_RAW
    PB_TAG(5)
    if (warp_u == 4) {
      pb_tc_fence_after();
      if (pb_elect_one()) {
        const uint64_t dm = pb_umma_desc(pb_smem_u32(sm + pl.off_mk), PB_TC_MK_LBO * 4u, 128u);
        const uint64_t dz = pb_umma_desc(pb_smem_u32(sm + pl.off_z), PB_TC_Z_LBO * 4u, 128u);
        pb_umma_tf32(tc.tmem, dm, dz, idesc, round == 0 ? 0u : 1u);
#pragma unroll
        for (int kb = 1; kb < 8; ++kb)
          pb_umma_tf32(tc.tmem, dm + (uint64_t)(kb * (2 * PB_TC_MK_LBO * 4 / 16)),
                       dz + (uint64_t)(kb * (2 * PB_TC_Z_LBO * 4 / 16)), idesc, 1u);
        pb_umma_commit(mb_d);
      }
      __syncwarp();
    }
  }
  if (MODE == 1) tc.wsel ^= 1u;  // the update phase below writes the next evaluation's output weights to w2nxt
  PB_TAG(3)
  // ---- readout A (warps 0-3): TMEM lane = hidden unit (lanes 64..127: the second halves' partial sums), 16 columns
  //      hi + 16 columns lo -> gx[half][i][j]; partial dot products for the output-layer weights -> gx[2][half][j]
  if (warp < 4) {
    pb_mbar_wait(mb_d, tc.ph_d);
    tc.ph_d ^= 1u;
    PB_MARK(24)
    pb_tc_fence_after();
    uint32_t g[32];
    pb_tmem_ld32(tc.tmem + ((uint32_t)(warp * 32) << 16) + (warp >= 2 ? 32u : 0u), g);
    pb_tmem_wait_ld();
    pb_tc_fence_before();
    const int j = tid & 63, hf = warp >> 1;
    float* gh = gx + hf * 1152;
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float gm = __uint_as_float(g[i]) + __uint_as_float(g[16 + i]);
      gh[i * 72 + j] = gm;
      if (i < inA && j < H) dot = fmaf(Wb[(i >> 2) * 256 + j * 4 + (i & 3)], gm, dot);  // Wb raw = current W1
    }
    gx[2304 + hf * 64 + j] = dot;
  }
  PB_ENDPHASE_RAW
  // ---- readout B (all threads): gradient of the log posterior and, fused, the leapfrog updates + next W operands
  PB_TAG(6)
  auto upd = [&](int idx, float gll, float& th_new) -> void {
    // gll = d loglik / d theta[idx]; MODE 0 stores it, MODE 1/2 apply the half-kick(s) and the drift
    if (MODE == 0) {
      G[idx] = gll;
      return;
    }
    const float th = TH[idx];
    const float gp = gll - th * lf.inv_pv;
    if (MODE == 2) G[idx] = gp;
    float p = lf.PM[idx] + lf.heps * gp;
    if (MODE == 1) {
      p = p + lf.heps * gp;
      th_new = th + lf.eps * (lf.MI[idx] * p);
      TH[idx] = th_new;
    }
    lf.PM[idx] = p;
  };
  {
    // first-layer weights, branch-free: a missing element works on a zero pad slot with a zero gradient (every
    // update of it is 0 -> 0); all loads first (the stores may alias for the compiler)
    float gl[PB_TC_RU];
#pragma unroll
    for (int u = 0; u < PB_TC_RU; ++u) {
      const float gsum = gx[tc.rgx[u]] + gx[1152 + tc.rgx[u]];
      gl[u] = (tc.rok >> u) & 1u ? w2s[tc.rw2[u]] * gsum : 0.f;
    }
#pragma unroll
    for (int u = 0; u < PB_TC_RU; ++u) {
      if (MODE == 0) {
        G[tc.ridx[u]] = gl[u];
      } else {
        // theta, p, Minv of these weights live in registers for the whole trajectory (pb_tc_stage_w<true>)
        const float gp = gl[u] - tc.rth[u] * lf.inv_pv;
        float p = tc.rpm[u] + lf.heps * gp;
        if (MODE == 1) {
          p = p + lf.heps * gp;
          const float tn = tc.rth[u] + lf.eps * (tc.rmi[u] * p);
          tc.rth[u] = tn;
          Wb[tc.rwo[u]] = tn;
          Wl[tc.rwo[u]] = tn - pb_trunc_tf32(tn);
        } else {  // last step: everything back to the resident arrays
          G[tc.ridx[u]] = gp;
          TH[tc.ridx[u]] = tc.rth[u];
          lf.PM[tc.ridx[u]] = p;
        }
        tc.rpm[u] = p;
      }
    }
  }
  if (tid < H) {  // the output weights are double-buffered: the loop above (other threads) still reads the current ones
    float w2n = 0.f;
    upd(L1.soff + tid * L1.wp, gx[2304 + tid] + gx[2304 + 64 + tid], w2n);
    if (MODE == 1) w2nxt[tid] = w2n;
  }
  if (tid == 64) {
    float a = 0.f, bn = 0.f;
    for (int w = 0; w < 4; ++w) a += rs[w];
    upd(L1.soff + H * L1.wp, a, bn);
  }
  if (MODE == 1) pb_fence_async_smem();
  pb_tc_fence_before();
  PB_ENDPHASE_RAW
  PB_TAG(0)
}

// plain gradient (initial point of a chain; pbnn_logposterior_grad)
template <bool BIG>
PB_DEV void pb_grad_tc(const PbPlan& pl, float* sm, float* TH, float* G, int nrows, float scale, int nt, PbTc& tc) {
  PbTcLeap lf = {nullptr, nullptr, 0.f, 0.f, 0.f};
  pb_tc_stage_w<false>(pl, sm, TH, G, lf, nt, tc);
  pb_tc_grad_core<0, BIG>(pl, sm, TH, G, nrows, scale, nt, tc, lf);
}
#endif
