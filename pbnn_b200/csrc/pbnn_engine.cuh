// Shared-memory resident chain engine: one CTA advances one chain (or ensemble member) with its
// weights, gradient and sampler state resident in shared memory for the whole trajectory.
//
// The bodies are written as a sequence of PHASES.  On the GPU a phase is executed by every thread of
// the CTA and ends with __syncthreads(); no thread-private value lives across a phase boundary.  That
// discipline lets tests/emu compile the very same bodies with g++ (PB_EMU) and run the phases as
// loops over `tid` -- a CPU emulator of the kernel logic used ONLY by the test-suite to check
// indexing before GPU time is spent.  It is not a product path.
//
// Data layout in shared memory (floats), see pb_build_plan():
//   nstate x [Ps]   resident parameter-sized arrays; layer l is W_aug[rows][wp] with the bias as row `in`
//   gpart           (ksmax-1) x [Ps] partial weight gradients of the split-K slices
//   A_1..A_nl       activations, FEATURE-MAJOR: A_l[feature][b], pitch ap = BT+4; A_l has a constant
//                   ones-row at index dims[l] so the bias and its gradient fall out of the GEMMs
//   X^T_aug, y^T    all input rows of one gradient evaluation, feature-major, resident (pitch xp); a batch
//                   tile is a column window of it.  (Fallback when they do not fit: A_0 / yT tile buffers
//                   streamed from global memory.)
// Three register-tiled FP32 GEMMs (4x4 outputs per thread, LDS.128 operands, conflict-free by pitch):
//   fwd : A_{l+1}[o][b] = act( sum_k W[k][o] * A_l[k][b] )            (k-major both operands)
//   dW  : G[i][o]      += sum_b A_l[i][b] * dZ[o][b]                 (dot form, both k-contiguous, split-K)
//   dH  : A_l[i][b]     = act'(A_l[i][b]) * sum_o W[i][o] * dZ[o][b]  (in place)
#pragma once
#include "pbnn_core.h"
#include "pbnn_rng.cuh"

#if defined(PB_EMU)
struct float4 {
  float x, y, z, w;
};
#define PB_PHASE for (int tid = 0; tid < nt; ++tid) {
#define PB_ENDPHASE }
#define PB_ATOMIC_OR(p, v) (*(p) |= (v))
#else
#define PB_PHASE            \
  {                         \
    const int tid = threadIdx.x;
#if defined(PB_PROFILE)
// per-phase cycle accounting of CTA 0 (profiling build only, -DPB_PROFILE): PB_TAG(n) names the phases that follow
__device__ unsigned long long pb_prof_cycles[64];
__device__ unsigned long long pb_prof_last;
__device__ int pb_prof_tag;
#define PB_TAG(n)                                          \
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) pb_prof_tag = (n);
#define PB_ENDPHASE                                                      \
  }                                                                      \
  __syncthreads();                                                       \
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {          \
    const unsigned long long now_ = clock64();                           \
    pb_prof_cycles[pb_prof_tag & 63] += now_ - pb_prof_last;             \
    pb_prof_cycles[32 + (pb_prof_tag & 31)] += 1;                        \
    pb_prof_last = now_;                                                 \
  }
// PB_MARK(n): charge the cycles since the last mark / phase end to slot n (subdivides a phase; thread 0 of CTA 0)
#define PB_MARK(n)                                                       \
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {          \
    const unsigned long long now_ = clock64();                           \
    pb_prof_cycles[(n) & 31] += now_ - pb_prof_last;                     \
    pb_prof_cycles[32 + ((n) & 31)] += 1;                                \
    pb_prof_last = now_;                                                 \
  }
#else
#define PB_MARK(n)
#define PB_ENDPHASE \
  }                 \
  __syncthreads();
#endif
#define PB_ATOMIC_OR(p, v) atomicOr((p), (v))
#define PB_ENDPHASE_RAW { PB_ENDPHASE
#endif
#ifndef PB_MARK
#define PB_MARK(n)
#endif
#ifndef PB_TAG
#define PB_TAG(n)
#endif

PB_DEV float4 pb_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
PB_DEV void pb_st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
PB_DEV int pb_min(int a, int b) { return a < b ? a : b; }
PB_DEV int pb_r4dev(int x) { return (x + 3) & ~3; }
PB_DEV bool pb_finite(float x) { return fabsf(x) <= 3.402823466e+38f; }
// Per-parameter square roots / reciprocals of the pSGLD and Adam updates: MUFU approximations on the device (2^-22
// relative each, on terms the parity contract resolves to 1e-5; the IEEE sqrtf and division are ~100 instructions per
// parameter and step -- as much as the bit-exact normal the same parameter draws), the exact ones in the emulator.
PB_DEV float pb_sqrt_fast(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}
PB_DEV float pb_rcp_fast(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}
#if defined(PB_EMU)
static inline float __fdividef(float a, float b) { return a / b; }
#endif

// packed FP32 pair FMA (Blackwell FFMA2, fma.rn.f32x2): same FLOP rate as two scalar FFMA at half the issue slots
#if defined(PB_EMU)
struct float2 {
  float x, y;
};
static inline float2 make_float2(float x, float y) {
  float2 r;
  r.x = x;
  r.y = y;
  return r;
}
static inline float2 pb_fma2(float2 a, float2 b, float2 c) { return make_float2(a.x * b.x + c.x, a.y * b.y + c.y); }
#else
PB_DEV float2 pb_fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#endif

template <int ACT>
PB_DEV float pb_act(float z) {
  return ACT == PBNN_ACT_RELU ? fmaxf(z, 0.f) : tanhf(z);
}
template <int ACT>
PB_DEV float pb_actgrad(float h) {  // derivative expressed through the activation OUTPUT
  return ACT == PBNN_ACT_RELU ? (h > 0.f ? 1.f : 0.f) : 1.f - h * h;
}

// shared-memory views (no pointer arrays: dynamic indexing would push them to local memory)
// GS (compile time) = state arrays in the global workspace; keeping it a template parameter lets the compiler see
// that the resident arrays are shared-memory pointers (LDS/STS instead of generic loads) in the common case
template <bool GS>
PB_DEV float* pb_state(const PbPlan& pl, float* sm, float* ws, int i) {
  return (GS ? ws : sm) + i * pl.Ps;
}
PB_DEV float* pb_A(const PbPlan& pl, float* sm, int l) { return sm + pl.aoff[l]; }
PB_DEV float* pb_scr(const PbPlan& pl, float* sm) { return sm + pl.off_scr; }
PB_DEV float* pb_llbuf(const PbPlan& pl, float* sm) { return sm + pl.off_scr + PB_SCR_MAIN + 64; }
PB_DEV float* pb_fbuf(const PbPlan& pl, float* sm) { return sm + pl.off_scr + PB_SCR_MAIN + 64 + 128; }
PB_DEV float* pb_acc(const PbPlan& pl, float* sm) { return sm + pl.off_acc; }
PB_DEV int* pb_idx(const PbPlan& pl, float* sm) { return reinterpret_cast<int*>(sm + pl.off_idx); }
PB_DEV float* pb_sc(const PbPlan& pl, float* sm) { return sm + pl.off_scal; }
PB_DEV uint32_t* pb_scu(const PbPlan& pl, float* sm) { return reinterpret_cast<uint32_t*>(sm + pl.off_scal); }

// input tile `tile` of the current gradient evaluation: pointer + pitch of X^T_aug and y^T
struct PbTile {
  const float* x;
  const float* y;
  int px, py;
};
PB_DEV PbTile pb_tile(const PbPlan& pl, float* sm, int tile) {
  PbTile t;
  if (pl.res_rows > 0) {
    t.x = sm + pl.off_xres + tile * pl.BT;
    t.y = sm + pl.off_yres + tile * pl.BT;
    t.px = t.py = pl.xp;
  } else {
    t.x = sm + pl.aoff[0];
    t.y = sm + pl.aoff[pl.nl + 1];
    t.px = t.py = pl.ap;
  }
  return t;
}

// visit every real parameter: f(resident_index, flat_index).  Consecutive tid -> consecutive columns,
// so flat addresses are contiguous inside a kernel row (coalesced trace / state traffic).
template <class F>
PB_DEV void pb_for_each_param(const PbPlan& pl, int tid, int nt, F f) {
  for (int l = 0; l < pl.nl; ++l) {
    const PbLayer& Ly = pl.L[l];
    const int total = Ly.inA * Ly.wp;
    int r = tid / Ly.wp, c = tid - r * Ly.wp;
    const int dr = nt / Ly.wp, dc = nt - dr * Ly.wp;
    for (int e = tid; e < total; e += nt) {
      if (c < Ly.out) f(Ly.soff + e, (r < Ly.in) ? (Ly.fw + r * Ly.out + c) : (Ly.fb + c));
      r += dr;
      c += dc;
      if (c >= Ly.wp) {
        c -= Ly.wp;
        ++r;
      }
    }
  }
}

// flat (ravel_pytree) index -> index in the resident padded layout
PB_DEV int pb_flat_to_resident(const PbPlan& pl, int fi) {
  int l = 0;
  for (int k = 1; k < pl.nl; ++k)
    if (fi >= pl.L[k].fb) l = k;
  const PbLayer& Ly = pl.L[l];
  int q = fi - Ly.fb;
  if (q < Ly.out) return Ly.soff + Ly.in * Ly.wp + q;  // bias row
  q -= Ly.out;
  const int r = (int)(((unsigned long long)(unsigned)q * Ly.div_out) >> 40);
  return Ly.soff + r * Ly.wp + (q - r * Ly.out);
}

// values an update needs from the state arrays, loaded BEFORE the threefry + erf_inv chains of a group are evaluated:
// with the state arrays in global memory (L2) the loads of a whole group are in flight while the integer work runs,
// instead of one dependent load -> compute -> store round trip per element
struct PbPre {
  float a, b, c, d, e;
};

// One jax.random.normal(key, (P,))[flat] per parameter: pre = pf(resident_index, flat_index) is loaded first, then
// f(resident_index, flat_index, z, pre) is called, iterating over the FLAT index (no pad slots, coalesced trace order).
// The threefry2x32 + erf_inv chains of ILP parameters are evaluated together: a single chain is ~250 cycles of
// dependent integer ops, so instruction-level parallelism (not more warps) is what hides its latency.
template <int ILP, class PF, class F>
PB_DEV void pb_flat_normal_ilp(const PbPlan& pl, int tid, int nt, PbKey key, PF pf, F f) {
  for (int p0 = tid; p0 < pl.P; p0 += ILP * nt) {
    int si[ILP];
    float z[ILP];
    PbPre pre[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      const int fi = p0 + j * nt;
      si[j] = fi < pl.P ? pb_flat_to_resident(pl, fi) : -1;
    }
#pragma unroll
    for (int j = 0; j < ILP; ++j)
      if (si[j] >= 0) pre[j] = pf(si[j], p0 + j * nt);
#pragma unroll
    for (int j = 0; j < ILP; ++j) z[j] = pb_normal_at(key, (uint32_t)(p0 + j * nt));
#pragma unroll
    for (int j = 0; j < ILP; ++j)
      if (si[j] >= 0) f(si[j], p0 + j * nt, z[j], pre[j]);
  }
}

// Large nets: iterate over the padded resident index instead (pad slots are a few % there, and the flat -> resident
// mapping per element costs more than they do); 4 chains in flight as above.
template <class PF, class F>
PB_DEV void pb_padded_normal_ilp4(const PbPlan& pl, int tid, int nt, PbKey key, PF pf, F f) {
  for (int l = 0; l < pl.nl; ++l) {
    const PbLayer& Ly = pl.L[l];
    const int total = Ly.inA * Ly.wp;
    const float inv_wp = 1.0f / (float)Ly.wp;
    for (int e0 = tid; e0 < total; e0 += 4 * nt) {
      int si[4], fi[4];
      float z[4];
      PbPre pre[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = e0 + j * nt;
        const int r = (int)(((float)e + 0.5f) * inv_wp);  // exact: (e + .5)/wp is >= .5/wp away from an integer
        const int c = e - r * Ly.wp;
        const bool ok = e < total && c < Ly.out;
        si[j] = ok ? Ly.soff + e : -1;
        fi[j] = (r < Ly.in) ? (Ly.fw + r * Ly.out + c) : (Ly.fb + c);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (si[j] >= 0) pre[j] = pf(si[j], fi[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) z[j] = pb_normal_at(key, (uint32_t)fi[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (si[j] >= 0) f(si[j], fi[j], z[j], pre[j]);
    }
  }
}

template <class PF, class F>
PB_DEV void pb_for_each_param_normal(const PbPlan& pl, int tid, int nt, PbKey key, PF pf, F f) {
  if (pl.P > 16 * nt) pb_padded_normal_ilp4(pl, tid, nt, key, pf, f);
  else if (pl.P > 3 * nt) pb_flat_normal_ilp<4>(pl, tid, nt, key, pf, f);
  else if (pl.P > nt) pb_flat_normal_ilp<2>(pl, tid, nt, key, pf, f);
  else pb_flat_normal_ilp<1>(pl, tid, nt, key, pf, f);
}

// element-wise pass over the padded resident arrays, 4 elements per thread in flight: all loads (ld -> PbPre) of a
// group are issued before the first store (the arrays may alias for the compiler)
template <class LD, class ST>
PB_DEV void pb_elementwise4(int n, int tid, int nt, LD ld, ST st) {
  for (int i0 = tid; i0 < n; i0 += 4 * nt) {
    PbPre v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i0 + j * nt < n) v[j] = ld(i0 + j * nt);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i0 + j * nt < n) st(i0 + j * nt, v[j]);
  }
}

// ------------------------------------------------------------------------------------------------
// GEMMs (per-thread bodies; warp w owns work units w, w+NW, ...)
// ------------------------------------------------------------------------------------------------
#define PB_FMA4x4(acc, w, h)                                                         \
  acc[0][0] += w.x * h.x; acc[0][1] += w.x * h.y; acc[0][2] += w.x * h.z; acc[0][3] += w.x * h.w; \
  acc[1][0] += w.y * h.x; acc[1][1] += w.y * h.y; acc[1][2] += w.y * h.z; acc[1][3] += w.y * h.w; \
  acc[2][0] += w.z * h.x; acc[2][1] += w.z * h.y; acc[2][2] += w.z * h.z; acc[2][3] += w.z * h.w; \
  acc[3][0] += w.w * h.x; acc[3][1] += w.w * h.y; acc[3][2] += w.w * h.z; acc[3][3] += w.w * h.w;

#define PB_ZERO4x4(acc)           \
  _Pragma("unroll") for (int i_ = 0; i_ < 4; ++i_) _Pragma("unroll") for (int j_ = 0; j_ < 4; ++j_) acc[i_][j_] = 0.f;

// warp tile: 16 outputs (4 lane-groups x 4) x 32 batch rows (8 lane-groups x 4)
template <int ACT, bool IDENT>
PB_DEV void pb_gemm_fwd(const PbLayer& Ly, const float* W, const float* Ain, int pin, float* Aout, int pout, int BT,
                        int tid, int nt) {
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const int lm = lane & 7, ln = lane >> 3;
  const int tiles_b = BT >> 5, tiles_o = (Ly.out + 15) >> 4;
  const int wp = Ly.wp, K = Ly.rows;
  for (int wt = warp; wt < tiles_b * tiles_o; wt += nw) {
    const int to = wt / tiles_b, tb = wt - to * tiles_b;
    const int o0 = to * 16 + ln * 4, b0 = tb * 32 + lm * 4;
    const float* wptr = W + pb_min(o0, wp - 4);
    const float* aptr = Ain + b0;
    float acc[4][4];
    PB_ZERO4x4(acc)
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 w = pb_ld4(wptr + (k + kk) * wp);
        const float4 h = pb_ld4(aptr + (k + kk) * pin);
        PB_FMA4x4(acc, w, h)
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o = o0 + j;
      if (o < Ly.out) {
        float4 v;
        if (IDENT) {
          v.x = acc[j][0]; v.y = acc[j][1]; v.z = acc[j][2]; v.w = acc[j][3];
        } else {
          v.x = pb_act<ACT>(acc[j][0]); v.y = pb_act<ACT>(acc[j][1]);
          v.z = pb_act<ACT>(acc[j][2]); v.w = pb_act<ACT>(acc[j][3]);
        }
        pb_st4(Aout + o * pout + b0, v);
      }
    }
  }
}

// Weight gradient.  Lane layout NI x NO (8x4 or 4x8): thread (li, lo) owns rows li + NI*j and columns
// lo + NO*j (j < 4), i.e. a warp tile of 4*NI input rows x 4*NO outputs; <= 8 distinct consecutive rows per
// LDS.128 => one wavefront each.  The batch (K) range is split into Ly.dw_ks slices; slice 0 accumulates into G,
// slice k > 0 into the partial array k-1 (summed by pb_reduce_gparts once per gradient evaluation).
template <int NI>
PB_DEV void pb_gemm_dw(const PbLayer& Ly, const float* Ain, int pin, const float* dZ, int pz, float* G, float* Gpart,
                       int Ps, int BT, bool first, int tid, int nt) {
  constexpr int NO = 32 / NI, TI = 4 * NI, TO = 4 * NO;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const int li = lane % NI, lo = lane / NI;
  const int tiles_i = (Ly.inA + TI - 1) / TI, tiles_o = (Ly.out + TO - 1) / TO;
  const int ks = Ly.dw_ks, kchunk = BT / ks;
  const int units = tiles_i * tiles_o * ks;
  for (int u = warp; u < units; u += nw) {
    const int t = u / ks, kslice = u - t * ks;
    const int ti = t / tiles_o, to = t - ti * tiles_o;
    const float* ap_[4];
    const float* dp_[4];
    int ir[4], oc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ir[j] = ti * TI + li + NI * j;
      oc[j] = to * TO + lo + NO * j;
      ap_[j] = Ain + pb_min(ir[j], Ly.inA - 1) * pin + kslice * kchunk;
      dp_[j] = dZ + pb_min(oc[j], Ly.out - 1) * pz + kslice * kchunk;
    }
    // dot form: the (x,y) / (z,w) halves of the LDS.128 operands are natural FFMA2 pairs; each output keeps an
    // (even k, odd k) accumulator pair that is summed at the end.
    // (Measured on B200: 8x8 / 8x4 register tiles were tried for the wide layers and were 5-15 % SLOWER -- with 16
    // warps per CTA they leave too few warp tiles; 4x4 tiles + more warps win.)
    float2 acc2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc2[i][j] = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int b = 0; b < kchunk; b += 4) {
      float4 a[4], d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[j] = pb_ld4(ap_[j] + b);
        d[j] = pb_ld4(dp_[j] + b);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc2[i][j] = pb_fma2(make_float2(a[i].x, a[i].y), make_float2(d[j].x, d[j].y), acc2[i][j]);
          acc2[i][j] = pb_fma2(make_float2(a[i].z, a[i].w), make_float2(d[j].z, d[j].w), acc2[i][j]);
        }
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = acc2[i][j].x + acc2[i][j].y;
    float* dst = (kslice == 0 ? G : Gpart + (kslice - 1) * Ps) + Ly.soff;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (ir[i] < Ly.inA && oc[j] < Ly.out) {
          float* g = dst + ir[i] * Ly.wp + oc[j];
          *g = first ? acc[i][j] : (*g + acc[i][j]);
        }
  }
}

// warp tile: 16 input features (4 lanes interleaved by 4) x 32 batch rows (8 lanes x 4)
template <int ACT>
PB_DEV void pb_gemm_dh(const PbLayer& Ly, const float* W, const float* dZ, float* Ain, int ap, int BT, int tid,
                       int nt) {
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const int lm = lane & 7, ln = lane >> 3;
  const int tiles_b = BT >> 5, tiles_i = (Ly.in + 15) >> 4;
  const int Ko = (Ly.out + 3) & ~3;
  for (int wt = warp; wt < tiles_b * tiles_i; wt += nw) {
    const int ti = wt / tiles_b, tb = wt - ti * tiles_b;
    const int b0 = tb * 32 + lm * 4;
    int ir[4];
    const float* wr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ir[j] = ti * 16 + ln + 4 * j;
      wr[j] = W + pb_min(ir[j], Ly.in - 1) * Ly.wp;
    }
    const float* dptr = dZ + b0;
    float acc[4][4];
    PB_ZERO4x4(acc)
#pragma unroll 2
    for (int o = 0; o < Ko; o += 4) {
      float4 w[4], d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        w[j] = pb_ld4(wr[j] + o);
        d[j] = pb_ld4(dptr + (o + j) * ap);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] += w[i].x * d[0].x; acc[i][1] += w[i].x * d[0].y; acc[i][2] += w[i].x * d[0].z; acc[i][3] += w[i].x * d[0].w;
        acc[i][0] += w[i].y * d[1].x; acc[i][1] += w[i].y * d[1].y; acc[i][2] += w[i].y * d[1].z; acc[i][3] += w[i].y * d[1].w;
        acc[i][0] += w[i].z * d[2].x; acc[i][1] += w[i].z * d[2].y; acc[i][2] += w[i].z * d[2].z; acc[i][3] += w[i].z * d[2].w;
        acc[i][0] += w[i].w * d[3].x; acc[i][1] += w[i].w * d[3].y; acc[i][2] += w[i].w * d[3].z; acc[i][3] += w[i].w * d[3].w;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (ir[i] < Ly.in) {
        float* hp = Ain + ir[i] * ap + b0;
        const float4 h = pb_ld4(hp);
        float4 v;
        v.x = acc[i][0] * pb_actgrad<ACT>(h.x);
        v.y = acc[i][1] * pb_actgrad<ACT>(h.y);
        v.z = acc[i][2] * pb_actgrad<ACT>(h.z);
        v.w = acc[i][3] * pb_actgrad<ACT>(h.w);
        pb_st4(hp, v);
      }
  }
}

// ------------------------------------------------------------------------------------------------
// tile-level passes
// ------------------------------------------------------------------------------------------------

// zero everything behind the resident state arrays, then set the ones-rows
PB_DEV void pb_init_acts(const PbPlan& pl, float* sm, int nt) {
  PB_PHASE
  for (int i = pl.off_gpart + tid; i < pl.off_idx; i += nt) sm[i] = 0.f;
  PB_ENDPHASE
  if (pl.h1 || pl.tc2) return;  // fused / tcgen05 paths: no activation buffers, the loaders write the ones column
  PB_PHASE
  for (int l = 1; l < pl.nl; ++l) {
    float* row = pb_A(pl, sm, l) + pl.L[l].in * pl.ap;
    for (int b = tid; b < pl.BT; b += nt) row[b] = 1.f;
  }
  if (pl.res_rows > 0) {
    float* row = sm + pl.off_xres + pl.d * pl.xp;
    for (int b = tid; b < pl.xp; b += nt) row[b] = 1.f;
  } else {
    float* row = pb_A(pl, sm, 0) + pl.d * pl.ap;
    for (int b = tid; b < pl.BT; b += nt) row[b] = 1.f;
  }
  PB_ENDPHASE
}

// gather `nrows` rows, feature-major, into buffer (xdst, ydst) of pitch `pitch` with `cap` columns:
// xdst[k][b] = X[row(b)][k]; columns >= nrows are zeroed.  row(b) = idxs ? idxs[row0+b] : row0+b
PB_DEV void pb_gather_rows(const PbPlan& pl, float* xdst, float* ydst, int pitch, int cap, const float* X,
                           const float* y, const int* idxs, int row0, int nrows, int nt) {
  PB_TAG(10)
  PB_PHASE
  const int d = pl.d, s = pl.sy;
  for (int item = tid; item < cap * d; item += nt) {
    const int b = item / d, k = item - b * d;
    float v = 0.f;
    if (b < nrows) {
      const int row = idxs ? idxs[row0 + b] : row0 + b;
      v = X[(size_t)row * d + k];
    }
    xdst[k * pitch + b] = v;
  }
  if (y)
    for (int item = tid; item < cap * s; item += nt) {
      const int b = item / s, o = item - b * s;
      float v = 0.f;
      if (b < nrows) {
        const int row = idxs ? idxs[row0 + b] : row0 + b;
        v = y[(size_t)row * s + o];
      }
      ydst[o * pitch + b] = v;
    }
  PB_ENDPHASE
}

// make all rows of the evaluation resident (res plans) -- once for a fixed minibatch / the full data set
PB_DEV void pb_load_resident(const PbPlan& pl, float* sm, const float* X, const float* y, const int* idxs, int nrows,
                             int nt) {
  if (pl.h1) {
    // fused path: row-major X_aug[b][4*dp] = (x_b, 1, 0...), y[b]
    PB_TAG(10)
    PB_PHASE
    const int xpitch = 4 * pl.h1_dp, d = pl.d;
    float* Xr = sm + pl.off_xr;
    float* yr = sm + pl.off_yr;
    for (int item = tid; item < nrows * xpitch; item += nt) {
      const int b = item / xpitch, k = item - b * xpitch;
      float v = k == d ? 1.f : 0.f;
      if (k < d) v = X[(size_t)(idxs ? idxs[b] : b) * d + k];
      Xr[item] = v;
    }
    for (int b = tid; b < nrows; b += nt) yr[b] = y[(size_t)(idxs ? idxs[b] : b)];
    PB_ENDPHASE
    return;
  }
  pb_gather_rows(pl, sm + pl.off_xres, sm + pl.off_yres, pl.xp, pl.xp - 4, X, y, idxs, 0, nrows, nt);
}

// forward through all layers for one tile.  want_r: also form the scaled residual r = (y - f) * scale (the
// cotangent of the last layer) into A_nl and the per-row log-likelihood terms.
template <int ACT>
PB_DEV void pb_forward_tile(const PbPlan& pl, float* sm, const float* TH, const PbTile& T, int nvalid, float scale,
                            bool want_r, bool first, int nt) {
  const int ap = pl.ap, BT = pl.BT;
  for (int l = 0; l < pl.nl - 1; ++l) {
    PB_TAG(1)
    PB_PHASE
    pb_gemm_fwd<ACT, false>(pl.L[l], TH + pl.L[l].soff, l == 0 ? T.x : pb_A(pl, sm, l), l == 0 ? T.px : ap,
                            pb_A(pl, sm, l + 1), ap, BT, tid, nt);
    PB_ENDPHASE
  }
  const int ll_ = pl.nl - 1;
  const PbLayer& Ly = pl.L[ll_];
  const float* Ain = ll_ == 0 ? T.x : pb_A(pl, sm, ll_);
  const int pin = ll_ == 0 ? T.px : ap;
  float* Aout = pb_A(pl, sm, pl.nl);
  float* scr = pb_scr(pl, sm);
  float* llbuf = pb_llbuf(pl, sm);
  if (pl.s == 1) {
    // f[b] = sum_k W[k] * A[k][b]; the k-range is split over NT/BT thread groups
    const int KS = (nt / BT) > 0 ? (nt / BT) : 1;
    const int chunk = (Ly.inA + KS - 1) / KS;
    PB_TAG(2)
    PB_PHASE
    const float* W = TH + Ly.soff;
    for (int item = tid; item < BT * KS; item += nt) {
      const int ks = item / BT, b = item - ks * BT;
      const int k1 = pb_min(Ly.inA, (ks + 1) * chunk);
      float a0 = 0.f, a1 = 0.f;
      int k = ks * chunk;
      for (; k + 1 < k1; k += 2) {
        a0 += W[k * Ly.wp] * Ain[k * pin + b];
        a1 += W[(k + 1) * Ly.wp] * Ain[(k + 1) * pin + b];
      }
      if (k < k1) a0 += W[k * Ly.wp] * Ain[k * pin + b];
      scr[ks * BT + b] = a0 + a1;
    }
    PB_ENDPHASE
    PB_TAG(3)
    PB_PHASE
    float* fbuf = pb_fbuf(pl, sm);
    for (int b = tid; b < BT; b += nt) {
      float f = 0.f;
      for (int ks = 0; ks < KS; ++ks) f += scr[ks * BT + b];
      fbuf[b] = f;
      if (want_r) {
        const bool valid = b < nvalid;
        const float res = T.y[b] - f;
        Aout[b] = valid ? res * scale : 0.f;
        const float ll = valid ? -0.5f * res * res * pl.inv_s2 : 0.f;
        llbuf[b] = first ? ll : llbuf[b] + ll;
      }
    }
    PB_ENDPHASE
  } else {
    PB_PHASE
    pb_gemm_fwd<ACT, true>(Ly, TH + Ly.soff, Ain, pin, Aout, ap, BT, tid, nt);
    PB_ENDPHASE
    if (want_r) {
      PB_PHASE
      for (int b = tid; b < BT; b += nt) {
        const bool valid = b < nvalid;
        float ll = 0.f;
        if (pl.lik == PBNN_LIK_HETEROSCEDASTIC) {
          // outputs (mu, log sigma^2): ll = -0.5 ls - 0.5 (y - mu)^2 exp(-ls)   (logprobs.py:37-42)
          const float mu = Aout[b], ls = Aout[ap + b];
          const float e = expf(-ls), res = T.y[b] - mu;
          Aout[b] = valid ? res * e * scale : 0.f;
          Aout[ap + b] = valid ? (-0.5f + 0.5f * res * res * e) * scale : 0.f;
          ll = valid ? -0.5f * ls - 0.5f * res * res * e : 0.f;
        } else {
          for (int o = 0; o < pl.s; ++o) {
            const float res = T.y[o * T.py + b] - Aout[o * ap + b];
            Aout[o * ap + b] = valid ? res * scale : 0.f;
            ll += valid ? -0.5f * res * res * pl.inv_s2 : 0.f;
          }
        }
        llbuf[b] = first ? ll : llbuf[b] + ll;
      }
      PB_ENDPHASE
    }
  }
}

// backward through all layers; G accumulates sum_b r[b] * df_b/dtheta (first: overwrite).
template <int ACT>
PB_DEV void pb_backward_tile(const PbPlan& pl, float* sm, const float* TH, float* G, const PbTile& T, bool first,
                             int nt) {
  const int ap = pl.ap, BT = pl.BT;
  float* Gpart = sm + pl.off_gpart;
  for (int l = pl.nl - 1; l >= 0; --l) {
    const PbLayer& Ly = pl.L[l];
    const float* Ain = l == 0 ? T.x : pb_A(pl, sm, l);
    const int pin = l == 0 ? T.px : ap;
    const float* dZ = pb_A(pl, sm, l + 1);
    if (l == pl.nl - 1 && pl.s == 1) {
      // rank-1 last layer: G[k] += <r, A[k][:]> and (l > 0) A[k][:] <- r * W[k] * act'(A[k][:]) = dZ_{l-1}.
      // work item = (row k, column segment); consecutive lanes take consecutive rows (conflict-free LDS.128)
      int SEG = 1;
      while (SEG * 2 * Ly.inA <= nt && SEG * 2 * 4 <= BT && SEG < 8) SEG *= 2;
      const int seglen = BT / SEG;
      float* scr = pb_scr(pl, sm);
      PB_TAG(4)
      PB_PHASE
      for (int item = tid; item < Ly.inA * SEG; item += nt) {
        const int seg = item / Ly.inA, k = item - seg * Ly.inA;
        const float* rowc = Ain + k * pin + seg * seglen;
        float* roww = pb_A(pl, sm, l) + k * ap + seg * seglen;  // only dereferenced when l > 0
        const float* rp = dZ + seg * seglen;
        const float wk = TH[Ly.soff + k * Ly.wp];
        const bool prop = (l > 0) && (k < Ly.in);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int b = 0; b < seglen; b += 4) {
          const float4 r = pb_ld4(rp + b);
          const float4 h = pb_ld4(rowc + b);
          a0 += r.x * h.x; a1 += r.y * h.y; a2 += r.z * h.z; a3 += r.w * h.w;
          if (prop) {
            float4 v;
            v.x = r.x * wk * pb_actgrad<ACT>(h.x);
            v.y = r.y * wk * pb_actgrad<ACT>(h.y);
            v.z = r.z * wk * pb_actgrad<ACT>(h.z);
            v.w = r.w * wk * pb_actgrad<ACT>(h.w);
            pb_st4(roww + b, v);
          }
        }
        scr[seg * Ly.inA + k] = (a0 + a1) + (a2 + a3);
      }
      PB_ENDPHASE
      PB_TAG(5)
      PB_PHASE
      for (int k = tid; k < Ly.inA; k += nt) {
        float a = 0.f;
        for (int seg = 0; seg < SEG; ++seg) a += scr[seg * Ly.inA + k];
        float* g = G + Ly.soff + k * Ly.wp;
        *g = first ? a : (*g + a);
      }
      PB_ENDPHASE
    } else {
      PB_TAG(6)
      PB_PHASE
      if (Ly.dw_ni == 8) pb_gemm_dw<8>(Ly, Ain, pin, dZ, ap, G, Gpart, pl.Ps, BT, first, tid, nt);
      else pb_gemm_dw<4>(Ly, Ain, pin, dZ, ap, G, Gpart, pl.Ps, BT, first, tid, nt);
      PB_ENDPHASE
      if (l > 0) {
        PB_TAG(7)
        PB_PHASE
        pb_gemm_dh<ACT>(Ly, TH + Ly.soff, dZ, pb_A(pl, sm, l), ap, BT, tid, nt);
        PB_ENDPHASE
      }
    }
  }
}

// likelihood part of the gradient over `nrows` rows.  Resident plans: the rows must already be resident
// (pb_load_resident).  Streaming plans: tiles are gathered from global memory (idxs, or contiguous rows); a single
// already-loaded tile can be re-used with tile_loaded = true.
template <int ACT>
PB_DEV void pb_grad_rows(const PbPlan& pl, float* sm, const float* TH, float* G, const float* X, const float* y,
                         const int* idxs, int nrows, float scale, bool tile_loaded, int nt) {
  const int ntiles = (nrows + pl.BT - 1) / pl.BT;
  for (int tile = 0; tile < ntiles; ++tile) {
    const int nvalid = pb_min(pl.BT, nrows - tile * pl.BT);
    if (pl.res_rows == 0 && !tile_loaded)
      pb_gather_rows(pl, pb_A(pl, sm, 0), pb_A(pl, sm, pl.nl + 1), pl.ap, pl.BT, X, y, idxs, tile * pl.BT, nvalid, nt);
    const PbTile T = pb_tile(pl, sm, tile);
    pb_forward_tile<ACT>(pl, sm, TH, T, nvalid, scale, true, tile == 0, nt);
    pb_backward_tile<ACT>(pl, sm, TH, G, T, tile == 0, nt);
  }
  PB_TAG(9)
  if (pl.ksmax > 1) {
    PB_TAG(8)
    PB_PHASE
    const float* Gpart = sm + pl.off_gpart;
    for (int i = tid; i < pl.Ps; i += nt) {
      float a = G[i];
      for (int k = 0; k < pl.ksmax - 1; ++k) a += Gpart[k * pl.Ps + i];
      G[i] = a;
    }
    PB_ENDPHASE
  }
  PB_TAG(0)
}

#if !defined(PB_EMU)
#ifndef PB_H1_NP
#define PB_H1_NP 2  // rows in flight per warp in the fused path (register budget: 128 regs at 2 CTAs/SM)
#endif
PB_DEV float pb_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;  // xor butterfly: every lane ends with the bit-identical sum
}

// Fused gradient for one-hidden-layer nets with a scalar output (BASELINE configs 0 and 1: 8-50-1, 13-50-1).
// Warp-autonomous: each warp owns a contiguous slice of the rows; lane j keeps the input weights of hidden units
// 2j, 2j+1 and their gradient accumulators in registers.  Rows are processed NP at a time so that the NP warp
// all-reduces (the long dependent chains) overlap: per row broadcast LDS.128 of x, 2 dot products, one all-reduce
// for f, then x is re-read (broadcast, one wavefront) for the rank-1 gradient accumulation.  No activation ever
// touches shared memory and there is no CTA barrier until the per-warp partial gradients are summed in a fixed
// order (deterministic).  DP = float4 chunks per augmented input row (compile time: no branches in the row loop).
template <int ACT, int DP>
PB_DEV void pb_grad_h1_t(const PbPlan& pl, float* sm, const float* TH, float* G, int nrows, float scale, int nt) {
  PB_TAG(1)
  constexpr int NP = PB_H1_NP, KD = 4 * DP;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const PbLayer& L0 = pl.L[0];
  const PbLayer& L1 = pl.L[1];
  const int H = L0.out, wp = L0.wp;
  const float* Xr = sm + pl.off_xr;
  const float* yr = sm + pl.off_yr;
  float* hpart = sm + pl.off_hpart;
  float* part = hpart + warp * pl.Ps;
  float* llbuf = pb_llbuf(pl, sm);
  const int j0 = 2 * lane;
  const bool ok0 = j0 < H, ok1 = j0 + 1 < H;
  // packed FP32 (Blackwell fma.rn.f32x2 / FFMA2): register pairs (k, k+1) of the weights and accumulators, so the
  // x pairs coming out of the LDS.128 are used as they are -- half the FMA issue slots of scalar FFMA.
  // (Measured alternatives, B200: re-reading the weights per row block to afford NP = 4..8 rows in flight is 5-12 %
  // slower than weights in registers with NP = 2; NP = 4 with weights in registers spills and is 2x slower.)
  float2 W0[KD / 2], W1[KD / 2], a0[KD / 2], a1[KD / 2];
#pragma unroll
  for (int k = 0; k < KD / 2; ++k) {
    const float* wr0 = TH + L0.soff + (2 * k) * wp;
    const float* wr1 = wr0 + wp;
    W0[k] = make_float2(ok0 ? wr0[j0] : 0.f, ok0 ? wr1[j0] : 0.f);
    W1[k] = make_float2(ok1 ? wr0[j0 + 1] : 0.f, ok1 ? wr1[j0 + 1] : 0.f);
    a0[k] = make_float2(0.f, 0.f);
    a1[k] = make_float2(0.f, 0.f);
  }
  const float v0 = ok0 ? TH[L1.soff + j0 * L1.wp] : 0.f;
  const float v1 = ok1 ? TH[L1.soff + (j0 + 1) * L1.wp] : 0.f;
  const float b2 = TH[L1.soff + H * L1.wp];
  float av0 = 0.f, av1 = 0.f, ab = 0.f, ll = 0.f;
  const int per = (nrows + nw - 1) / nw;
  const int n0 = warp * per, n1 = pb_min(nrows, n0 + per);
  for (int n = n0; n < n1; n += NP) {
    float h0[NP], h1[NP], pf[NP];
    int row[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      row[p] = pb_min(n + p, n1 - 1);
      const float* xp_ = Xr + row[p] * KD;
      float2 z0 = make_float2(0.f, 0.f), z1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < DP; ++c) {
        const float4 t = pb_ld4(xp_ + 4 * c);
        const float2 lo = make_float2(t.x, t.y), hi = make_float2(t.z, t.w);
        z0 = __ffma2_rn(lo, W0[2 * c], z0);
        z1 = __ffma2_rn(lo, W1[2 * c], z1);
        z0 = __ffma2_rn(hi, W0[2 * c + 1], z0);
        z1 = __ffma2_rn(hi, W1[2 * c + 1], z1);
      }
      h0[p] = pb_act<ACT>(z0.x + z0.y);
      h1[p] = pb_act<ACT>(z1.x + z1.y);
      pf[p] = h0[p] * v0 + h1[p] * v1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int p = 0; p < NP; ++p) pf[p] += __shfl_xor_sync(0xffffffffu, pf[p], o);
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const bool valid = n + p < n1;
      const float res = yr[row[p]] - (pf[p] + b2);
      const float r = valid ? res * scale : 0.f;
      ll += valid ? -0.5f * res * res * pl.inv_s2 : 0.f;
      const float dz0 = r * v0 * pb_actgrad<ACT>(h0[p]), dz1 = r * v1 * pb_actgrad<ACT>(h1[p]);
      const float2 d0 = make_float2(dz0, dz0), d1 = make_float2(dz1, dz1);
      const float* xp_ = Xr + row[p] * KD;
#pragma unroll
      for (int c = 0; c < DP; ++c) {
        const float4 t = pb_ld4(xp_ + 4 * c);
        const float2 lo = make_float2(t.x, t.y), hi = make_float2(t.z, t.w);
        a0[2 * c] = __ffma2_rn(lo, d0, a0[2 * c]);
        a1[2 * c] = __ffma2_rn(lo, d1, a1[2 * c]);
        a0[2 * c + 1] = __ffma2_rn(hi, d0, a0[2 * c + 1]);
        a1[2 * c + 1] = __ffma2_rn(hi, d1, a1[2 * c + 1]);
      }
      av0 += r * h0[p];
      av1 += r * h1[p];
      ab += r;
    }
  }
#pragma unroll
  for (int k = 0; k < KD / 2; ++k) {
    if (2 * k < L0.inA) {
      if (ok0) part[L0.soff + (2 * k) * wp + j0] = a0[k].x;
      if (ok1) part[L0.soff + (2 * k) * wp + j0 + 1] = a1[k].x;
    }
    if (2 * k + 1 < L0.inA) {
      if (ok0) part[L0.soff + (2 * k + 1) * wp + j0] = a0[k].y;
      if (ok1) part[L0.soff + (2 * k + 1) * wp + j0 + 1] = a1[k].y;
    }
  }
  if (ok0) part[L1.soff + j0 * L1.wp] = av0;
  if (ok1) part[L1.soff + (j0 + 1) * L1.wp] = av1;
  if (lane == 0) {
    part[L1.soff + H * L1.wp] = ab;
    llbuf[warp] = ll;
  }
  if (tid >= nw && tid < 128) llbuf[tid] = 0.f;
  PB_ENDPHASE_RAW
  PB_TAG(8)
  for (int i = 4 * tid; i < pl.Ps; i += 4 * nt) {  // Ps is a multiple of 4; fixed summation order over the warps
    float4 a = pb_ld4(hpart + i);
    for (int w = 1; w < nw; ++w) {
      const float4 b = pb_ld4(hpart + w * pl.Ps + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    pb_st4(G + i, a);
  }
  PB_ENDPHASE_RAW
  PB_TAG(0)
}

template <int ACT>
PB_DEV void pb_grad_h1(const PbPlan& pl, float* sm, const float* TH, float* G, int nrows, float scale, int nt) {
  switch (pl.h1_dp) {
    case 1: pb_grad_h1_t<ACT, 1>(pl, sm, TH, G, nrows, scale, nt); break;
    case 2: pb_grad_h1_t<ACT, 2>(pl, sm, TH, G, nrows, scale, nt); break;
    case 3: pb_grad_h1_t<ACT, 3>(pl, sm, TH, G, nrows, scale, nt); break;
    default: pb_grad_h1_t<ACT, 4>(pl, sm, TH, G, nrows, scale, nt); break;
  }
}
#endif

#include "pbnn_tc.cuh"
#include "pbnn_tc2.cuh"

// likelihood gradient over the rows of one evaluation: fused one-hidden-layer path or the tiled GEMM engine
// ENG selects the gradient engine at COMPILE time (0 = tiled GEMM engine, 1 = fused one-hidden-layer path): separate
// kernels keep the code of each small (the HMC kernel lost 5 % when unrelated unrolled code was inlined into it).
template <int ACT, int ENG>
PB_DEV void pb_grad(const PbPlan& pl, float* sm, const float* TH, float* G, const float* X, const float* y,
                    const int* idxs, int nrows, float scale, bool tile_loaded, int nt) {
#if !defined(PB_EMU)
  if (ENG == 1) {
    pb_grad_h1<ACT>(pl, sm, TH, G, nrows, scale, nt);
    return;
  }
#endif
  pb_grad_rows<ACT>(pl, sm, TH, G, X, y, idxs, nrows, scale, tile_loaded, nt);
}

// deterministic block sums of K (<=3) quantities whose per-thread partials sit in scr[k*PB_MAX_NT + tid]
PB_DEV void pb_block_sums(const PbPlan& pl, float* sm, int K, float* dst, int nt) {
  const int nw = nt >> 5;
  float* scr = pb_scr(pl, sm);
  PB_PHASE
  if (tid < K * nw) {
    const int k = tid / nw, w = tid - k * nw;
    const float* p = scr + k * PB_MAX_NT + w * 32;
    float a = 0.f;
    for (int i = 0; i < 32; ++i) a += p[i];
    scr[PB_SCR_MAIN + k * 16 + w] = a;
  }
  PB_ENDPHASE
  PB_PHASE
  if (tid < K) {
    float a = 0.f;
    for (int w = 0; w < nw; ++w) a += scr[PB_SCR_MAIN + tid * 16 + w];
    dst[tid] = a;
  }
  PB_ENDPHASE
}

// flat [P] (global) -> resident padded layout; pads zeroed
PB_DEV void pb_load_state(const PbPlan& pl, float* dst, const float* src_flat, int nt) {
  PB_PHASE
  for (int i = tid; i < pl.Ps; i += nt) dst[i] = 0.f;
  PB_ENDPHASE
  if (src_flat) {
    PB_PHASE
    pb_for_each_param(pl, tid, nt, [&](int si, int fi) { dst[si] = src_flat[fi]; });
    PB_ENDPHASE
  }
}

PB_DEV void pb_store_state(const PbPlan& pl, const float* src, float* dst_flat, uint32_t* flagword, int nt) {
  PB_PHASE
  bool bad = false;
  pb_for_each_param(pl, tid, nt, [&](int si, int fi) {
    const float v = src[si];
    dst_flat[fi] = v;
    bad |= !pb_finite(v);
  });
  if (bad && flagword) PB_ATOMIC_OR(flagword, 1u);
  PB_ENDPHASE
}

// ------------------------------------------------------------------------------------------------
// minibatch index generation (src/pbnn/utils/data.py:69-78): key <- split(key)[1]; randint(key,(B,),0,N)
// ------------------------------------------------------------------------------------------------
PB_DEV void pb_gen_advance(const PbPlan& pl, float* sm, int times, int nt) {
  PB_PHASE
  if (tid == 0) {
    uint32_t* scu = pb_scu(pl, sm);
    PbKey k;
    k.k0 = scu[0];
    k.k1 = scu[1];
    for (int i = 0; i < times; ++i) k = pb_split_at(k, 1u);
    scu[0] = k.k0;
    scu[1] = k.k1;
  }
  PB_ENDPHASE
}

PB_DEV void pb_gen_indices(const PbPlan& pl, float* sm, int B, int N, uint32_t mult, int nt) {
  PB_PHASE
  const uint32_t* scu = pb_scu(pl, sm);
  int* idx = pb_idx(pl, sm);
  PbKey k;
  k.k0 = scu[0];
  k.k1 = scu[1];
  const PbKey k1 = pb_split_at(k, 0u), k2 = pb_split_at(k, 1u);
  for (int b = tid; b < B; b += nt) idx[b] = pb_randint_at(k1, k2, (uint32_t)b, (uint32_t)N, mult);
  PB_ENDPHASE
}

// caller-supplied indices are clamped to [0, N) like a JAX gather (an out-of-range index must not become an
// out-of-bounds read); bit 2 of the chain's flag word records that it happened
PB_DEV void pb_copy_indices(const PbPlan& pl, float* sm, const int32_t* src, int B, int N, int nt) {
  PB_PHASE
  int* idx = pb_idx(pl, sm);
  bool clamped = false;
  for (int b = tid; b < B; b += nt) {
    int v = src ? src[b] : b;  // nullptr: rows 0..B-1
    if (v < 0 || v >= N) {
      v = v < 0 ? 0 : N - 1;
      clamped = true;
    }
    idx[b] = v;
  }
  if (clamped) PB_ATOMIC_OR(pb_scu(pl, sm) + 2, 4u);
  PB_ENDPHASE
}

// ------------------------------------------------------------------------------------------------
// SG-MCMC / Adam chain body: one chain, T steps.
//   SGLD  : langevin.py:21-78 + blackjax.sgld          theta += eps*g + sqrt(2 eps) xi
//   pSGLD : langevin.py:81-140 + sgmcmc/diffusions.py:33-60
//   SGHMC : hamiltonian.py:80-140 + blackjax.sghmc
//   ADAM  : deep_ensembles.py:59-75 + optax.adam
//   GRAD  : test hook, one gradient evaluation
// resident arrays: st[0]=theta, st[1]=g, st[2]=aux (p | V | m), st[3]=aux2 (v)
// ------------------------------------------------------------------------------------------------
template <int ALGO, int ACT, bool GS, int ENG>
PB_DEV void pb_sg_chain(const PbPlan& pl, const PbSgArgs& a, float* sm, int c, int nt, void* ctx = nullptr) {
  // ENG: 0 tiled GEMM engine, 1 fused one-hidden-layer path, 2 two-hidden-layer tcgen05 engine (ctx = the CTA's PbTc2)
  float* ws = a.ws ? a.ws + (size_t)c * pl.nstate * pl.Ps : nullptr;
  float* TH = pb_state<GS>(pl, sm, ws, 0);
  float* G = pb_state<GS>(pl, sm, ws, 1);
  float* A1 = pb_state<GS>(pl, sm, ws, pl.nstate > 2 ? 2 : 0);
  float* A2 = pb_state<GS>(pl, sm, ws, pl.nstate > 3 ? 3 : 0);
  float* sc = pb_sc(pl, sm);
  uint32_t* scu = pb_scu(pl, sm);
  const int* sidx = pb_idx(pl, sm);
  const int P = pl.P, B = a.B;
  const size_t cP = (size_t)c * P;
  PbKey ck;
  ck.k0 = a.keys ? a.keys[2 * c] : 0u;
  ck.k1 = a.keys ? a.keys[2 * c + 1] : 0u;

  auto load_res = [&]() {
#if !defined(PB_EMU)
    if constexpr (ENG == 2) {
      pb_tc2_stage_x(pl, sm, *static_cast<PbTc2*>(ctx), a.X, a.y, sidx, B, nt);
      return;
    }
#endif
    pb_load_resident(pl, sm, a.X, a.y, sidx, B, nt);
  };
  auto grad = [&](bool tile_loaded_) {
    pb_grad<ACT, (ENG == 2 ? 0 : ENG)>(pl, sm, TH, G, a.X, a.y, sidx, B, a.scale, tile_loaded_, nt);
  };
#if !defined(PB_EMU)
  // ENG 2: gradient + parameter update in one pass (pbnn_tc2.cuh); the functors own the sampler arithmetic
  // (the engine supplies pre.a = theta from its operand images and writes the new theta to o0 / o1: trace row and
  // final state -- W1, b1, W2 have no fp32 copy in TH while the chain runs)
  auto t2grad = [&](PbKey nkey_, bool noise_, float* o0, float* o1, auto pf_, auto uf_) {
    pb_grad_tc2(pl, sm, TH, B, a.scale, nt, *static_cast<PbTc2*>(ctx), nkey_, noise_, o0, o1, scu + 2, pf_, uf_);
  };
  auto pf_theta = [&](int si, int fi) { return PbPre(); };
  auto trace_row = [&](int t_) -> float* {
    return (a.trace && t_ >= a.burnin && ((t_ - a.burnin) % a.thin) == 0)
               ? a.trace + ((size_t)c * a.T_out + (size_t)((t_ - a.burnin) / a.thin)) * P
               : nullptr;
  };
  auto final_dst = [&](int t_) -> float* { return (t_ == a.T - 1 && !a.no_theta_store) ? a.theta + cP : nullptr; };
#endif

  pb_init_acts(pl, sm, nt);
  pb_load_state(pl, TH, a.theta + cP, nt);
  if (ENG != 2 || ALGO == PB_ALGO_PSGLD) pb_load_state(pl, G, nullptr, nt);
  if (ALGO == PB_ALGO_PSGLD) {
    pb_load_state(pl, G, a.has_state ? a.aux1 + cP : nullptr, nt);
    pb_load_state(pl, A1, a.has_state ? a.aux2 + cP : nullptr, nt);
  }
  if (ALGO == PB_ALGO_SGHMC) pb_load_state(pl, A1, nullptr, nt);
  if (ALGO == PB_ALGO_ADAM) {
    pb_load_state(pl, A1, a.aux1 + cP, nt);
    pb_load_state(pl, A2, a.aux2 + cP, nt);
  }
  PB_PHASE
  if (tid == 0) {
    scu[0] = ck.k0;
    scu[1] = ck.k1;
    scu[2] = 0u;
  }
  PB_ENDPHASE
#if !defined(PB_EMU)
  if constexpr (ENG == 2) pb_tc2_restage(pl, sm, *static_cast<PbTc2*>(ctx), TH, nt);  // operand images of theta0
#endif

  const bool res = pl.res_rows > 0;
  const bool single_tile = B <= pl.BT;
  const bool explicit_idx = a.idx != nullptr || a.contig;
  const bool per_step =
      explicit_idx ? (a.idx_steps > 1 || ALGO == PB_ALGO_ADAM) : (a.mb_mode == PBNN_MB_PER_STEP);
  const int32_t* cidx = a.idx ? a.idx + (size_t)c * a.idx_steps * B : nullptr;
  int gen_pos = 0;  // how many draws of the generator have been consumed (uniform across threads)

  if (ALGO == PB_ALGO_GRAD) {
    pb_copy_indices(pl, sm, cidx, B, a.N, nt);
    if (res) load_res();
#if !defined(PB_EMU)
    if constexpr (ENG == 2) {
      t2grad(ck, false, nullptr, nullptr, pf_theta, [&](int si, int fi, const PbPre& v, float gll, float z) {
        a.aux1[cP + fi] = gll - v.a * pl.inv_pv;
        return v.a;
      });
    } else
#endif
      grad(false);
    PB_PHASE
    if (ENG != 2) pb_for_each_param(pl, tid, nt, [&](int si, int fi) { a.aux1[cP + fi] = G[si] - TH[si] * pl.inv_pv; });
    float* scr = pb_scr(pl, sm);
    const float* llbuf = pb_llbuf(pl, sm);
    float v = 0.f;
    for (int b = tid; b < pl.BT; b += nt) v += llbuf[b];
    scr[tid] = v;
    PB_ENDPHASE
    pb_block_sums(pl, sm, 1, sc + 8, nt);
    PB_PHASE
    if (tid == 0 && a.aux2) a.aux2[c] = sc[8];
    PB_ENDPHASE
    return;
  }

  // pSGLD init (psgld.py:25-29): g0 = grad(theta0, draw #init_draw), V0 = g0^2
  if (ALGO == PB_ALGO_PSGLD && !a.has_state) {
    if (explicit_idx) {
      pb_copy_indices(pl, sm, cidx, B, a.N, nt);
    } else {
      pb_gen_advance(pl, sm, a.init_draw - gen_pos, nt);
      gen_pos = a.init_draw;
      pb_gen_indices(pl, sm, B, a.N, a.randint_mult, nt);
    }
    if (res) load_res();
#if !defined(PB_EMU)
    if constexpr (ENG == 2) {
      t2grad(ck, false, nullptr, nullptr, pf_theta, [&](int si, int fi, const PbPre& v, float gll, float z) {
        const float g = gll - v.a * pl.inv_pv;
        G[si] = g;
        A1[si] = g * g;
        return v.a;
      });
    } else
#endif
    {
      grad(false);
      PB_PHASE
      for (int i = tid; i < pl.Ps; i += nt) {
        const float g = G[i] - TH[i] * pl.inv_pv;
        G[i] = g;
        A1[i] = g * g;
      }
      PB_ENDPHASE
    }
  }

  // the minibatch the traced loop re-uses (quirk Q1), or the state before the first per-step draw
  bool tile_loaded = false;
  if (!per_step) {
    if (explicit_idx) {
      pb_copy_indices(pl, sm, cidx, B, a.N, nt);
    } else {
      pb_gen_advance(pl, sm, a.which_draw - gen_pos, nt);
      gen_pos = a.which_draw;
      pb_gen_indices(pl, sm, B, a.N, a.randint_mult, nt);
    }
    if (res) {
      load_res();
    } else if (single_tile) {
      pb_gather_rows(pl, pb_A(pl, sm, 0), pb_A(pl, sm, pl.nl + 1), pl.ap, pl.BT, a.X, a.y, sidx, 0, B, nt);
      tile_loaded = true;
    }
  } else if (!explicit_idx) {
    // per-step mode: draw #1 is consumed by kern.init (langevin.py:60), step t uses draw t0 + t + 2
    const long long skip = 1 + a.t0 - gen_pos;
    pb_gen_advance(pl, sm, (int)skip, nt);
  }

  for (int t = 0; t < a.T; ++t) {
    const long long tg = a.t0 + t;
    const PbKey kt = a.direct_key ? ck : pb_split_at(ck, (uint32_t)tg);  // split(rng_key, T)[t]
    const float eps = a.step_sizes ? a.step_sizes[t < a.n_step_sizes ? t : a.n_step_sizes - 1] : a.const_step;
    if (per_step) {
      if (explicit_idx) {
        pb_copy_indices(pl, sm, cidx ? cidx + (size_t)t * B : nullptr, B, a.N, nt);
      } else {
        pb_gen_advance(pl, sm, 1, nt);
        pb_gen_indices(pl, sm, B, a.N, a.randint_mult, nt);
      }
      if (res) load_res();
    }

    if (ALGO == PB_ALGO_SGLD && ENG == 2) {
#if !defined(PB_EMU)
      if constexpr (ENG == 2) {
        const float ns = sqrtf(2.0f * a.temperature * eps);
        const float* cvc = a.cv_center ? a.cv_center + cP : nullptr;
        const float* cve = a.cv_center ? a.cv_center_est + cP : nullptr;
        if (cvc) {
          t2grad(
              kt, true, trace_row(t), final_dst(t),
              [&](int si, int fi) {
                PbPre v;
                v.c = cvc[fi];
                v.d = cve[fi];
                return v;
              },
              [&](int si, int fi, const PbPre& v, float gll, float z) {
                const float th = v.a;
                const float g = (v.c + (gll - th * pl.inv_pv)) - v.d;
                return th + eps * g + ns * z;
              });
        } else {
          t2grad(kt, true, trace_row(t), final_dst(t), pf_theta,
                 [&](int si, int fi, const PbPre& v, float gll, float z) {
                   const float th = v.a;
                   return th + eps * (gll - th * pl.inv_pv) + ns * z;
                 });
        }
      }
#endif
    } else if (ALGO == PB_ALGO_SGLD) {
      grad(tile_loaded);
      const float ns = sqrtf(2.0f * a.temperature * eps);
      const float* cvc = a.cv_center ? a.cv_center + cP : nullptr;
      const float* cve = a.cv_center ? a.cv_center_est + cP : nullptr;
      PB_PHASE
      pb_for_each_param_normal(
          pl, tid, nt, kt,
          [&](int si, int fi) {
            PbPre v;
            v.a = TH[si];
            v.b = G[si];
            v.c = cvc ? cvc[fi] : 0.f;
            v.d = cvc ? cve[fi] : 0.f;
            return v;
          },
          [&](int si, int fi, float z, const PbPre& v) {
            const float th = v.a;
            float g = v.b - th * pl.inv_pv;
            if (cvc) g = (v.c + g) - v.d;  // blackjax control_variates: grad_center + grad_est - grad_center_est
            TH[si] = th + eps * g + ns * z;
          });
      PB_ENDPHASE
    } else if (ALGO == PB_ALGO_PSGLD && ENG == 2) {
#if !defined(PB_EMU)
      if constexpr (ENG == 2) {
        // fused order: [theta update of step 0 by the generic pass] then per step: trace of theta_t, gradient at
        // theta_t with (g, V) update AND the theta update of step t + 1 in the drains (diffusions.py:42-59 unrolled)
        if (t == 0) {
          PB_PHASE
          pb_for_each_param_normal(
              pl, tid, nt, kt,
              [&](int si, int fi) {
                PbPre v;
                v.a = A1[si];
                v.b = TH[si];
                v.c = G[si];
                return v;
              },
              [&](int si, int fi, float z, const PbPre& v) {
                const float pre = pb_rcp_fast(pb_sqrt_fast(v.a) + 1e-6f);
                TH[si] = v.b + eps * pre * v.c + pb_sqrt_fast(2.0f * pre * eps) * z;
              });
          PB_ENDPHASE
          pb_tc2_restage(pl, sm, *static_cast<PbTc2*>(ctx), TH, nt);
        }
        if (t == 0 && trace_row(0)) {  // (later rows are written by the drains of the step before)
          float* dst = trace_row(0);
          PB_PHASE
          pb_for_each_param(pl, tid, nt, [&](int si, int fi) { dst[fi] = TH[si]; });
          PB_ENDPHASE
        }
        const bool more = t + 1 < a.T;
        const PbKey kn = a.direct_key ? ck : pb_split_at(ck, (uint32_t)(tg + 1));
        const float epsn =
            a.step_sizes ? a.step_sizes[t + 1 < a.n_step_sizes ? t + 1 : a.n_step_sizes - 1] : a.const_step;
        t2grad(
            kn, more, more ? trace_row(t + 1) : nullptr, final_dst(t),
            [&](int si, int fi) {
              PbPre v;
              v.b = A1[si];
              return v;
            },
            [&](int si, int fi, const PbPre& v, float gll, float z) {
              const float g = gll - v.a * pl.inv_pv;
              const float V = a.alpha * v.b + (1.0f - a.alpha) * (g * g);
              G[si] = g;
              A1[si] = V;
              if (!more) return v.a;
              const float pre = pb_rcp_fast(pb_sqrt_fast(V) + 1e-6f);
              return v.a + epsn * pre * g + pb_sqrt_fast(2.0f * pre * epsn) * z;
            });
        continue;  // (the trace of this step has been written above)
      }
#endif
    } else if (ALGO == PB_ALGO_PSGLD) {
      PB_PHASE
      pb_for_each_param_normal(
          pl, tid, nt, kt,
          [&](int si, int fi) {
            PbPre v;
            v.a = A1[si];
            v.b = TH[si];
            v.c = G[si];
            return v;
          },
          [&](int si, int fi, float z, const PbPre& v) {
            const float pre = pb_rcp_fast(pb_sqrt_fast(v.a) + 1e-6f);
            TH[si] = v.b + eps * pre * v.c + pb_sqrt_fast(2.0f * pre * eps) * z;
          });
      PB_ENDPHASE
      grad(tile_loaded);
      PB_PHASE
      pb_elementwise4(
          pl.Ps, tid, nt,
          [&](int i) {
            PbPre v;
            v.a = G[i];
            v.b = TH[i];
            v.c = A1[i];
            return v;
          },
          [&](int i, const PbPre& v) {
            const float g = v.a - v.b * pl.inv_pv;
            G[i] = g;
            A1[i] = a.alpha * v.c + (1.0f - a.alpha) * (g * g);
          });
      PB_ENDPHASE
    } else if (ALGO == PB_ALGO_SGHMC) {
      PB_PHASE
      // momentum resampling p = mu + sigma z (blackjax.sghmc: generate_gaussian_noise(rng_key, position, ...); which
      // of (mu, sigma) carries sqrt(step_size) is a parameter -- see DESIGN.md and include/pbnn_b200.h)
      const float seps = sqrtf(eps);
      const float mom_mu = a.mom_mu0 + a.mom_mu1 * seps, mom_sig = a.mom_sig0 + a.mom_sig1 * seps;
      pb_for_each_param_normal(
          pl, tid, nt, kt, [&](int si, int fi) { return PbPre(); },
          [&](int si, int fi, float z, const PbPre& v) { A1[si] = mom_mu + mom_sig * z; });
      PB_ENDPHASE
      const float ns = sqrtf(2.0f * eps * (a.alpha - a.beta) * a.temperature);
      const float* cvc = a.cv_center ? a.cv_center + cP : nullptr;
      const float* cve = a.cv_center ? a.cv_center_est + cP : nullptr;
      const float fric = 1.0f - a.alpha;
      for (int l = 0; l < a.L; ++l) {
        const PbKey kl = pb_split_at(kt, (uint32_t)l);  // split(keys[t], L)[l]
#if !defined(PB_EMU)
        if constexpr (ENG == 2) {
          if (cvc) {
            t2grad(
                kl, true, l == a.L - 1 ? trace_row(t) : nullptr, l == a.L - 1 ? final_dst(t) : nullptr,
                [&](int si, int fi) {
                  PbPre v;
                  v.b = A1[si];
                  v.d = cvc[fi];
                  v.e = cve[fi];
                  return v;
                },
                [&](int si, int fi, const PbPre& v, float gll, float z) {
                  const float th = v.a, p = v.b;
                  const float g = (v.d + (gll - th * pl.inv_pv)) - v.e;
                  A1[si] = fric * p + eps * g + ns * z;
                  return th + p;
                });
          } else {
            t2grad(
                kl, true, l == a.L - 1 ? trace_row(t) : nullptr, l == a.L - 1 ? final_dst(t) : nullptr,
                [&](int si, int fi) {
                  PbPre v;
                  v.b = A1[si];
                  return v;
                },
                [&](int si, int fi, const PbPre& v, float gll, float z) {
                  const float th = v.a, p = v.b;
                  A1[si] = fric * p + eps * (gll - th * pl.inv_pv) + ns * z;
                  return th + p;
                });
          }
          continue;
        }
#endif
        grad(tile_loaded);
        PB_PHASE
        pb_for_each_param_normal(
            pl, tid, nt, kl,
            [&](int si, int fi) {
              PbPre v;
              v.a = TH[si];
              v.b = A1[si];
              v.c = G[si];
              v.d = cvc ? cvc[fi] : 0.f;
              v.e = cvc ? cve[fi] : 0.f;
              return v;
            },
            [&](int si, int fi, float z, const PbPre& v) {
              const float th = v.a, p = v.b;
              float g = v.c - th * pl.inv_pv;
              if (cvc) g = (v.d + g) - v.e;
              TH[si] = th + p;
              A1[si] = fric * p + eps * g + ns * z;
            });
        PB_ENDPHASE
      }
    } else if (ALGO == PB_ALGO_ADAM && ENG == 2) {
#if !defined(PB_EMU)
      if constexpr (ENG == 2) {
        const float cnt = (float)(a.count0 + t + 1);
        const float ibc1 = 1.0f / (1.0f - powf(a.b1, cnt)), ibc2 = 1.0f / (1.0f - powf(a.b2, cnt));
        t2grad(
            ck, false, trace_row(t), final_dst(t),
            [&](int si, int fi) {
              PbPre p;
              p.c = A1[si];
              p.d = A2[si];
              return p;
            },
            [&](int si, int fi, const PbPre& p, float gll, float z) {
              const float th = p.a;
              const float g = -(gll - th * pl.inv_pv);  // gradient of the loss -logposterior
              const float m = (1.0f - a.b1) * g + a.b1 * p.c;
              const float v = (1.0f - a.b2) * (g * g) + a.b2 * p.d;
              A1[si] = m;
              A2[si] = v;
              return th + (-a.lr) * ((m * ibc1) * pb_rcp_fast(pb_sqrt_fast(v * ibc2) + a.eps));
            });
      }
#endif
    } else if (ALGO == PB_ALGO_ADAM) {
      grad(false);
      const float cnt = (float)(a.count0 + t + 1);
      const float ibc1 = 1.0f / (1.0f - powf(a.b1, cnt)), ibc2 = 1.0f / (1.0f - powf(a.b2, cnt));
      PB_PHASE
      pb_elementwise4(
          pl.Ps, tid, nt,
          [&](int i) {
            PbPre p;
            p.a = TH[i];
            p.b = G[i];
            p.c = A1[i];
            p.d = A2[i];
            return p;
          },
          [&](int i, const PbPre& p) {
            const float th = p.a;
            const float g = -(p.b - th * pl.inv_pv);  // gradient of the loss -logposterior
            const float m = (1.0f - a.b1) * g + a.b1 * p.c;
            const float v = (1.0f - a.b2) * (g * g) + a.b2 * p.d;
            A1[i] = m;
            A2[i] = v;
            TH[i] = th + (-a.lr) * ((m * ibc1) * pb_rcp_fast(pb_sqrt_fast(v * ibc2) + a.eps));
          });
      PB_ENDPHASE
    }

    if (ENG != 2 && a.trace && t >= a.burnin && ((t - a.burnin) % a.thin) == 0) {  // (ENG 2: written by the drains)
      float* dst = a.trace + ((size_t)c * a.T_out + (size_t)((t - a.burnin) / a.thin)) * P;
      PB_PHASE
      pb_for_each_param(pl, tid, nt, [&](int si, int fi) { dst[fi] = TH[si]; });
      PB_ENDPHASE
    }
  }

  // (ENG 2: the drains of the last gradient evaluation have stored the final state and raised the non-finite flag)
  if (ENG != 2 && !a.no_theta_store) pb_store_state(pl, TH, a.theta + cP, scu + 2, nt);
  if (ALGO == PB_ALGO_PSGLD || ALGO == PB_ALGO_ADAM) {
    pb_store_state(pl, ALGO == PB_ALGO_PSGLD ? G : A1, a.aux1 + cP, nullptr, nt);
    pb_store_state(pl, ALGO == PB_ALGO_PSGLD ? A1 : A2, a.aux2 + cP, nullptr, nt);
  }
  PB_PHASE
  if (tid == 0 && a.flags) a.flags[c] = scu[2];
  PB_ENDPHASE
}

// ------------------------------------------------------------------------------------------------
// HMC chain body (hamiltonian.py:18-77 + blackjax.hmc 1.2.5: velocity Verlet, one value_and_grad per
// leapfrog, Metropolis accept with one uniform; target = logprior + sum_i loglik_i, pipelines/hmc.py:59-62)
// resident arrays: st[0]=theta' (proposal), st[1]=grad', st[2]=momentum, st[3]=theta, st[4]=grad, st[5]=M^-1
// scalar slots: sc[8]=loglik sum, sc[9]=sum theta'^2, sc[10]=KE, sc[12]=logp(current), sc[13]=KE0,
//               scu[14]=accept flag, scu[15]=accept count
// ------------------------------------------------------------------------------------------------
PB_DEV void pb_hmc_logp_reduce(const PbPlan& pl, float* sm, const float* TH, const float* PM, const float* MI,
                               int nt) {
  PB_PHASE
  float* scr = pb_scr(pl, sm);
  const float* llbuf = pb_llbuf(pl, sm);
  float ll = 0.f, t2 = 0.f, ke = 0.f;
  for (int b = tid; b < pl.BT; b += nt) ll += llbuf[b];
  for (int i = tid; i < pl.Ps; i += nt) {
    t2 += TH[i] * TH[i];
    ke += (MI[i] * PM[i]) * PM[i];
  }
  scr[tid] = ll;
  scr[PB_MAX_NT + tid] = t2;
  scr[2 * PB_MAX_NT + tid] = ke;
  PB_ENDPHASE
  pb_block_sums(pl, sm, 3, pb_sc(pl, sm) + 8, nt);
}

template <int ACT, bool GS, int ENG>
PB_DEV void pb_hmc_chain(const PbPlan& pl, const PbHmcArgs& a, float* sm, int c, int nt, int s_begin = 0,
                         int s_end = -1, bool resume = false, void* tcx = nullptr) {
  // iterations [s_begin, s_end) of chain c; resume: continue a chain whose state another CTA left in global memory
  // (balanced schedule of the tcgen05 kernel); tcx: the CTA's tcgen05 context (ENG >= 2: set up by the kernel)
  if (s_end < 0) s_end = a.S;
  float* ws = a.ws ? a.ws + (size_t)c * pl.nstate * pl.Ps : nullptr;
  float* TH = pb_state<GS>(pl, sm, ws, 0);
  float* G = pb_state<GS>(pl, sm, ws, 1);
  float* PM = pb_state<GS>(pl, sm, ws, 2);
  float* TH0 = pb_state<GS>(pl, sm, ws, 3);
  float* G0 = pb_state<GS>(pl, sm, ws, 4);
  float* MI = pb_state<GS>(pl, sm, ws, 5);
  float* sc = pb_sc(pl, sm);
  uint32_t* scu = pb_scu(pl, sm);
  const int P = pl.P;
  const size_t cP = (size_t)c * P;
  PbKey ck;
  ck.k0 = a.keys ? a.keys[2 * c] : 0u;
  ck.k1 = a.keys ? a.keys[2 * c + 1] : 0u;
  const float eps = a.step_size, heps = a.step_size * 0.5f;

  // ENG >= 2: tcgen05 engine (pbnn_tc.cuh): buffers, operand staging and TMEM are set up once per CTA by the kernel
  if (ENG < 2) pb_init_acts(pl, sm, nt);
  pb_load_state(pl, TH, resume ? nullptr : a.theta + cP, nt);
  pb_load_state(pl, G, nullptr, nt);
  pb_load_state(pl, PM, nullptr, nt);
  pb_load_state(pl, MI, a.inv_mass, nt);
#if !defined(PB_EMU)
  PbTc tc_local;
  PbTc& tc = tcx ? *static_cast<PbTc*>(tcx) : tc_local;
#define PB_HMC_GRAD()                                                      \
  if constexpr (ENG >= 2) pb_grad_tc<ENG == 3>(pl, sm, TH, G, a.N, pl.inv_s2, nt, tc); \
  else pb_grad<ACT, ENG>(pl, sm, TH, G, a.X, a.y, nullptr, a.N, pl.inv_s2, false, nt);
#else
#define PB_HMC_GRAD() pb_grad<ACT, ENG>(pl, sm, TH, G, a.X, a.y, nullptr, a.N, pl.inv_s2, false, nt);
#endif
  if (ENG < 2 && pl.res_rows > 0) pb_load_resident(pl, sm, a.X, a.y, nullptr, a.N, nt);
  if (!resume) {
    PB_PHASE
    if (tid == 0) {
      scu[2] = 0u;
      scu[15] = 0u;
    }
    PB_ENDPHASE

    // init: (logp, grad) at theta0
    PB_HMC_GRAD()
    PB_PHASE
    for (int i = tid; i < pl.Ps; i += nt) {
      const float g = G[i] - TH[i] * pl.inv_pv;
      G[i] = g;
      G0[i] = g;
      TH0[i] = TH[i];
    }
    PB_ENDPHASE
    pb_hmc_logp_reduce(pl, sm, TH, PM, MI, nt);
    PB_PHASE
    if (tid == 0) sc[12] = (-0.5f * pl.inv_pv * sc[9] - (float)P * pl.logprior_const) + sc[8];
    if (a.grad_out) pb_for_each_param(pl, tid, nt, [&](int si, int fi) { a.grad_out[cP + fi] = G0[si]; });
    PB_ENDPHASE
  }
#if !defined(PB_EMU)
  else {
    // L2 loads (.cg): the previous owner of the chain published these with a release store (pb_hmc_tc_kernel)
    PB_PHASE
    pb_for_each_param(pl, tid, nt, [&](int si, int fi) {
      const float t = __ldcg(a.theta + cP + fi);
      TH[si] = t;
      TH0[si] = t;
      G0[si] = __ldcg(a.sched_g0 + cP + fi);
    });
    if (tid == 0) {
      sc[12] = __ldcg(a.sched_logp + c);
      scu[15] = (uint32_t)__ldcg(a.sched_acc + c);
      scu[2] = __ldcg(a.sched_flags + c);
    }
    PB_ENDPHASE
  }
#endif

  for (int s = s_begin; s < s_end; ++s) {
    const PbKey ks = pb_split_at(ck, (uint32_t)(a.t0 + s));
    const PbKey km = pb_split_at(ks, 0u), ka = pb_split_at(ks, 1u);
    // momentum ~ sqrt(1/Minv) * N(0, I); proposal starts at the current state
    PB_TAG(11)
    PB_PHASE
    float ke = 0.f;
    // (plain iterator: the momentum draw is 1/L of the work and the unrolled ILP variants cost the HMC kernel 5 %
    // through code size -- measured)
    pb_for_each_param(pl, tid, nt, [&](int si, int fi) {
      const float z = pb_normal_at(km, (uint32_t)fi);
      const float mi = MI[si];
      const float p = sqrtf(1.0f / mi) * z;
      PM[si] = p;
      ke += (mi * p) * p;
      TH[si] = TH0[si];
      G[si] = G0[si];
    });
    pb_scr(pl, sm)[tid] = ke;
    PB_ENDPHASE
    pb_block_sums(pl, sm, 1, sc + 13, nt);

#if !defined(PB_EMU)
    if constexpr (ENG >= 2) {
      // tcgen05 engine (ENG 2: N <= 512, ENG 3: larger N): half-kicks, drift and operand staging are fused into the gradient's readout (pbnn_tc.cuh)
      const PbTcLeap lf = {PM, MI, eps, heps, pl.inv_pv};
      if (a.L > 0) {
        pb_tc_stage_w<true>(pl, sm, TH, G, lf, nt, tc);
        for (int l = 0; l < a.L - 1; ++l) pb_tc_grad_core<1, ENG == 3>(pl, sm, TH, G, a.N, pl.inv_s2, nt, tc, lf);
        pb_tc_grad_core<2, ENG == 3>(pl, sm, TH, G, a.N, pl.inv_s2, nt, tc, lf);
      }
    } else
#endif
    for (int l = 0; l < a.L; ++l) {
      PB_TAG(12)
      PB_PHASE
      for (int i = tid; i < pl.Ps; i += nt) {
        const float p = PM[i] + heps * G[i];
        PM[i] = p;
        TH[i] = TH[i] + eps * (MI[i] * p);
      }
      PB_ENDPHASE
      PB_HMC_GRAD()
      PB_TAG(13)
      PB_PHASE
      for (int i = tid; i < pl.Ps; i += nt) {
        const float g = G[i] - TH[i] * pl.inv_pv;
        G[i] = g;
        PM[i] = PM[i] + heps * g;
      }
      PB_ENDPHASE
    }
    PB_TAG(14)
    pb_hmc_logp_reduce(pl, sm, TH, PM, MI, nt);
    PB_TAG(15)
    PB_PHASE
    if (tid == 0) {
      const float lp1 = (-0.5f * pl.inv_pv * sc[9] - (float)P * pl.logprior_const) + sc[8];
      const float e0 = -sc[12] + 0.5f * sc[13];
      const float e1 = -lp1 + 0.5f * sc[10];
      float delta = e0 - e1;
      if (delta != delta) delta = -INFINITY;
      const float pacc = fminf(1.0f, expf(delta));
      const float u = fmaxf(0.0f, pb_bits_to_uniform(pb_bits_at(ka, 0u)));
      const bool acc = u < pacc;
      scu[14] = acc ? 1u : 0u;
      if (acc) {
        sc[12] = lp1;
        scu[15] += 1u;
      }
      if (a.info_delta) a.info_delta[(size_t)c * a.S + s] = delta;
      if (a.info_accept) a.info_accept[(size_t)c * a.S + s] = acc ? 1 : 0;
    }
    PB_ENDPHASE
    const bool emit = a.trace && s >= a.burnin && ((s - a.burnin) % a.thin) == 0;
    float* dst = emit ? a.trace + ((size_t)c * a.S_out + (size_t)((s - a.burnin) / a.thin)) * P : nullptr;
    PB_TAG(16)
    PB_PHASE
    const bool acc = scu[14] != 0u;
    if (acc)
      for (int i = tid; i < pl.Ps; i += nt) {
        TH0[i] = TH[i];
        G0[i] = G[i];
      }
    if (emit) pb_for_each_param(pl, tid, nt, [&](int si, int fi) { dst[fi] = acc ? TH[si] : TH0[si]; });
    PB_ENDPHASE
  }

  if (!a.no_store) pb_store_state(pl, TH0, a.theta + cP, scu + 2, nt);
  PB_PHASE
  if (a.sched_g0 && s_end < a.S) {  // hand the chain over: gradient, log-density and counters of the current state
    pb_for_each_param(pl, tid, nt, [&](int si, int fi) { a.sched_g0[cP + fi] = G0[si]; });
    if (tid == 0) {
      a.sched_logp[c] = sc[12];
      a.sched_acc[c] = (int32_t)scu[15];
      a.sched_flags[c] = scu[2];
    }
  } else if (tid == 0) {
    if (a.logp) a.logp[c] = sc[12];
    if (a.accept_count) a.accept_count[c] = (int32_t)scu[15];
    if (a.flags) a.flags[c] = scu[2];
  }
  PB_ENDPHASE
#undef PB_HMC_GRAD
}

// ------------------------------------------------------------------------------------------------
// predictive: CTA (mt, ch) evaluates samples [ch*chunk, (ch+1)*chunk) on test rows [mt*BT, (mt+1)*BT)
// (predict_fn, langevin.py:75-76) and accumulates sum_s f, sum_s f^2 for the moments.
// ------------------------------------------------------------------------------------------------
template <int ACT, bool GS>
PB_DEV void pb_predict_block(const PbPlan& pl, const PbPredArgs& a, float* sm, int mt, int ch, int nt) {
  float* TH = pb_state<GS>(pl, sm, a.ws ? a.ws + ((size_t)mt * a.nchunk + ch) * pl.Ps : nullptr, 0);
  float* acc = pb_acc(pl, sm);
  const int ap = pl.ap, BT = pl.BT, s = pl.s, P = pl.P;
  const int row0 = mt * BT;
  const int nvalid = pb_min(BT, a.M - row0);
  const int accn = pb_r4dev(s) * ap;
  pb_init_acts(pl, sm, nt);
  pb_gather_rows(pl, pb_A(pl, sm, 0), nullptr, pl.ap, pl.BT, a.Xtest, nullptr, nullptr, row0, nvalid, nt);
  const PbTile T = pb_tile(pl, sm, 0);
  const int s0 = ch * a.chunk, s1 = pb_min(a.S, s0 + a.chunk);
  for (int smp = s0; smp < s1; ++smp) {
    pb_load_state(pl, TH, a.samples + (size_t)smp * P, nt);
    pb_forward_tile<ACT>(pl, sm, TH, T, nvalid, 0.f, false, true, nt);
    PB_PHASE
    const float* fbuf = pb_fbuf(pl, sm);
    const float* Aout = pb_A(pl, sm, pl.nl);
    for (int item = tid; item < s * BT; item += nt) {
      const int o = item / BT, b = item - o * BT;
      const float f = (s == 1) ? fbuf[b] : Aout[o * ap + b];
      acc[o * ap + b] += f;
      acc[accn + o * ap + b] += f * f;
      if (a.preds && b < nvalid) a.preds[((size_t)smp * a.M + row0 + b) * s + o] = f;
    }
    PB_ENDPHASE
  }
  if (a.part) {
    PB_PHASE
    for (int item = tid; item < s * BT; item += nt) {
      const int b = item / s, o = item - b * s;
      if (b < nvalid) {
        const size_t base = ((size_t)ch * 2) * a.M * s + (size_t)(row0 + b) * s + o;
        a.part[base] = acc[o * ap + b];
        a.part[base + (size_t)a.M * s] = acc[accn + o * ap + b];
      }
    }
    PB_ENDPHASE
  }
}
