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
//   A_0..A_nl       activations, FEATURE-MAJOR: A_l[feature][b], pitch ap = BT+4; A_l has a constant
//                   ones-row at index dims[l] so the bias and its gradient fall out of the GEMMs
//   yT              targets, feature-major
// Three register-tiled FP32 GEMMs (4x4 outputs per thread, LDS.128 operands, conflict-free by pitch):
//   fwd : A_{l+1}[o][b] = act( sum_k W[k][o] * A_l[k][b] )            (k-major both operands)
//   dW  : G[i][o]      += sum_b A_l[i][b] * dZ[o][b]                 (dot form, both k-contiguous)
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
#define PB_ENDPHASE \
  }                 \
  __syncthreads();
#define PB_ATOMIC_OR(p, v) atomicOr((p), (v))
#endif

PB_DEV float4 pb_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
PB_DEV void pb_st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
PB_DEV int pb_min(int a, int b) { return a < b ? a : b; }
PB_DEV int pb_r4dev(int x) { return (x + 3) & ~3; }
PB_DEV bool pb_finite(float x) { return fabsf(x) <= 3.402823466e+38f; }

template <int ACT>
PB_DEV float pb_act(float z) {
  return ACT == PBNN_ACT_RELU ? fmaxf(z, 0.f) : tanhf(z);
}
template <int ACT>
PB_DEV float pb_actgrad(float h) {  // derivative expressed through the activation OUTPUT
  return ACT == PBNN_ACT_RELU ? (h > 0.f ? 1.f : 0.f) : 1.f - h * h;
}

struct PbSmem {
  float* st[8];
  float* A[PBNN_MAX_LAYERS + 2];
  float* yT;
  float* scr;    // [1024 + 64] reduction scratch
  float* llbuf;  // [128] per-row log-likelihood terms
  float* fbuf;   // [128] network outputs of the tile (s == 1)
  float* acc;    // [2 * r4(s) * ap] predictive accumulators
  int* idx;
  float* sc;     // 32 scalar slots
  uint32_t* scu;
};

PB_DEV PbSmem pb_carve(const PbPlan& pl, float* sm) {
  PbSmem S;
  for (int i = 0; i < 8; ++i) S.st[i] = sm + (i < pl.nstate ? i : 0) * pl.Ps;
  for (int l = 0; l <= pl.nl; ++l) S.A[l] = sm + pl.aoff[l];
  S.yT = sm + pl.aoff[pl.nl + 1];
  S.scr = sm + pl.off_scr;
  S.llbuf = S.scr + 1024 + 64;
  S.fbuf = S.llbuf + 128;
  S.acc = S.fbuf + 128;
  S.idx = reinterpret_cast<int*>(sm + pl.off_idx);
  S.sc = sm + pl.off_scal;
  S.scu = reinterpret_cast<uint32_t*>(sm + pl.off_scal);
  return S;
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

// ------------------------------------------------------------------------------------------------
// GEMMs (per-thread bodies; warp w owns warp-tiles w, w+NW, ...)
// ------------------------------------------------------------------------------------------------
#define PB_FMA4x4(acc, w, h)                                                         \
  acc[0][0] += w.x * h.x; acc[0][1] += w.x * h.y; acc[0][2] += w.x * h.z; acc[0][3] += w.x * h.w; \
  acc[1][0] += w.y * h.x; acc[1][1] += w.y * h.y; acc[1][2] += w.y * h.z; acc[1][3] += w.y * h.w; \
  acc[2][0] += w.z * h.x; acc[2][1] += w.z * h.y; acc[2][2] += w.z * h.z; acc[2][3] += w.z * h.w; \
  acc[3][0] += w.w * h.x; acc[3][1] += w.w * h.y; acc[3][2] += w.w * h.z; acc[3][3] += w.w * h.w;

// warp tile: 16 outputs (4 lane-groups x 4) x 32 batch rows (8 lane-groups x 4)
template <int ACT, bool IDENT>
PB_DEV void pb_gemm_fwd(const PbLayer& Ly, const float* W, const float* Ain, float* Aout, int ap, int BT, int tid,
                        int nt) {
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
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 w = pb_ld4(wptr + (k + kk) * wp);
        const float4 h = pb_ld4(aptr + (k + kk) * ap);
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
        pb_st4(Aout + o * ap + b0, v);
      }
    }
  }
}

#define PB_DOT4(a, d) (a.x * d.x + a.y * d.y + a.z * d.z + a.w * d.w)

// warp tile: 32 input rows (8 lanes, interleaved by 8) x 16 outputs (4 lanes, interleaved by 4)
PB_DEV void pb_gemm_dw(const PbLayer& Ly, const float* Ain, const float* dZ, float* G, int ap, int BT, bool first,
                       int tid, int nt) {
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const int li = lane & 7, lo = lane >> 3;
  const int tiles_i = (Ly.inA + 31) >> 5, tiles_o = (Ly.out + 15) >> 4;
  for (int wt = warp; wt < tiles_i * tiles_o; wt += nw) {
    const int ti = wt / tiles_o, to = wt - ti * tiles_o;
    const float* ap_[4];
    const float* dp_[4];
    int ir[4], oc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ir[j] = ti * 32 + li + 8 * j;
      oc[j] = to * 16 + lo + 4 * j;
      ap_[j] = Ain + pb_min(ir[j], Ly.inA - 1) * ap;
      dp_[j] = dZ + pb_min(oc[j], Ly.out - 1) * ap;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int b = 0; b < BT; b += 4) {
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
          acc[i][j] += a[i].x * d[j].x;
          acc[i][j] += a[i].y * d[j].y;
          acc[i][j] += a[i].z * d[j].z;
          acc[i][j] += a[i].w * d[j].w;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (ir[i] < Ly.inA && oc[j] < Ly.out) {
          float* g = G + Ly.soff + ir[i] * Ly.wp + oc[j];
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
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
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

// zero the activation region, set the ones-rows.  (pads of the resident arrays are zeroed by the loaders)
PB_DEV void pb_init_acts(const PbPlan& pl, const PbSmem& S, float* sm, int nt) {
  PB_PHASE
  for (int i = pl.off_act + tid; i < pl.off_idx; i += nt) sm[i] = 0.f;
  PB_ENDPHASE
  PB_PHASE
  for (int l = 0; l < pl.nl; ++l) {
    float* row = S.A[l] + pl.L[l].in * pl.ap;
    for (int b = tid; b < pl.BT; b += nt) row[b] = 1.f;
  }
  PB_ENDPHASE
}

// gather BT rows into the feature-major input tile: A_0[k][b] = X[row(b)][k]
PB_DEV void pb_load_tile(const PbPlan& pl, const PbSmem& S, const float* X, const float* y, const int* idxs,
                         int row0, int nvalid, bool with_y, int nt) {
  PB_PHASE
  const int d = pl.d, s = pl.s, ap = pl.ap;
  for (int item = tid; item < pl.BT * d; item += nt) {
    const int b = item / d, k = item - b * d;
    float v = 0.f;
    if (b < nvalid) {
      const int row = idxs ? idxs[row0 + b] : row0 + b;
      v = X[(size_t)row * d + k];
    }
    S.A[0][k * ap + b] = v;
  }
  if (with_y)
    for (int item = tid; item < pl.BT * s; item += nt) {
      const int b = item / s, o = item - b * s;
      float v = 0.f;
      if (b < nvalid) {
        const int row = idxs ? idxs[row0 + b] : row0 + b;
        v = y[(size_t)row * s + o];
      }
      S.yT[o * ap + b] = v;
    }
  PB_ENDPHASE
}

// forward through all layers for the resident tile.  want_r: also form the scaled residual
// r = (y - f) * scale (the cotangent of the last layer) into A_nl and the per-row log-likelihood terms.
template <int ACT>
PB_DEV void pb_forward_tile(const PbPlan& pl, const PbSmem& S, const float* TH, int nvalid, float scale, bool want_r,
                            bool first, int nt) {
  const int ap = pl.ap, BT = pl.BT;
  for (int l = 0; l < pl.nl - 1; ++l) {
    PB_PHASE
    pb_gemm_fwd<ACT, false>(pl.L[l], TH + pl.L[l].soff, S.A[l], S.A[l + 1], ap, BT, tid, nt);
    PB_ENDPHASE
  }
  const PbLayer& Ly = pl.L[pl.nl - 1];
  const float* Ain = S.A[pl.nl - 1];
  float* Aout = S.A[pl.nl];
  if (pl.s == 1) {
    // f[b] = sum_k W[k] * A[k][b]; the k-range is split over NT/BT thread groups
    const int KS = (nt / BT) > 0 ? (nt / BT) : 1;
    const int chunk = (Ly.inA + KS - 1) / KS;
    PB_PHASE
    const float* W = TH + Ly.soff;
    for (int item = tid; item < BT * KS; item += nt) {
      const int ks = item / BT, b = item - ks * BT;
      const int k1 = pb_min(Ly.inA, (ks + 1) * chunk);
      float a = 0.f;
      for (int k = ks * chunk; k < k1; ++k) a += W[k * Ly.wp] * Ain[k * ap + b];
      S.scr[ks * BT + b] = a;
    }
    PB_ENDPHASE
    PB_PHASE
    for (int b = tid; b < BT; b += nt) {
      float f = 0.f;
      for (int ks = 0; ks < KS; ++ks) f += S.scr[ks * BT + b];
      S.fbuf[b] = f;
      if (want_r) {
        const bool valid = b < nvalid;
        const float res = S.yT[b] - f;
        Aout[b] = valid ? res * scale : 0.f;
        const float ll = valid ? -0.5f * res * res * pl.inv_s2 : 0.f;
        S.llbuf[b] = first ? ll : S.llbuf[b] + ll;
      }
    }
    PB_ENDPHASE
  } else {
    PB_PHASE
    pb_gemm_fwd<ACT, true>(Ly, TH + Ly.soff, Ain, Aout, ap, BT, tid, nt);
    PB_ENDPHASE
    if (want_r) {
      PB_PHASE
      for (int b = tid; b < BT; b += nt) {
        const bool valid = b < nvalid;
        float ll = 0.f;
        for (int o = 0; o < pl.s; ++o) {
          const float res = S.yT[o * ap + b] - Aout[o * ap + b];
          Aout[o * ap + b] = valid ? res * scale : 0.f;
          ll += valid ? -0.5f * res * res * pl.inv_s2 : 0.f;
        }
        S.llbuf[b] = first ? ll : S.llbuf[b] + ll;
      }
      PB_ENDPHASE
    }
  }
}

// backward through all layers; G accumulates sum_b r[b] * df_b/dtheta (first: overwrite).
template <int ACT>
PB_DEV void pb_backward_tile(const PbPlan& pl, const PbSmem& S, const float* TH, float* G, bool first, int nt) {
  const int ap = pl.ap, BT = pl.BT;
  for (int l = pl.nl - 1; l >= 0; --l) {
    const PbLayer& Ly = pl.L[l];
    float* Ain = S.A[l];
    const float* dZ = S.A[l + 1];
    if (l == pl.nl - 1 && pl.s == 1) {
      // rank-1 last layer: one thread per input row k: G[k] += <r, A[k][:]>, then A[k][:] becomes dZ_{l-1}
      PB_PHASE
      for (int k = tid; k < Ly.inA; k += nt) {
        float* row = Ain + k * ap;
        const float wk = TH[Ly.soff + k * Ly.wp];
        const bool prop = (l > 0) && (k < Ly.in);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int b = 0; b < BT; b += 4) {
          const float4 r = pb_ld4(dZ + b);
          const float4 h = pb_ld4(row + b);
          a0 += r.x * h.x; a1 += r.y * h.y; a2 += r.z * h.z; a3 += r.w * h.w;
          if (prop) {
            float4 v;
            v.x = r.x * wk * pb_actgrad<ACT>(h.x);
            v.y = r.y * wk * pb_actgrad<ACT>(h.y);
            v.z = r.z * wk * pb_actgrad<ACT>(h.z);
            v.w = r.w * wk * pb_actgrad<ACT>(h.w);
            pb_st4(row + b, v);
          }
        }
        const float a = (a0 + a1) + (a2 + a3);
        float* g = G + Ly.soff + k * Ly.wp;
        *g = first ? a : (*g + a);
      }
      PB_ENDPHASE
    } else {
      PB_PHASE
      pb_gemm_dw(Ly, Ain, dZ, G, ap, BT, first, tid, nt);
      PB_ENDPHASE
      if (l > 0) {
        PB_PHASE
        pb_gemm_dh<ACT>(Ly, TH + Ly.soff, dZ, Ain, ap, BT, tid, nt);
        PB_ENDPHASE
      }
    }
  }
}

// likelihood part of the gradient over `nrows` rows (minibatch via idxs, or contiguous rows from row0).
// If `resident` the single tile already sits in A_0 / yT (fixed minibatch, nrows <= BT).
template <int ACT>
PB_DEV void pb_grad_rows(const PbPlan& pl, const PbSmem& S, const float* TH, float* G, const float* X, const float* y,
                         const int* idxs, int row0, int nrows, float scale, bool resident, int nt) {
  const int ntiles = (nrows + pl.BT - 1) / pl.BT;
  for (int tile = 0; tile < ntiles; ++tile) {
    const int nvalid = pb_min(pl.BT, nrows - tile * pl.BT);
    if (!resident) pb_load_tile(pl, S, X, y, idxs, row0 + tile * pl.BT, nvalid, true, nt);
    pb_forward_tile<ACT>(pl, S, TH, nvalid, scale, true, tile == 0, nt);
    pb_backward_tile<ACT>(pl, S, TH, G, tile == 0, nt);
  }
}

// deterministic block sums of K (<=4) quantities whose per-thread partials sit in scr[k*256 + tid]
PB_DEV void pb_block_sums(const PbSmem& S, int K, float* dst, int nt) {
  const int nw = nt >> 5;
  PB_PHASE
  if (tid < K * nw) {
    const int k = tid / nw, w = tid - k * nw;
    const float* p = S.scr + k * 256 + w * 32;
    float a = 0.f;
    for (int i = 0; i < 32; ++i) a += p[i];
    S.scr[1024 + k * 8 + w] = a;
  }
  PB_ENDPHASE
  PB_PHASE
  if (tid < K) {
    float a = 0.f;
    for (int w = 0; w < nw; ++w) a += S.scr[1024 + tid * 8 + w];
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
PB_DEV void pb_gen_advance(const PbSmem& S, int times, int nt) {
  PB_PHASE
  if (tid == 0) {
    PbKey k;
    k.k0 = S.scu[0];
    k.k1 = S.scu[1];
    for (int i = 0; i < times; ++i) k = pb_split_at(k, 1u);
    S.scu[0] = k.k0;
    S.scu[1] = k.k1;
  }
  PB_ENDPHASE
}

PB_DEV void pb_gen_indices(const PbSmem& S, int B, int N, uint32_t mult, int nt) {
  PB_PHASE
  PbKey k;
  k.k0 = S.scu[0];
  k.k1 = S.scu[1];
  const PbKey k1 = pb_split_at(k, 0u), k2 = pb_split_at(k, 1u);
  for (int b = tid; b < B; b += nt) S.idx[b] = pb_randint_at(k1, k2, (uint32_t)b, (uint32_t)N, mult);
  PB_ENDPHASE
}

PB_DEV void pb_copy_indices(const PbSmem& S, const int32_t* src, int B, int nt) {
  PB_PHASE
  for (int b = tid; b < B; b += nt) S.idx[b] = src ? src[b] : b;  // nullptr: rows 0..B-1
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
template <int ALGO, int ACT>
PB_DEV void pb_sg_chain(const PbPlan& pl, const PbSgArgs& a, float* sm, int c, int nt) {
  const PbSmem S = pb_carve(pl, sm);
  float* TH = S.st[0];
  float* G = S.st[1];
  float* A1 = S.st[2];
  float* A2 = S.st[3];
  const int P = pl.P, B = a.B;
  const size_t cP = (size_t)c * P;
  PbKey ck;
  ck.k0 = a.keys ? a.keys[2 * c] : 0u;
  ck.k1 = a.keys ? a.keys[2 * c + 1] : 0u;

  pb_init_acts(pl, S, sm, nt);
  pb_load_state(pl, TH, a.theta + cP, nt);
  pb_load_state(pl, G, nullptr, nt);
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
    S.scu[0] = ck.k0;
    S.scu[1] = ck.k1;
    S.scu[2] = 0u;
  }
  PB_ENDPHASE

  const bool single_tile = B <= pl.BT;
  const bool explicit_idx = a.idx != nullptr || a.contig;
  const bool per_step =
      explicit_idx ? (a.idx_steps > 1 || ALGO == PB_ALGO_ADAM) : (a.mb_mode == PBNN_MB_PER_STEP);
  const int32_t* cidx = a.idx ? a.idx + (size_t)c * a.idx_steps * B : nullptr;
  int gen_pos = 0;  // how many draws of the generator have been consumed (uniform across threads)

  if (ALGO == PB_ALGO_GRAD) {
    pb_copy_indices(S, cidx, B, nt);
    pb_grad_rows<ACT>(pl, S, TH, G, a.X, a.y, S.idx, 0, B, a.scale, false, nt);
    PB_PHASE
    pb_for_each_param(pl, tid, nt, [&](int si, int fi) { a.aux1[cP + fi] = G[si] - TH[si] * pl.inv_pv; });
    S.scr[tid] = 0.f;
    for (int b = tid; b < pl.BT; b += nt) S.scr[tid] += S.llbuf[b];
    PB_ENDPHASE
    pb_block_sums(S, 1, S.sc + 8, nt);
    PB_PHASE
    if (tid == 0 && a.aux2) a.aux2[c] = S.sc[8];
    PB_ENDPHASE
    return;
  }

  // pSGLD init (psgld.py:25-29): g0 = grad(theta0, draw #init_draw), V0 = g0^2
  if (ALGO == PB_ALGO_PSGLD && !a.has_state) {
    if (explicit_idx) {
      pb_copy_indices(S, cidx, B, nt);
    } else {
      pb_gen_advance(S, a.init_draw - gen_pos, nt);
      gen_pos = a.init_draw;
      pb_gen_indices(S, B, a.N, a.randint_mult, nt);
    }
    pb_grad_rows<ACT>(pl, S, TH, G, a.X, a.y, S.idx, 0, B, a.scale, false, nt);
    PB_PHASE
    for (int i = tid; i < pl.Ps; i += nt) {
      const float g = G[i] - TH[i] * pl.inv_pv;
      G[i] = g;
      A1[i] = g * g;
    }
    PB_ENDPHASE
  }

  // the minibatch the traced loop re-uses (quirk Q1), or the state before the first per-step draw
  if (!per_step) {
    if (explicit_idx) {
      pb_copy_indices(S, cidx, B, nt);
    } else {
      pb_gen_advance(S, a.which_draw - gen_pos, nt);
      gen_pos = a.which_draw;
      pb_gen_indices(S, B, a.N, a.randint_mult, nt);
    }
    if (single_tile) pb_load_tile(pl, S, a.X, a.y, S.idx, 0, B, true, nt);
  } else if (!explicit_idx) {
    // per-step mode: draw #1 is consumed by kern.init (langevin.py:60), step t uses draw t0 + t + 2
    const long long skip = 1 + a.t0 - gen_pos;
    pb_gen_advance(S, (int)skip, nt);
  }

  for (int t = 0; t < a.T; ++t) {
    const long long tg = a.t0 + t;
    const PbKey kt = a.direct_key ? ck : pb_split_at(ck, (uint32_t)tg);  // split(rng_key, T)[t]
    const float eps = a.step_sizes ? a.step_sizes[t < a.n_step_sizes ? t : a.n_step_sizes - 1] : a.const_step;
    if (per_step) {
      if (explicit_idx) {
        pb_copy_indices(S, cidx ? cidx + (size_t)t * B : nullptr, B, nt);
      } else {
        pb_gen_advance(S, 1, nt);
        pb_gen_indices(S, B, a.N, a.randint_mult, nt);
      }
    }
    const bool resident = single_tile && !per_step;

    if (ALGO == PB_ALGO_SGLD) {
      pb_grad_rows<ACT>(pl, S, TH, G, a.X, a.y, S.idx, 0, B, a.scale, resident, nt);
      const float ns = sqrtf(2.0f * eps);
      PB_PHASE
      pb_for_each_param(pl, tid, nt, [&](int si, int fi) {
        const float th = TH[si];
        const float g = G[si] - th * pl.inv_pv;
        TH[si] = th + eps * g + ns * pb_normal_at(kt, (uint32_t)fi);
      });
      PB_ENDPHASE
    } else if (ALGO == PB_ALGO_PSGLD) {
      PB_PHASE
      pb_for_each_param(pl, tid, nt, [&](int si, int fi) {
        const float pre = 1.0f / (sqrtf(A1[si]) + 1e-6f);
        TH[si] = TH[si] + eps * pre * G[si] + sqrtf(2.0f * pre * eps) * pb_normal_at(kt, (uint32_t)fi);
      });
      PB_ENDPHASE
      pb_grad_rows<ACT>(pl, S, TH, G, a.X, a.y, S.idx, 0, B, a.scale, resident, nt);
      PB_PHASE
      for (int i = tid; i < pl.Ps; i += nt) {
        const float g = G[i] - TH[i] * pl.inv_pv;
        G[i] = g;
        A1[i] = a.alpha * A1[i] + (1.0f - a.alpha) * (g * g);
      }
      PB_ENDPHASE
    } else if (ALGO == PB_ALGO_SGHMC) {
      PB_PHASE
      pb_for_each_param(pl, tid, nt, [&](int si, int fi) { A1[si] = pb_normal_at(kt, (uint32_t)fi); });
      PB_ENDPHASE
      const float ns = sqrtf(2.0f * eps * (a.alpha - a.beta));
      const float fric = 1.0f - a.alpha;
      for (int l = 0; l < a.L; ++l) {
        const PbKey kl = pb_split_at(kt, (uint32_t)l);  // split(keys[t], L)[l]
        pb_grad_rows<ACT>(pl, S, TH, G, a.X, a.y, S.idx, 0, B, a.scale, resident, nt);
        PB_PHASE
        pb_for_each_param(pl, tid, nt, [&](int si, int fi) {
          const float th = TH[si], p = A1[si];
          const float g = G[si] - th * pl.inv_pv;
          TH[si] = th + p;
          A1[si] = fric * p + eps * g + ns * pb_normal_at(kl, (uint32_t)fi);
        });
        PB_ENDPHASE
      }
    } else if (ALGO == PB_ALGO_ADAM) {
      pb_grad_rows<ACT>(pl, S, TH, G, a.X, a.y, S.idx, 0, B, a.scale, false, nt);
      const float cnt = (float)(a.count0 + t + 1);
      const float bc1 = 1.0f - powf(a.b1, cnt), bc2 = 1.0f - powf(a.b2, cnt);
      PB_PHASE
      for (int i = tid; i < pl.Ps; i += nt) {
        const float th = TH[i];
        const float g = -(G[i] - th * pl.inv_pv);  // gradient of the loss -logposterior
        const float m = (1.0f - a.b1) * g + a.b1 * A1[i];
        const float v = (1.0f - a.b2) * (g * g) + a.b2 * A2[i];
        A1[i] = m;
        A2[i] = v;
        TH[i] = th + (-a.lr) * ((m / bc1) / (sqrtf(v / bc2) + a.eps));
      }
      PB_ENDPHASE
    }

    if (a.trace && t >= a.burnin && ((t - a.burnin) % a.thin) == 0) {
      float* dst = a.trace + ((size_t)c * a.T_out + (size_t)((t - a.burnin) / a.thin)) * P;
      PB_PHASE
      pb_for_each_param(pl, tid, nt, [&](int si, int fi) { dst[fi] = TH[si]; });
      PB_ENDPHASE
    }
  }

  if (!a.no_theta_store) pb_store_state(pl, TH, a.theta + cP, S.scu + 2, nt);
  if (ALGO == PB_ALGO_PSGLD || ALGO == PB_ALGO_ADAM) {
    pb_store_state(pl, ALGO == PB_ALGO_PSGLD ? G : A1, a.aux1 + cP, nullptr, nt);
    pb_store_state(pl, ALGO == PB_ALGO_PSGLD ? A1 : A2, a.aux2 + cP, nullptr, nt);
  }
  PB_PHASE
  if (tid == 0 && a.flags) a.flags[c] = S.scu[2];
  PB_ENDPHASE
}

// ------------------------------------------------------------------------------------------------
// HMC chain body (hamiltonian.py:18-77 + blackjax.hmc 1.2.5: velocity Verlet, one value_and_grad per
// leapfrog, Metropolis accept with one uniform; target = logprior + sum_i loglik_i, pipelines/hmc.py:59-62)
// resident arrays: st[0]=theta' (proposal), st[1]=grad', st[2]=momentum, st[3]=theta, st[4]=grad, st[5]=M^-1
// scalar slots: sc[8]=loglik sum, sc[9]=sum theta'^2, sc[10]=KE, sc[12]=logp(current), sc[13]=KE0,
//               scu[14]=accept flag, scu[15]=accept count
// ------------------------------------------------------------------------------------------------
template <int ACT>
PB_DEV void pb_hmc_logp_reduce(const PbPlan& pl, const PbSmem& S, const float* TH, const float* PM, const float* MI,
                               int nt) {
  PB_PHASE
  float ll = 0.f, t2 = 0.f, ke = 0.f;
  for (int b = tid; b < pl.BT; b += nt) ll += S.llbuf[b];
  for (int i = tid; i < pl.Ps; i += nt) {
    t2 += TH[i] * TH[i];
    ke += (MI[i] * PM[i]) * PM[i];
  }
  S.scr[tid] = ll;
  S.scr[256 + tid] = t2;
  S.scr[512 + tid] = ke;
  PB_ENDPHASE
  pb_block_sums(S, 3, S.sc + 8, nt);
}

template <int ACT>
PB_DEV void pb_hmc_chain(const PbPlan& pl, const PbHmcArgs& a, float* sm, int c, int nt) {
  const PbSmem S = pb_carve(pl, sm);
  float* TH = S.st[0];
  float* G = S.st[1];
  float* PM = S.st[2];
  float* TH0 = S.st[3];
  float* G0 = S.st[4];
  float* MI = S.st[5];
  const int P = pl.P;
  const size_t cP = (size_t)c * P;
  PbKey ck;
  ck.k0 = a.keys[2 * c];
  ck.k1 = a.keys[2 * c + 1];
  const float eps = a.step_size, heps = a.step_size * 0.5f;

  pb_init_acts(pl, S, sm, nt);
  pb_load_state(pl, TH, a.theta + cP, nt);
  pb_load_state(pl, G, nullptr, nt);
  pb_load_state(pl, PM, nullptr, nt);
  pb_load_state(pl, MI, a.inv_mass, nt);
  PB_PHASE
  if (tid == 0) {
    S.scu[2] = 0u;
    S.scu[15] = 0u;
  }
  PB_ENDPHASE

  // init: (logp, grad) at theta0
  pb_grad_rows<ACT>(pl, S, TH, G, a.X, a.y, nullptr, 0, a.N, pl.inv_s2, false, nt);
  PB_PHASE
  for (int i = tid; i < pl.Ps; i += nt) {
    const float g = G[i] - TH[i] * pl.inv_pv;
    G[i] = g;
    G0[i] = g;
    TH0[i] = TH[i];
  }
  PB_ENDPHASE
  pb_hmc_logp_reduce<ACT>(pl, S, TH, PM, MI, nt);
  PB_PHASE
  if (tid == 0) S.sc[12] = (-0.5f * pl.inv_pv * S.sc[9] - (float)P * pl.logprior_const) + S.sc[8];
  PB_ENDPHASE

  for (int s = 0; s < a.S; ++s) {
    const PbKey ks = pb_split_at(ck, (uint32_t)(a.t0 + s));
    const PbKey km = pb_split_at(ks, 0u), ka = pb_split_at(ks, 1u);
    // momentum ~ sqrt(1/Minv) * N(0, I); proposal starts at the current state
    PB_PHASE
    float ke = 0.f;
    pb_for_each_param(pl, tid, nt, [&](int si, int fi) {
      const float mi = MI[si];
      const float p = sqrtf(1.0f / mi) * pb_normal_at(km, (uint32_t)fi);
      PM[si] = p;
      ke += (mi * p) * p;
      TH[si] = TH0[si];
      G[si] = G0[si];
    });
    S.scr[tid] = ke;
    PB_ENDPHASE
    pb_block_sums(S, 1, S.sc + 13, nt);

    for (int l = 0; l < a.L; ++l) {
      PB_PHASE
      for (int i = tid; i < pl.Ps; i += nt) {
        const float p = PM[i] + heps * G[i];
        PM[i] = p;
        TH[i] = TH[i] + eps * (MI[i] * p);
      }
      PB_ENDPHASE
      pb_grad_rows<ACT>(pl, S, TH, G, a.X, a.y, nullptr, 0, a.N, pl.inv_s2, false, nt);
      PB_PHASE
      for (int i = tid; i < pl.Ps; i += nt) {
        const float g = G[i] - TH[i] * pl.inv_pv;
        G[i] = g;
        PM[i] = PM[i] + heps * g;
      }
      PB_ENDPHASE
    }
    pb_hmc_logp_reduce<ACT>(pl, S, TH, PM, MI, nt);
    PB_PHASE
    if (tid == 0) {
      const float lp1 = (-0.5f * pl.inv_pv * S.sc[9] - (float)P * pl.logprior_const) + S.sc[8];
      const float e0 = -S.sc[12] + 0.5f * S.sc[13];
      const float e1 = -lp1 + 0.5f * S.sc[10];
      float delta = e0 - e1;
      if (delta != delta) delta = -INFINITY;
      const float pacc = fminf(1.0f, expf(delta));
      const float u = fmaxf(0.0f, pb_bits_to_uniform(pb_bits_at(ka, 0u)));
      const bool acc = u < pacc;
      S.scu[14] = acc ? 1u : 0u;
      if (acc) {
        S.sc[12] = lp1;
        S.scu[15] += 1u;
      }
      if (a.info_delta) a.info_delta[(size_t)c * a.S + s] = delta;
      if (a.info_accept) a.info_accept[(size_t)c * a.S + s] = acc ? 1 : 0;
    }
    PB_ENDPHASE
    const bool emit = a.trace && s >= a.burnin && ((s - a.burnin) % a.thin) == 0;
    float* dst = emit ? a.trace + ((size_t)c * a.S_out + (size_t)((s - a.burnin) / a.thin)) * P : nullptr;
    PB_PHASE
    const bool acc = S.scu[14] != 0u;
    if (acc)
      for (int i = tid; i < pl.Ps; i += nt) {
        TH0[i] = TH[i];
        G0[i] = G[i];
      }
    if (emit) pb_for_each_param(pl, tid, nt, [&](int si, int fi) { dst[fi] = acc ? TH[si] : TH0[si]; });
    PB_ENDPHASE
  }

  pb_store_state(pl, TH0, a.theta + cP, S.scu + 2, nt);
  PB_PHASE
  if (tid == 0) {
    if (a.logp) a.logp[c] = S.sc[12];
    if (a.accept_count) a.accept_count[c] = (int32_t)S.scu[15];
    if (a.flags) a.flags[c] = S.scu[2];
  }
  PB_ENDPHASE
}

// ------------------------------------------------------------------------------------------------
// predictive: CTA (mt, ch) evaluates samples [ch*chunk, (ch+1)*chunk) on test rows [mt*BT, (mt+1)*BT)
// (predict_fn, langevin.py:75-76) and accumulates sum_s f, sum_s f^2 for the moments.
// ------------------------------------------------------------------------------------------------
template <int ACT>
PB_DEV void pb_predict_block(const PbPlan& pl, const PbPredArgs& a, float* sm, int mt, int ch, int nt) {
  const PbSmem S = pb_carve(pl, sm);
  float* TH = S.st[0];
  const int ap = pl.ap, BT = pl.BT, s = pl.s, P = pl.P;
  const int row0 = mt * BT;
  const int nvalid = pb_min(BT, a.M - row0);
  const int accn = pb_r4dev(s) * ap;
  pb_init_acts(pl, S, sm, nt);
  PB_PHASE
  for (int i = tid; i < 2 * accn; i += nt) S.acc[i] = 0.f;
  PB_ENDPHASE
  pb_load_tile(pl, S, a.Xtest, nullptr, nullptr, row0, nvalid, false, nt);
  const int s0 = ch * a.chunk, s1 = pb_min(a.S, s0 + a.chunk);
  for (int smp = s0; smp < s1; ++smp) {
    pb_load_state(pl, TH, a.samples + (size_t)smp * P, nt);
    pb_forward_tile<ACT>(pl, S, TH, nvalid, 0.f, false, true, nt);
    PB_PHASE
    for (int item = tid; item < s * BT; item += nt) {
      const int o = item / BT, b = item - o * BT;
      const float f = (s == 1) ? S.fbuf[b] : S.A[pl.nl][o * ap + b];
      S.acc[o * ap + b] += f;
      S.acc[accn + o * ap + b] += f * f;
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
        a.part[base] = S.acc[o * ap + b];
        a.part[base + (size_t)a.M * s] = S.acc[accn + o * ap + b];
      }
    }
    PB_ENDPHASE
  }
}
