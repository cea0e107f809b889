// pbnn_b200: Adam for small two-hidden-layer relu nets with a scalar output (BASELINE configs[3]: a deep ensemble of 512
// MLPs 8-50-50-1, B = 32) -- src/pbnn/deep_ensembles.py:59-75 (one optax.adam step per minibatch on the negative
// log-posterior) over src/pbnn/models.py:16-35 and benchmark/pipelines/logprobs.py:6-17; also the MAP trainer.
//
// Nets d-H-H-1 with H <= 64, d + 1 <= 16, B <= 32: the 50 x 50 layer is too small for a 128-row MMA tile (B = 32 leaves
// it three-quarters empty) and the tiled FP32 engine spends ~45 k instructions per step on it (4 x 4 register tiles for
// 32 x 50 outputs keep 4 of 8 warps busy, a parameter-sized gradient array and a separate update pass).  Here a CTA of
// FOUR warps owns a member for all its steps; the weights live in the CTA's shared memory, Adam's moments in global
// memory (L2), and a step alternates between two lane layouts:
//   row layout (lane = minibatch row)   forward of both layers and dZ1 = (dZ2 W2^T) o [Z1 > 0]: broadcast LDS.128 of
//                                       four weights x the lane's own activation (fma.rn.f32x2), no cross-lane
//                                       reduction anywhere; the four warps split the OUTPUT units of each phase;
//   unit layout (lane = output pair)    the weight gradients (reductions over the rows: rank-1 updates in registers,
//                                       8 input units per pass, the passes dealt over the warps), each followed at once by the
//                                       Adam update of exactly those weights -- no gradient array, no update pass.
// Four CTA barriers per step; the next step's minibatch rows arrive by cp.async behind the weight-gradient passes.
// Same arithmetic contract as the other engines (FP32, fixed summation order); a member's result does not depend on
// the launch it runs in.
#pragma once
#if !defined(PB_EMU)

#define PB_W2_NW 4  // warps per member

template <int DP>  // float4 chunks of an augmented input row
PB_DEV void pb_adam_w2_member(const PbPlan& pl, const PbSgArgs& a, float* sm, int c) {
  constexpr int KD = 4 * DP, KH = KD / 2, NW = PB_W2_NW;
  const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const PbLayer& L0 = pl.L[0];
  const PbLayer& L1 = pl.L[1];
  const PbLayer& L2 = pl.L[2];
  const int H = L0.out, d = pl.d, P = pl.P, B = a.B;
  const int JR = pl.w2_jr, PA = pl.w2_pa, nq = (H + 3) >> 2;
  float* const xs0 = sm + pl.w2_ox;   // [2][32][KD]  minibatch rows (x, 1, 0..) of this step / the next one, then y [2][32]
  float* const ys0 = xs0 + 2 * 32 * KD;
  float* const W1s = sm + pl.w2_ow1;  // [KD][64]  row k < d: kernel row, row d: bias, else 0
  float* const W2s = sm + pl.w2_ow2;  // [JR][64]  row j < H: kernel row, row H: bias, else 0
  float* const w3s = sm + pl.w2_ow3;  // [64] w3, then [NW][32] partial outputs, [32] residuals
  float* const fps = w3s + 64;
  float* const rs = fps + NW * 32;
  float* const H1s = sm + pl.w2_oh;   // [32][PA]  relu(Z1), column H = 1 (the bias input of layer 2)
  float* const S2s = sm + pl.w2_os;   // [32][PA]  relu(Z2)
  float* const Z2s = sm + pl.w2_oz;   // [32][PA]  dZ2 = r w3 [Z2 > 0]
  float* const D1s = sm + pl.w2_od;   // [32][PA]  dZ1
  const size_t cP = (size_t)c * P;
  float* const M = a.aux1 + cP;
  float* const V = a.aux2 + cP;
  const int n0 = 2 * lane;  // unit layout: this lane's pair of output units
  const bool ok0 = n0 < H, ok1 = n0 + 1 < H;
  const int n0c = n0 < JR ? n0 : 0;  // (lanes beyond the padded width read column 0 and discard)
  float* const h1row = H1s + lane * PA;  // row layout: this lane's rows
  float* const s2row = S2s + lane * PA;
  float* const z2row = Z2s + lane * PA;
  float* const d1row = D1s + lane * PA;
  uint32_t fl = 0u;

  // ---- member set-up: weights -> shared memory ----------------------------------------------------------------------
  for (int i = tid; i < pl.w2_warp_floats; i += 32 * NW) sm[i] = 0.f;
  __syncthreads();
  for (int e = tid; e < (d + 1) * H; e += 32 * NW) {
    const int k = e / H, j = e - k * H;
    W1s[k * 64 + j] = a.theta[cP + (k < d ? L0.fw + k * H + j : L0.fb + j)];
  }
  for (int e = tid; e < (H + 1) * H; e += 32 * NW) {
    const int j = e / H, n = e - j * H;
    W2s[j * 64 + n] = a.theta[cP + (j < H ? L1.fw + j * H + n : L1.fb + n)];
  }
  for (int n = tid; n < H; n += 32 * NW) w3s[n] = a.theta[cP + L2.fw + n];
  float b3 = a.theta[cP + L2.fb];  // (replicated in every thread: all apply the same update)
  __syncthreads();

  const int32_t* cidx = a.idx ? a.idx + (size_t)c * a.idx_steps * B : nullptr;
  const bool valid = lane < B;

  // Adam on one weight (optax.adam): gll = d loglik / d th.  Every phase issues ALL the moment loads of its weights
  // before its row loop and all the stores after the arithmetic (a load -> update -> store round trip per weight would
  // cost ~1 k cycles each: the stores may alias the next loads for the compiler).
  float ibc1 = 0.f, ibc2 = 0.f;
  auto adam = [&](float th, float gll, float& m, float& v, bool ok) -> float {
    const float g = -(gll - th * pl.inv_pv);  // gradient of the loss -logposterior
    m = (1.0f - a.b1) * g + a.b1 * m;
    v = (1.0f - a.b2) * (g * g) + a.b2 * v;
    // (sqrt.approx / rcp.approx: 2^-22 relative each, on a step of relative size lr -- the IEEE sqrtf + the range
    //  handling of __fdividef cost ~30 of the 45 instructions of this update, 3 051 times per step)
    float sq, rc;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(v * ibc2));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(sq + a.eps));
    const float tn = th + (-a.lr) * ((m * ibc1) * rc);
    return ok ? tn : th;
  };

  // Minibatch rows.  Warp 0 fetches them a step ahead with cp.async straight into shared memory (double-buffered), the
  // row INDEX two steps ahead: index -> row are two dependent L2 round trips (~3 k cycles) that no phase has to wait
  // for any more.  Indices are clamped like a JAX gather; bit 2 of the flag word records a clamp.
  auto fetch_index = [&](int t) -> int {
    int v = 0;
    if (valid && t < a.T) {
      v = a.contig ? lane : cidx[(size_t)(a.idx_steps > 1 ? t : 0) * B + lane];
      if (v < 0 || v >= a.N) {
        v = v < 0 ? 0 : a.N - 1;
        fl |= 4u;
      }
    }
    return v;
  };
  auto issue_row = [&](int v, int buf) {  // lane's row -> xs[buf][lane][0..d), y -> ys[buf][lane] (asynchronous)
    const float* xp = a.X + (size_t)v * d;
    const uint32_t dst = pb_smem_u32(xs0 + (buf * 32 + lane) * KD);
    if ((d & 3) == 0 && (reinterpret_cast<uintptr_t>(a.X) & 15) == 0) {
      for (int k = 0; k < d; k += 4)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 4u * (uint32_t)k), "l"(xp + k) : "memory");
    } else {
      for (int k = 0; k < d; ++k)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * (uint32_t)k), "l"(xp + k) : "memory");
    }
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(pb_smem_u32(ys0 + buf * 32 + lane)), "l"(a.y + v)
                 : "memory");
  };
  int vn = 0;
  if (warp == 0) {
    for (int buf = 0; buf < 2; ++buf)  // the constant tail of an augmented row: (1, 0, ..)
      for (int k = d; k < KD; ++k) xs0[(buf * 32 + lane) * KD + k] = k == d ? 1.f : 0.f;
    issue_row(fetch_index(0), 0);
    vn = fetch_index(1);
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();

  for (int t = 0; t < a.T; ++t) {
    if (tid == 0) {  // bias corrections of this step (published behind the first barrier)
      const float cnt = (float)(a.count0 + t + 1);
      rs[33] = 1.0f / (1.0f - powf(a.b1, cnt));
      rs[34] = 1.0f / (1.0f - powf(a.b2, cnt));
    }

    // ---- A. this step's minibatch row of the lane (every warp keeps it in registers), fetched during the previous step ---
    const float* const xs = xs0 + (t & 1) * 32 * KD;
    float xr[KD];
#pragma unroll
    for (int q = 0; q < DP; ++q) {
      const float4 x4 = pb_ld4(xs + lane * KD + 4 * q);
      xr[4 * q] = x4.x;
      xr[4 * q + 1] = x4.y;
      xr[4 * q + 2] = x4.z;
      xr[4 * q + 3] = x4.w;
    }
    const float yb = ys0[(t & 1) * 32 + lane];

    PB_MARK(1)
    // ---- B. layer 1 (row layout; warp w: output quads w, w + 4, ..): H1s[b][j] = relu(x_b . W1[:, j] + b1_j) -----------
    for (int q = warp; q < nq; q += NW) {
      const float* wq = W1s + 4 * q;
      float2 z01 = make_float2(0.f, 0.f), z23 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < KD; ++k) {  // (rows > d: x = 0 times w = 0)
        const float4 w = pb_ld4(wq + k * 64);
        z01 = __ffma2_rn(make_float2(xr[k], xr[k]), make_float2(w.x, w.y), z01);
        z23 = __ffma2_rn(make_float2(xr[k], xr[k]), make_float2(w.z, w.w), z23);
      }
      pb_st4(h1row + 4 * q, float4{fmaxf(z01.x, 0.f), fmaxf(z01.y, 0.f), fmaxf(z23.x, 0.f), fmaxf(z23.y, 0.f)});
    }
    if (warp == (H >> 2) % NW) h1row[H] = 1.f;  // bias input of layer 2: column H belongs to a quad of THIS warp (stored
                                                //  above, same thread) or to none
    __syncthreads();
    ibc1 = rs[33];
    ibc2 = rs[34];

    PB_MARK(2)
    // ---- C. layer 2 (row layout; warp w: output quads w, w + 4, w + 8, w + 12): S2s = relu(z2), partial w3 . relu(z2) ---
    {
      float2 acc[4][2];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u][0] = acc[u][1] = make_float2(0.f, 0.f);
      const float* wr = W2s + 4 * warp;
      const float* hp = h1row;
      const int nqw = (nq - warp + NW - 1) / NW;
      for (int jq = 0; jq < (JR >> 2); ++jq, hp += 4) {
        const float4 h4 = pb_ld4(hp);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj, wr += 64) {
          const float hv = jj == 0 ? h4.x : jj == 1 ? h4.y : jj == 2 ? h4.z : h4.w;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (u < nqw) {  // (warp-uniform: this warp's quads w, w + 4, .. below the net's width)
              const float4 w = pb_ld4(wr + 16 * u);
              acc[u][0] = __ffma2_rn(make_float2(hv, hv), make_float2(w.x, w.y), acc[u][0]);
              acc[u][1] = __ffma2_rn(make_float2(hv, hv), make_float2(w.z, w.w), acc[u][1]);
            }
        }
      }
      float fpart = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int q = warp + NW * u;
        if (q < nq) {
          const float4 w3 = pb_ld4(w3s + 4 * q);
          const float4 h2 = float4{fmaxf(acc[u][0].x, 0.f), fmaxf(acc[u][0].y, 0.f), fmaxf(acc[u][1].x, 0.f),
                                   fmaxf(acc[u][1].y, 0.f)};
          fpart = fmaf(h2.x, w3.x, fpart);
          fpart = fmaf(h2.y, w3.y, fpart);
          fpart = fmaf(h2.z, w3.z, fpart);
          fpart = fmaf(h2.w, w3.w, fpart);
          pb_st4(s2row + 4 * q, h2);
        }
      }
      fps[warp * 32 + lane] = fpart;
    }
    __syncthreads();

    PB_MARK(3)
    // ---- D. residual (every warp), dZ2 = r w3 [z2 > 0] for ALL units in registers (the next phase contracts over them);
    //         warp w publishes its own quads of it ------------------------------------------------------------------------
    const float f = ((fps[lane] + fps[32 + lane]) + fps[64 + lane]) + fps[96 + lane];
    const float r = valid ? (yb - (f + b3)) * a.scale : 0.f;
    if (warp == 0) rs[lane] = r;
    float2 dz[16][2];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      dz[q][0] = dz[q][1] = make_float2(0.f, 0.f);
      if (q < nq) {
        const float4 h2 = pb_ld4(s2row + 4 * q);
        const float4 w3 = pb_ld4(w3s + 4 * q);
        const float4 g = float4{h2.x > 0.f ? r * w3.x : 0.f, h2.y > 0.f ? r * w3.y : 0.f, h2.z > 0.f ? r * w3.z : 0.f,
                                h2.w > 0.f ? r * w3.w : 0.f};
        dz[q][0] = make_float2(g.x, g.y);
        dz[q][1] = make_float2(g.z, g.w);
        if ((q & (NW - 1)) == warp) pb_st4(z2row + 4 * q, g);
      }
    }

    PB_MARK(4)
    // ---- F. dZ1 (row layout; warp w: input quads w, w + 4, ..): D1s[b][j] = [h1_bj > 0] sum_n dZ2_bn W2[j][n] -------------
    for (int jq = warp; jq < nq; jq += NW) {
      const float4 h4 = pb_ld4(h1row + 4 * jq);
      float e4[4];
      const float* wr = W2s + (4 * jq) * 64;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj, wr += 64) {
        const float hv = jj == 0 ? h4.x : jj == 1 ? h4.y : jj == 2 ? h4.z : h4.w;
        float2 s0 = make_float2(0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;  // (four chains: 32 dependent FFMA2 otherwise)
#pragma unroll
        for (int q = 0; q < 16; q += 2) {
          const float4 w = pb_ld4(wr + 4 * q), w2 = pb_ld4(wr + 4 * q + 4);
          s0 = __ffma2_rn(make_float2(w.x, w.y), dz[q][0], s0);
          s1 = __ffma2_rn(make_float2(w.z, w.w), dz[q][1], s1);
          s2 = __ffma2_rn(make_float2(w2.x, w2.y), dz[q + 1][0], s2);
          s3 = __ffma2_rn(make_float2(w2.z, w2.w), dz[q + 1][1], s3);
        }
        const float2 s = __fadd2_rn(__fadd2_rn(s0, s1), __fadd2_rn(s2, s3));
        e4[jj] = (4 * jq + jj < H && hv > 0.f) ? s.x + s.y : 0.f;  // (row H of W2s is the bias, not an input unit)
      }
      pb_st4(d1row + 4 * jq, float4{e4[0], e4[1], e4[2], e4[3]});
    }
    __syncthreads();  // Z2s, D1s, rs complete; every read of the OLD W2 is done
    int vn2 = 0;
    if (warp == 0) {  // next step's rows and the index after that: in flight behind the weight-gradient passes
      issue_row(vn, (t + 1) & 1);
      vn2 = fetch_index(t + 2);
    }

    PB_MARK(5)
    // ---- G. dW2, db2 (unit layout, 8 input units per pass, pass p by warp p % NW) + Adam on exactly those weights --------
    for (int j0 = 8 * warp; j0 <= H; j0 += 8 * NW) {
      if (!ok0) continue;  // (lanes beyond the net's width own no weights; nothing below is warp-collective)
      float2 acc[8], mo[8], vo[8];
      const int o1 = ok1 ? 1 : 0;  // (odd H: the last lane's second unit does not exist -- it reads its first one twice)
#pragma unroll
      for (int i = 0; i < 8; ++i) {  // the moments of this pass: in flight during the row loop
        const int j = j0 + i <= H ? j0 + i : H;  // (rows beyond the bias row re-read it and never store)
        const int fi = (j < H ? L1.fw + j * H : L1.fb) + n0;
        mo[i] = make_float2(M[fi], M[fi + o1]);
        vo[i] = make_float2(V[fi], V[fi + o1]);
        acc[i] = make_float2(0.f, 0.f);
      }
      const bool two = j0 + 4 < JR;  // (the pass covers one or two quads of input units)
      const float* hp = H1s + j0;
      const float* zp = Z2s + n0;
#pragma unroll 4
      for (int b = 0; b < 32; ++b, hp += PA, zp += PA) {  // (unrolled: the loads of four rows are in flight together)
        const float2 d2 = *reinterpret_cast<const float2*>(zp);
        const float4 h4 = pb_ld4(hp);
        acc[0] = __ffma2_rn(make_float2(h4.x, h4.x), d2, acc[0]);
        acc[1] = __ffma2_rn(make_float2(h4.y, h4.y), d2, acc[1]);
        acc[2] = __ffma2_rn(make_float2(h4.z, h4.z), d2, acc[2]);
        acc[3] = __ffma2_rn(make_float2(h4.w, h4.w), d2, acc[3]);
        if (two) {
          const float4 g4 = pb_ld4(hp + 4);
          acc[4] = __ffma2_rn(make_float2(g4.x, g4.x), d2, acc[4]);
          acc[5] = __ffma2_rn(make_float2(g4.y, g4.y), d2, acc[5]);
          acc[6] = __ffma2_rn(make_float2(g4.z, g4.z), d2, acc[6]);
          acc[7] = __ffma2_rn(make_float2(g4.w, g4.w), d2, acc[7]);
        }
      }
      float* wp = W2s + j0 * 64 + n0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (j0 + i <= H) {  // (uniform)
          const int j = j0 + i;
          const int fi = (j < H ? L1.fw + j * H : L1.fb) + n0;
          float2 th = *reinterpret_cast<const float2*>(wp + i * 64);
          th.x = adam(th.x, acc[i].x, mo[i].x, vo[i].x, true);
          th.y = adam(th.y, acc[i].y, mo[i].y, vo[i].y, ok1);
          *reinterpret_cast<float2*>(wp + i * 64) = th;
          M[fi] = mo[i].x;
          V[fi] = vo[i].x;
          if (ok1) {
            M[fi + 1] = mo[i].y;
            V[fi + 1] = vo[i].y;
          }
        }
      }
    }

    PB_MARK(6)
    // ---- H. dW1, db1 (unit layout, warp NW - 1) and dw3, db3 (warp NW - 2: its last dW2 pass is the short one) + Adam ----
    if (warp == NW - 1) {
      float2 a0[KH], a1[KH], mo[KD], vo[KD];
#pragma unroll
      for (int k = 0; k < KH; ++k) a0[k] = a1[k] = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int fi = (k < d ? L0.fw + k * H : L0.fb) + n0;
        const int f0 = (ok0 && k <= d) ? fi : 0, f1 = (ok1 && k <= d) ? fi + 1 : 0;
        mo[k] = make_float2(M[f0], M[f1]);
        vo[k] = make_float2(V[f0], V[f1]);
      }
      const float* dp = D1s + n0c;
      const float* xp = xs;
#pragma unroll 4
      for (int b = 0; b < 32; ++b, dp += PA, xp += KD) {
        const float2 d1 = *reinterpret_cast<const float2*>(dp);
        const float2 dx = make_float2(d1.x, d1.x), dy = make_float2(d1.y, d1.y);
#pragma unroll
        for (int q = 0; q < DP; ++q) {
          const float4 x4 = pb_ld4(xp + 4 * q);
          a0[2 * q] = __ffma2_rn(make_float2(x4.x, x4.y), dx, a0[2 * q]);
          a1[2 * q] = __ffma2_rn(make_float2(x4.x, x4.y), dy, a1[2 * q]);
          a0[2 * q + 1] = __ffma2_rn(make_float2(x4.z, x4.w), dx, a0[2 * q + 1]);
          a1[2 * q + 1] = __ffma2_rn(make_float2(x4.z, x4.w), dy, a1[2 * q + 1]);
        }
      }
#pragma unroll
      for (int k = 0; k < KD; ++k)
        if (k <= d) {
          float2 th = *reinterpret_cast<const float2*>(W1s + k * 64 + n0);
          th.x = adam(th.x, (k & 1) ? a0[k >> 1].y : a0[k >> 1].x, mo[k].x, vo[k].x, ok0);
          th.y = adam(th.y, (k & 1) ? a1[k >> 1].y : a1[k >> 1].x, mo[k].y, vo[k].y, ok1);
          *reinterpret_cast<float2*>(W1s + k * 64 + n0) = th;
        }
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int fi = (k < d ? L0.fw + k * H : L0.fb) + n0;
        if (ok0 && k <= d) {
          M[fi] = mo[k].x;
          V[fi] = vo[k].x;
        }
        if (ok1 && k <= d) {
          M[fi + 1] = mo[k].y;
          V[fi + 1] = vo[k].y;
        }
      }
    }
    if (warp == NW - 2) {
      float m3[3], v3[3];
      m3[0] = M[ok0 ? L2.fw + n0 : 0];
      v3[0] = V[ok0 ? L2.fw + n0 : 0];
      m3[1] = M[ok1 ? L2.fw + n0 + 1 : 0];
      v3[1] = V[ok1 ? L2.fw + n0 + 1 : 0];
      m3[2] = M[L2.fb];
      v3[2] = V[L2.fb];
      float2 g3 = make_float2(0.f, 0.f);
      float gb3 = 0.f;
      const float* sp = S2s + n0c;
#pragma unroll 4
      for (int b = 0; b < 32; ++b, sp += PA) {
        const float2 h2 = *reinterpret_cast<const float2*>(sp);
        const float rb = rs[b];
        g3 = __ffma2_rn(make_float2(rb, rb), h2, g3);  // dw3_n = sum_b r_b relu(z2_bn)
        gb3 += rb;                                      // db3   = sum_b r_b
      }
      float2 th = *reinterpret_cast<const float2*>(w3s + n0);
      th.x = adam(th.x, g3.x, m3[0], v3[0], ok0);
      th.y = adam(th.y, g3.y, m3[1], v3[1], ok1);
      *reinterpret_cast<float2*>(w3s + n0) = th;
      const float b3n = adam(b3, gb3, m3[2], v3[2], true);
      if (lane == 0) rs[32] = b3n;  // (published to the other warps behind the barrier below)
      if (ok0) {
        M[L2.fw + n0] = m3[0];
        V[L2.fw + n0] = v3[0];
      }
      if (ok1) {
        M[L2.fw + n0 + 1] = m3[1];
        V[L2.fw + n0 + 1] = v3[1];
      }
      if (lane == 0) {
        M[L2.fb] = m3[2];
        V[L2.fb] = v3[2];
      }
    }
    if (warp == 0) {
      asm volatile("cp.async.wait_all;" ::: "memory");
      vn = vn2;
    }
    __syncthreads();  // the new weights and the next rows are visible to the next step
    b3 = rs[32];
    PB_MARK(7)
  }

  // ---- final state ----------------------------------------------------------------------------------------------------
  bool bad = false;
  if (!a.no_theta_store) {
    bad = !pb_finite(b3);
    for (int e = tid; e < (d + 1) * H; e += 32 * NW) {
      const int k = e / H, j = e - k * H;
      const float v = W1s[k * 64 + j];
      a.theta[cP + (k < d ? L0.fw + k * H + j : L0.fb + j)] = v;
      bad |= !pb_finite(v);
    }
    for (int e = tid; e < (H + 1) * H; e += 32 * NW) {
      const int j = e / H, n = e - j * H;
      const float v = W2s[j * 64 + n];
      a.theta[cP + (j < H ? L1.fw + j * H + n : L1.fb + n)] = v;
      bad |= !pb_finite(v);
    }
    for (int n = tid; n < H; n += 32 * NW) {
      const float v = w3s[n];
      a.theta[cP + L2.fw + n] = v;
      bad |= !pb_finite(v);
    }
    if (tid == 0) a.theta[cP + L2.fb] = b3;
  }
  fl |= bad ? 1u : 0u;
  if (a.flags && fl) atomicOr(a.flags + c, fl);
}

template <int DP>
__global__ void __launch_bounds__(32 * PB_W2_NW, 4)
pb_adam_w2_kernel(const __grid_constant__ PbPlan pl, const __grid_constant__ PbSgArgs a) {
  extern __shared__ __align__(16) float pb_smem[];
  const int c = (int)blockIdx.x;
#if defined(PB_PROFILE)
  if (blockIdx.x == 0 && threadIdx.x == 0) pb_prof_last = clock64();
#endif
  if (threadIdx.x == 0 && a.flags) a.flags[c] = 0u;
  __syncthreads();
  pb_adam_w2_member<DP>(pl, a, pb_smem, c);
}
#endif
