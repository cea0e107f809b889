// pbnn_b200: warp-per-chain engine for the SG samplers / Adam on one-hidden-layer nets with a scalar output
// (BASELINE configs 0 / 1 family: 8-50-1, 13-50-1; H <= 64, d + 1 <= 16, B <= 256) when there are many chains.
// Reference path: src/pbnn/mcmc/langevin.py:21-78 (sgld), :81-140 (pSGLD), src/pbnn/mcmc/hamiltonian.py:80-140 (sghmc),
// src/pbnn/deep_ensembles.py:59-75 (Adam) over src/pbnn/models.py:16-35 and benchmark/pipelines/logprobs.py:6-17.
//
// One WARP owns one chain for its whole life; nothing is shared between warps, so the kernel has no CTA barrier at all
// and an SM is kept busy by 16+ independent chains whose integer (threefry + erf_inv) and FP32 (gradient) phases
// interleave on the two pipes.  Lane j holds the first-layer weights (and bias) of hidden units 2j, 2j + 1, their output
// weights and a replica of the output bias IN REGISTERS for all T steps; the log-likelihood gradient of those weights
// is lane-local too (rank-1 updates in registers), so the only cross-lane traffic of a step is one xor-butterfly per
// data row (the network output) and the noise exchange:
//   * noise: jax.random.normal(key_t, (P,))[i] depends on (key_t, i) only, so lane l draws i = l, l + 32, ... -- every
//     lane does ceil(P / 32) threefry blocks instead of the 2 KD + 3 slots it owns (25 of 32 lanes hold weights at
//     H = 50) -- and scatters them through a per-warp table (flat index -> [slot][lane], built once per chain) into a
//     shared-memory array the owners read at immediate offsets; trace rows take the same road backwards, so the
//     global stores are whole 128-byte lines;
//   * auxiliary state (pSGLD g, V; SGHMC momentum; Adam m, v) sits in per-warp shared memory as [slot][lane].
// Same arithmetic contract as the other engines (FP32, fixed summation order, bit-exact noise); the chain's results do
// not depend on which warp / CTA runs it.
#pragma once
#if !defined(PB_EMU)

#define PB_W1_WPC PB_W1_WARPS  // warps (= chains) per CTA (pbnn_core.h)


// register slots of a lane: u < KD first-layer row k = u of unit 2 lane; KD <= u < 2 KD the same of unit 2 lane + 1
// (row d = bias, rows > d pads); 2 KD, 2 KD + 1 the two output weights; 2 KD + 2 the output bias (replicated in every
// lane: all lanes apply the identical update, stores of it are same-value same-address)
#define PB_W1_SLOT(W0_, W1_, v0_, v1_, b2_, u)                                                        \
  (*((u) < KD ? (((u) & 1) ? &W0_[((u) >> 1) % KH].y : &W0_[((u) >> 1) % KH].x)                        \
              : (u) < 2 * KD ? (((u) & 1) ? &W1_[(((u) - KD) >> 1) % KH].y : &W1_[(((u) - KD) >> 1) % KH].x) \
                             : (u) == 2 * KD ? &v0_ : (u) == 2 * KD + 1 ? &v1_ : &b2_))

// NWC warps per chain: 1 (many chains: every warp its own chain) or 4 (few chains, SGLD on relu nets: a CTA per chain --
// every warp keeps a replica of the weights and applies the identical update, the warps split the hidden units of the
// forward, the rows of the backward (partial gradients summed through shared memory) and the noise draws: a quarter of
// the step latency of a lone warp for BASELINE configs[0], one chain)
template <int ALGO, int ACT, int DP, int NWC = 1>
PB_DEV void pb_sgw_chain(const PbPlan& pl, const PbSgArgs& a, float* wsm, int c, int lane) {
  constexpr int KD = 4 * DP, KH = KD / 2, NS = 2 * KD + 3, NP = 2, CT = 32 * NWC;
  static_assert(NWC == 1 || (ALGO == PB_ALGO_SGLD && ACT == PBNN_ACT_RELU), "the CTA-per-chain variant: SGLD, relu");
  const int wc = NWC == 1 ? 0 : (int)(threadIdx.x >> 5);  // warp within the chain
  const int ctid = 32 * wc + lane;
  auto csync = [&]() {
    if constexpr (NWC == 1) __syncwarp();
    else __syncthreads();
  };
  const PbLayer& L0 = pl.L[0];
  const PbLayer& L1 = pl.L[1];
  const int H = L0.out, d = pl.d, P = pl.P, B = a.B;
  float* xs = wsm + pl.w1_ox;
  float* ys = wsm + pl.w1_oy;
  float* zs = wsm + pl.w1_oz;  // [NS + 1][32]: noise / trace staging in slot order (row NS: dump row for i >= P)
  uint16_t* tab = reinterpret_cast<uint16_t*>(wsm + pl.w1_ot);  // flat index i -> slot * 32 + owner lane
  int* ib = reinterpret_cast<int*>(wsm + pl.w1_oi);
  float* ax1 = wsm + pl.w1_oa + lane;  // [slot][lane]
  float* ax2 = ax1 + NS * 32;
  const int j0 = 2 * lane;
  const size_t cP = (size_t)c * P;
  uint32_t fl = 0u;
  PbKey ck;
  ck.k0 = a.keys ? a.keys[2 * c] : 0u;
  ck.k1 = a.keys ? a.keys[2 * c + 1] : 0u;

  // slot u -> (holds a parameter?, flat ravel_pytree index); straight-line code (u is a compile-time constant after
  // unrolling, the row test and the base offset are warp-uniform): pad slots answer index 0 and every memory access
  // they make is either harmless (loads) or predicated off (stores)
  const bool ok0 = j0 < H, ok1 = j0 + 1 < H;
  auto slot_ok = [&](int u) -> bool {
    if (u < 2 * KD) return ((u < KD ? u : u - KD) <= d) && (u < KD ? ok0 : ok1);
    return u == 2 * KD ? ok0 : (u == 2 * KD + 1 ? ok1 : true);
  };
  auto slot_fi = [&](int u) -> int {
    int fi;
    if (u < 2 * KD) {
      const int unit = u < KD ? 0 : 1, k = u - unit * KD;
      fi = (k < d ? L0.fw + k * H : L0.fb) + j0 + unit;
    } else {
      fi = u < 2 * KD + 2 ? L1.fw + j0 + (u - 2 * KD) : L1.fb;
    }
    return slot_ok(u) ? fi : 0;
  };

  float2 W0[KH], W1[KH], a0[KH], a1[KH];
  float v0, v1, b2, av0 = 0.f, av1 = 0.f, ab = 0.f;
#define TH_(u) PB_W1_SLOT(W0, W1, v0, v1, b2, u)
#define GA_(u) PB_W1_SLOT(a0, a1, av0, av1, ab, u)
#pragma unroll
  for (int u = 0; u < NS; ++u) {
    const float v = a.theta[cP + slot_fi(u)];
    TH_(u) = slot_ok(u) ? v : 0.f;
  }

  // flat index -> position in the slot-major staging array (the output bias lives in lane 0's copy; readers broadcast)
  for (int i = ctid; i < ((P + 4 * CT - 1) / (4 * CT)) * (4 * CT); i += CT) {
    int o = NS * 32 + lane;  // dump row
    if (i < P) {
      int u, j;
      if (i >= L1.fw) {
        j = i - L1.fw;
        u = 2 * KD + (j & 1);
      } else if (i == L1.fb) {
        j = 0;
        u = 2 * KD + 2;
      } else if (i >= L0.fw) {
        const int k = (i - L0.fw) / H;
        j = (i - L0.fw) - k * H;
        u = (j & 1) * KD + k;
      } else {
        j = i - L0.fb;
        u = (j & 1) * KD + d;
      }
      o = u * 32 + (j >> 1);
    }
    tab[i] = (uint16_t)o;
  }
  for (int i = ctid; i < (NS + 1) * 32; i += CT) zs[i] = 0.f;  // pad slots read 0 forever
  for (int i = ctid; i < 64; i += CT) (wsm + pl.w1_ow + KD * 64 + 64)[i] = 0.f;  // ballots of units never evaluated
  csync();
#define ZS_(u) zs[(u) * 32 + ((u) == NS - 1 ? 0 : lane)]

  // ---- minibatch ------------------------------------------------------------------------------------------------
  auto gather = [&]() {  // rows ib[0..B) -> xs[b][KD] = (x_b, 1, 0...), ys[b]
    csync();
    for (int item = ctid; item < B * KD; item += CT) {
      const int b = item / KD, k = item - b * KD;
      float v = k == d ? 1.f : 0.f;
      if (k < d) v = a.X[(size_t)ib[b] * d + k];
      xs[item] = v;
    }
    for (int b = ctid; b < B; b += CT) ys[b] = a.y[ib[b]];
    csync();
  };
  auto copy_indices = [&](const int32_t* src) {  // clamped like a JAX gather; bit 2 of the flag word records it
    csync();
    for (int b = ctid; b < B; b += CT) {
      int v = src ? src[b] : b;
      if (v < 0 || v >= a.N) {
        v = v < 0 ? 0 : a.N - 1;
        fl |= 4u;
      }
      ib[b] = v;
    }
  };
  PbKey gk = ck;  // minibatch generator key (utils/data.py:69-78): key <- split(key)[1]; randint(key, (B,), 0, N)
  auto gen_advance = [&](int times) {
    for (int i = 0; i < times; ++i) gk = pb_split_at(gk, 1u);
  };
  auto gen_indices = [&]() {
    csync();
    const PbKey k1 = pb_split_at(gk, 0u), k2 = pb_split_at(gk, 1u);
    for (int b = ctid; b < B; b += CT) ib[b] = pb_randint_at(k1, k2, (uint32_t)b, (uint32_t)a.N, a.randint_mult);
  };

  // ---- gradient of the (scaled) log-likelihood at the registers' theta: a0, a1, av0, av1, ab ----------------------
  // relu nets take two layouts per 32-row block.  FORWARD with lane = data row: the weights are published to shared memory
  // once per evaluation and read back as broadcasts, four hidden units per LDS.128, so a lane computes its own row's
  // output with no cross-lane reduction at all (the unit-per-lane layout needs a 5-stage butterfly per row), and the relu
  // pattern of a unit over the 32 rows is ONE ballot.  BACKWARD with lane = hidden-unit pair as before (the weight
  // gradient must end up next to the weights): per row a broadcast of the residual and of x, the unit's ballot bit
  // selects r or 0, rank-1 update in registers.  The output-layer gradient needs no pass over the rows:
  // sum_b r_b relu(z_bj) = sum_k W[k][j] (sum_b r_b m_bj x_bk), the same accumulators before the factor v_j is applied.
  auto grad = [&]() {
#pragma unroll
    for (int k = 0; k < KH; ++k) {
      a0[k] = make_float2(0.f, 0.f);
      a1[k] = make_float2(0.f, 0.f);
    }
    av0 = av1 = ab = 0.f;
    if constexpr (ACT == PBNN_ACT_RELU) {
      // [KD][64] first-layer rows (row d = bias, rows > d = the zero pads), then v[64], ballots[64], r[32]
      float* const ws = wsm + pl.w1_ow;
      float* const vs = ws + KD * 64;
      uint32_t* const mt = reinterpret_cast<uint32_t*>(vs + 64);
      float* const rs = vs + 128;
      float* const fps = rs + 32;                  // [NWC][32] partial outputs        (NWC > 1)
      float* const gps = fps + NWC * 32;           // [NWC][NS][32] partial gradients  (NWC > 1)
      csync();
      if (wc == 0) {  // (the replicas agree: one warp publishes)
#pragma unroll
        for (int k = 0; k < KD; ++k) *reinterpret_cast<float2*>(ws + k * 64 + j0) = make_float2(TH_(k), TH_(KD + k));
        *reinterpret_cast<float2*>(vs + j0) = make_float2(v0, v1);
      }
      const int nq = (H + 3) >> 2;
      float rsum = 0.f;
      for (int rb = 0; rb < B; rb += 32) {
        csync();
        const int row = rb + lane;
        const bool valid = row < B;
        const int rowc = valid ? row : B - 1;
        float4 xr[DP];
#pragma unroll
        for (int q = 0; q < DP; ++q) xr[q] = pb_ld4(xs + rowc * KD + 4 * q);
        float f = 0.f;
        const float* wq = ws + 4 * wc;
        // d a multiple of 4 (8-50-1): the bias is the only live row of the last float4 chunk -- it starts the sums
        // and the chunk is skipped (rows > d: x = 0 times w = 0 otherwise)
        const bool bias_first = KD > 4 && d == KD - 4;
        for (int q = wc; q < nq; q += NWC, wq += 4 * NWC) {  // (NWC > 1: the warps split the hidden units)
          float2 z01 = make_float2(0.f, 0.f), z23 = make_float2(0.f, 0.f);
          if (bias_first) {
            const float4 w = pb_ld4(wq + (KD - 4) * 64);
            z01 = make_float2(w.x, w.y);
            z23 = make_float2(w.z, w.w);
          }
#pragma unroll
          for (int k = 0; k < KD; ++k) {
            if (k >= KD - 4 && bias_first) break;
            const float4 w = pb_ld4(wq + k * 64);
            const float4 t = xr[k >> 2];
            const float xk = (k & 3) == 0 ? t.x : (k & 3) == 1 ? t.y : (k & 3) == 2 ? t.z : t.w;
            z01 = __ffma2_rn(make_float2(xk, xk), make_float2(w.x, w.y), z01);
            z23 = __ffma2_rn(make_float2(xk, xk), make_float2(w.z, w.w), z23);
          }
          const float4 v = pb_ld4(wq + KD * 64);
          f = fmaf(fmaxf(z01.x, 0.f), v.x, f);
          f = fmaf(fmaxf(z01.y, 0.f), v.y, f);
          f = fmaf(fmaxf(z23.x, 0.f), v.z, f);
          f = fmaf(fmaxf(z23.y, 0.f), v.w, f);
          const uint32_t m_0 = __ballot_sync(0xffffffffu, z01.x > 0.f), m_1 = __ballot_sync(0xffffffffu, z01.y > 0.f);
          const uint32_t m_2 = __ballot_sync(0xffffffffu, z23.x > 0.f), m_3 = __ballot_sync(0xffffffffu, z23.y > 0.f);
          if (lane == 0) *reinterpret_cast<uint4*>(mt + 4 * q) = make_uint4(m_0, m_1, m_2, m_3);
        }
        if constexpr (NWC > 1) {
          fps[wc * 32 + lane] = f;
          __syncthreads();
          f = 0.f;
#pragma unroll
          for (int w = 0; w < NWC; ++w) f += fps[w * 32 + lane];  // (fixed order: every warp gets the same bits)
        }
        const float res = ys[rowc] - (f + b2);
        const float r = valid ? res * a.scale : 0.f;
        if (wc == 0) rs[lane] = r;
        rsum += r;
        csync();
        const uint32_t m0 = mt[j0], m1 = mt[j0 + 1];
        const int nr = pb_min(32, B - rb);
        for (int bb = wc * (32 / NWC); bb < pb_min(nr, (wc + 1) * (32 / NWC)); ++bb) {  // (NWC > 1: the warps split the rows)
          const float rr = rs[bb];
          const uint32_t bit = 1u << bb;
          const float d0s = (m0 & bit) ? rr : 0.f, d1s = (m1 & bit) ? rr : 0.f;
          const float2 d0 = make_float2(d0s, d0s), d1 = make_float2(d1s, d1s);
          const float* xp_ = xs + (rb + bb) * KD;
#pragma unroll
          for (int q = 0; q < DP; ++q) {
            const float4 t = pb_ld4(xp_ + 4 * q);
            const float2 lo = make_float2(t.x, t.y), hi = make_float2(t.z, t.w);
            a0[2 * q] = __ffma2_rn(lo, d0, a0[2 * q]);
            a1[2 * q] = __ffma2_rn(lo, d1, a1[2 * q]);
            a0[2 * q + 1] = __ffma2_rn(hi, d0, a0[2 * q + 1]);
            a1[2 * q + 1] = __ffma2_rn(hi, d1, a1[2 * q + 1]);
          }
        }
      }
      ab = pb_warp_sum(rsum);
      if constexpr (NWC > 1) {  // partial first-layer gradients of the warps -> every warp sums all of them, same order
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KH; ++k) {
          *reinterpret_cast<float2*>(gps + ((wc * KH + k) * 32 + lane) * 4) = a0[k];
          *reinterpret_cast<float2*>(gps + ((wc * KH + k) * 32 + lane) * 4 + 2) = a1[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KH; ++k) {
          float2 t0 = make_float2(0.f, 0.f), t1 = make_float2(0.f, 0.f);
#pragma unroll
          for (int w = 0; w < NWC; ++w) {
            t0 = __fadd2_rn(t0, *reinterpret_cast<const float2*>(gps + ((w * KH + k) * 32 + lane) * 4));
            t1 = __fadd2_rn(t1, *reinterpret_cast<const float2*>(gps + ((w * KH + k) * 32 + lane) * 4 + 2));
          }
          a0[k] = t0;
          a1[k] = t1;
        }
      }
      float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < KH; ++k) {
        s0 = __ffma2_rn(W0[k], a0[k], s0);
        s1 = __ffma2_rn(W1[k], a1[k], s1);
      }
      av0 = s0.x + s0.y;
      av1 = s1.x + s1.y;
#pragma unroll
      for (int k = 0; k < KH; ++k) {
        a0[k] = make_float2(a0[k].x * v0, a0[k].y * v0);
        a1[k] = make_float2(a1[k].x * v1, a1[k].y * v1);
      }
    } else {
    for (int n = 0; n < B; n += NP) {
      float h0[NP], h1[NP], pf[NP];
      float4 xr[NP][DP];  // the rows stay in registers for the rank-1 update below
      int row[NP];
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        row[p] = pb_min(n + p, B - 1);
        const float* xp_ = xs + row[p] * KD;
        float2 z0 = make_float2(0.f, 0.f), z1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < DP; ++q) {
          const float4 t = xr[p][q] = pb_ld4(xp_ + 4 * q);
          const float2 lo = make_float2(t.x, t.y), hi = make_float2(t.z, t.w);
          z0 = __ffma2_rn(lo, W0[2 * q], z0);
          z1 = __ffma2_rn(lo, W1[2 * q], z1);
          z0 = __ffma2_rn(hi, W0[2 * q + 1], z0);
          z1 = __ffma2_rn(hi, W1[2 * q + 1], z1);
        }
        h0[p] = pb_act<ACT>(z0.x + z0.y);
        h1[p] = pb_act<ACT>(z1.x + z1.y);
        pf[p] = h0[p] * v0 + h1[p] * v1;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {  // both rows' butterflies share one packed add per stage
        static_assert(NP == 2, "packed butterfly");
        const float2 got = make_float2(__shfl_xor_sync(0xffffffffu, pf[0], o), __shfl_xor_sync(0xffffffffu, pf[1], o));
        const float2 sum = __fadd2_rn(make_float2(pf[0], pf[1]), got);
        pf[0] = sum.x;
        pf[1] = sum.y;
      }
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const bool valid = n + p < B;
        const float res = ys[row[p]] - (pf[p] + b2);
        const float r = valid ? res * a.scale : 0.f;
        const float dz0 = r * v0 * pb_actgrad<ACT>(h0[p]), dz1 = r * v1 * pb_actgrad<ACT>(h1[p]);
        const float2 d0 = make_float2(dz0, dz0), d1 = make_float2(dz1, dz1);
#pragma unroll
        for (int q = 0; q < DP; ++q) {
          const float4 t = xr[p][q];
          const float2 lo = make_float2(t.x, t.y), hi = make_float2(t.z, t.w);
          a0[2 * q] = __ffma2_rn(lo, d0, a0[2 * q]);
          a1[2 * q] = __ffma2_rn(lo, d1, a1[2 * q]);
          a0[2 * q + 1] = __ffma2_rn(hi, d0, a0[2 * q + 1]);
          a1[2 * q + 1] = __ffma2_rn(hi, d1, a1[2 * q + 1]);
        }
        av0 += r * h0[p];
        av1 += r * h1[p];
        ab += r;
      }
    }
    }  // (other activations: the unit-per-lane forward, one butterfly per row pair)
  };

  // ---- noise: jax.random.normal(key, (P,))[i] -> zs[tab[i]], lane l draws i = l + 32 m (4 threefry chains in flight)
  auto gen_noise = [&](PbKey key) {
    csync();  // the previous readers are done with zs
    for (int p0 = ctid; p0 < P; p0 += 4 * CT) {
      uint32_t idx[4];
      uint32_t to[4];
      float z[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        idx[j] = (uint32_t)(p0 + CT * j);
        to[j] = tab[p0 + CT * j];  // (the table is padded to a multiple of 4 CT: i >= P lands in the dump row)
      }
      pb_normal_n<4, true>(key, idx, z);
#pragma unroll
      for (int j = 0; j < 4; ++j) zs[to[j]] = z[j];
    }
    csync();
  };

  auto trace_row = [&](int t_) -> float* {
    return (a.trace && t_ >= a.burnin && ((t_ - a.burnin) % a.thin) == 0)
               ? a.trace + ((size_t)c * a.T_out + (size_t)((t_ - a.burnin) / a.thin)) * P
               : nullptr;
  };
  auto store_theta = [&](float* dst) {  // registers -> slot-major staging -> flat order, coalesced
    csync();
    if (wc == 0) {
#pragma unroll
      for (int u = 0; u < NS; ++u) zs[u * 32 + lane] = TH_(u);  // (pad slots hold 0; the b2 replicas agree)
    }
    csync();
    for (int i = ctid; i < P; i += CT) dst[i] = zs[tab[i]];
  };

  // ---- chain set-up (mirrors pb_sg_chain) -------------------------------------------------------------------------
  if (ALGO == PB_ALGO_PSGLD || ALGO == PB_ALGO_ADAM) {
    const bool have = ALGO == PB_ALGO_ADAM || a.has_state;
#pragma unroll
    for (int u = 0; u < NS; ++u) {
      const bool ld = have && slot_ok(u);
      ax1[u * 32] = ld ? a.aux1[cP + slot_fi(u)] : 0.f;
      ax2[u * 32] = ld ? a.aux2[cP + slot_fi(u)] : 0.f;
    }
  }
  if (ALGO == PB_ALGO_SGHMC) {
#pragma unroll
    for (int u = 0; u < NS; ++u) ax1[u * 32] = 0.f;
  }

  const bool explicit_idx = a.idx != nullptr || a.contig;
  const bool per_step =
      explicit_idx ? (a.idx_steps > 1 || ALGO == PB_ALGO_ADAM) : (a.mb_mode == PBNN_MB_PER_STEP);
  const int32_t* cidx = a.idx ? a.idx + (size_t)c * a.idx_steps * B : nullptr;
  int gen_pos = 0;

  // pSGLD init (psgld.py:25-29): g0 = grad(theta0, draw #init_draw), V0 = g0^2
  if (ALGO == PB_ALGO_PSGLD && !a.has_state) {
    if (explicit_idx) {
      copy_indices(cidx);
    } else {
      gen_advance(a.init_draw - gen_pos);
      gen_pos = a.init_draw;
      gen_indices();
    }
    gather();
    grad();
#pragma unroll
    for (int u = 0; u < NS; ++u) {
      const float g = GA_(u) - TH_(u) * pl.inv_pv;
      ax1[u * 32] = g;
      ax2[u * 32] = g * g;
    }
  }

  if (!per_step) {  // the minibatch the traced loop re-uses (quirk Q1)
    if (explicit_idx) {
      copy_indices(cidx);
    } else {
      gen_advance(a.which_draw - gen_pos);
      gen_pos = a.which_draw;
      gen_indices();
    }
    gather();
  } else if (!explicit_idx) {
    gen_advance((int)(1 + a.t0 - gen_pos));  // draw #1 belongs to kern.init, step t uses draw t0 + t + 2
  }

  for (int t = 0; t < a.T; ++t) {
    const long long tg = a.t0 + t;
    const PbKey kt = a.direct_key ? ck : pb_split_at(ck, (uint32_t)tg);  // split(rng_key, T)[t]
    const float eps = a.step_sizes ? a.step_sizes[t < a.n_step_sizes ? t : a.n_step_sizes - 1] : a.const_step;
    if (per_step) {
      if (explicit_idx) {
        copy_indices(cidx ? cidx + (size_t)t * B : nullptr);
      } else {
        gen_advance(1);
        gen_indices();
      }
      gather();
    }
    const float* cvc = a.cv_center ? a.cv_center + cP : nullptr;
    const float* cve = a.cv_center ? a.cv_center_est + cP : nullptr;

    if (ALGO == PB_ALGO_SGLD) {
      grad();
      gen_noise(kt);
      const float ns = sqrtf(2.0f * a.temperature * eps);
#pragma unroll
      for (int u = 0; u < NS; ++u) {
        const float th = TH_(u);
        float g = GA_(u) - th * pl.inv_pv;
        if (cvc) g = (cvc[slot_fi(u)] + g) - cve[slot_fi(u)];  // blackjax control_variates
        TH_(u) = th + eps * g + ns * ZS_(u);  // (pad slots: theta = g = z = 0)
      }
    } else if (ALGO == PB_ALGO_PSGLD) {  // diffusions.py:42-59: move with the STALE (g, V), then refresh them
      gen_noise(kt);
#pragma unroll
      for (int u = 0; u < NS; ++u) {
        // (sqrt.approx / rcp.approx, 2^-22 relative each: the IEEE sqrtf and division cost ~100 instructions per slot --
        //  more than the slot's share of the noise generation)
        const float pre = pb_rcp_fast(pb_sqrt_fast(ax2[u * 32]) + 1e-6f);
        TH_(u) = TH_(u) + eps * pre * ax1[u * 32] + pb_sqrt_fast(2.0f * pre * eps) * ZS_(u);  // (pad slots: 0)
      }
      grad();
#pragma unroll
      for (int u = 0; u < NS; ++u) {
        const float g = GA_(u) - TH_(u) * pl.inv_pv;
        ax1[u * 32] = g;
        ax2[u * 32] = a.alpha * ax2[u * 32] + (1.0f - a.alpha) * (g * g);
      }
    } else if (ALGO == PB_ALGO_SGHMC) {
      const float seps = sqrtf(eps);
      const float mom_mu = a.mom_mu0 + a.mom_mu1 * seps, mom_sig = a.mom_sig0 + a.mom_sig1 * seps;
      gen_noise(kt);
#pragma unroll
      for (int u = 0; u < NS; ++u) {
        ax1[u * 32] = slot_ok(u) ? mom_mu + mom_sig * ZS_(u) : 0.f;
      }
      const float ns = sqrtf(2.0f * eps * (a.alpha - a.beta) * a.temperature);
      const float fric = 1.0f - a.alpha;
      for (int l = 0; l < a.L; ++l) {
        const PbKey kl = pb_split_at(kt, (uint32_t)l);  // split(keys[t], L)[l]
        grad();
        gen_noise(kl);
#pragma unroll
        for (int u = 0; u < NS; ++u) {
          const float th = TH_(u), p = ax1[u * 32];
          float g = GA_(u) - th * pl.inv_pv;
          if (cvc) g = (cvc[slot_fi(u)] + g) - cve[slot_fi(u)];
          TH_(u) = th + p;  // (pad slots: everything 0)
          ax1[u * 32] = fric * p + eps * g + ns * ZS_(u);
        }
      }
    } else if (ALGO == PB_ALGO_ADAM) {
      grad();
      const float cnt = (float)(a.count0 + t + 1);
      const float ibc1 = 1.0f / (1.0f - powf(a.b1, cnt)), ibc2 = 1.0f / (1.0f - powf(a.b2, cnt));
#pragma unroll
      for (int u = 0; u < NS; ++u) {
        const float th = TH_(u);
        const float g = -(GA_(u) - th * pl.inv_pv);  // gradient of the loss -logposterior
        const float m = (1.0f - a.b1) * g + a.b1 * ax1[u * 32];
        const float v = (1.0f - a.b2) * (g * g) + a.b2 * ax2[u * 32];
        ax1[u * 32] = m;
        ax2[u * 32] = v;
        const float tn = th + (-a.lr) * ((m * ibc1) * pb_rcp_fast(pb_sqrt_fast(v * ibc2) + a.eps));
        TH_(u) = slot_ok(u) ? tn : 0.f;
      }
    }
    float* dst = trace_row(t);
    if (dst) store_theta(dst);
  }

  if (!a.no_theta_store) {
    store_theta(a.theta + cP);
#pragma unroll
    for (int u = 0; u < NS; ++u) fl |= pb_finite(TH_(u)) ? 0u : 1u;
  }
  if (ALGO == PB_ALGO_PSGLD || ALGO == PB_ALGO_ADAM) {
#pragma unroll
    for (int u = 0; u < NS; ++u) {
      if (slot_ok(u)) {
        a.aux1[cP + slot_fi(u)] = ax1[u * 32];
        a.aux2[cP + slot_fi(u)] = ax2[u * 32];
      }
    }
  }
  fl = __reduce_or_sync(0xffffffffu, fl);
  if constexpr (NWC == 1) {
    if (lane == 0 && a.flags) a.flags[c] = fl;
  } else {
    if (lane == 0 && a.flags && fl) atomicOr(a.flags + c, fl);  // (zeroed by the kernel before the chain starts)
  }
#undef TH_
#undef GA_
#undef ZS_
}

// few chains (SGLD, relu): one CTA of four warps per chain
template <int DP>
__global__ void __launch_bounds__(128, 4)
pb_sgc_kernel(const __grid_constant__ PbPlan pl, const __grid_constant__ PbSgArgs a) {
  extern __shared__ __align__(16) float pb_smem[];
  const int c = (int)blockIdx.x;
  if (threadIdx.x == 0 && a.flags) a.flags[c] = 0u;
  __syncthreads();
  pb_sgw_chain<PB_ALGO_SGLD, PBNN_ACT_RELU, DP, 4>(pl, a, pb_smem, c, (int)threadIdx.x & 31);
}

template <int ALGO, int ACT, int DP>
__global__ void __launch_bounds__(32 * PB_W1_WPC, 16 / PB_W1_WPC)
pb_sgw_kernel(const __grid_constant__ PbPlan pl, const __grid_constant__ PbSgArgs a) {
  extern __shared__ __align__(16) float pb_smem[];
  const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
  const int c = (int)blockIdx.x * PB_W1_WPC + warp;
  if (c >= a.n_chains) return;  // (no CTA barrier anywhere in this kernel)
  pb_sgw_chain<ALGO, ACT, DP>(pl, a, pb_smem + (size_t)warp * pl.w1_warp_floats, c, lane);
}
#endif
