// Launch plan + argument blocks shared by the CUDA kernels and the host-side C ABI.
// Plain C++ (no CUDA types) so the same header serves nvcc and the g++-built kernel-logic emulator
// under tests/emu (test infrastructure only).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pbnn_b200.h"

#define PB_SMEM_LIMIT_BYTES (227 * 1024)
#ifndef PB_MAX_NT
#define PB_MAX_NT 512
#endif
#define PB_SCR_MAIN (3 * PB_MAX_NT)                       // per-thread partials of up to 3 block reductions
#define PB_SCR_FLOATS (PB_SCR_MAIN + 64 + 128 + 128)    // + second level [64], llbuf[128], fbuf[128]
#define PB_GPART_BUDGET_BYTES (24 * 1024)      // split-K partial gradient buffers

// One Dense layer in the resident (shared-memory) layout.
//   W_aug : rows x wp floats, rows = roundup4(in+1); row r<in is kernel row r, row `in` is the bias,
//           remaining rows and the columns >= out are zero pads that are never updated.
//   wp    : pitch, multiple of 4 with wp % 8 == 4 so that 8 consecutive rows hit 8 distinct 16-byte
//           bank groups (conflict-free LDS.128 in the k-contiguous GEMMs).
//   dw_ni : lane layout of the weight-gradient GEMM: 8 -> warp tile 32(in) x 16(out), 4 -> 16(in) x 32(out)
//   dw_ks : split-K factor of that GEMM (slices > 0 accumulate into partial buffers)
struct PbLayer {
  int in, out;
  int inA;   // in + 1
  int rows;  // roundup4(in + 1)
  int wp;
  int soff;  // float offset of W_aug inside a resident parameter array
  int fb, fw;  // flat (ravel_pytree) offsets of bias and kernel
  int dw_ni, dw_ks;
  unsigned long long div_out;  // ceil(2^40 / out): q / out == (q * div_out) >> 40 exactly for q < 2^28
};

struct PbPlan {
  int nl, act, P, Ps;
  int d, s;
  int lik, sy;      // likelihood kind; number of target columns (1 for heteroscedastic, else s)
  float lik_scale;  // 1/sigma^2 (homoscedastic) or 1
  PbLayer L[PBNN_MAX_LAYERS];
  int BT, ap;                     // batch-tile rows (multiple of 32) and activation pitch (BT + 4)
  int aoff[PBNN_MAX_LAYERS + 2];  // tile buffers A_0 .. A_nl, then yT (A_0 / yT have 0 rows when res_rows > 0)
  int res_rows;                   // > 0: all input rows of a gradient evaluation are resident, feature-major
  int xp, off_xres, off_yres;     //      X^T_aug [r4(d+1)][xp], y^T [r4(s)][xp], xp = roundup(rows, BT) + 4
  int nstate;                     // resident parameter-sized arrays
  int ksmax, off_gpart;           // (ksmax - 1) partial gradient arrays for split-K
  int gstate;                     // resident arrays live in a global-memory workspace (model too large for smem)
  int h1, h1_dp;                  // fused one-hidden-layer path: X row-major [rows][4*h1_dp], per-warp partial grads
  int h1_nw;                      // warps per CTA the fused path should use (0: the caller's default)
  int off_xr, off_yr, off_hpart;
  // warp-per-chain engine (pbnn_warp.cuh; SG samplers / Adam on the same nets, many chains): float offsets inside a
  // warp's private shared-memory block of w1_warp_floats floats -- rows [B][4 h1_dp], targets, indices, noise, aux state
  int w1, w1_warp_floats, w1_ox, w1_oy, w1_oi, w1_ot, w1_oz, w1_ow, w1_oa;
  // Adam for small two-hidden-layer relu nets, four warps per member (pbnn_warp2.cuh): float offsets inside the CTA's block;
  // w2_jr = roundup4(H + 1) rows of the layer-2 weights, w2_pa = activation pitch (= 4 mod 8: conflict-free LDS.128)
  int w2, w2_warp_floats, w2_dp, w2_jr, w2_pa, w2_ox, w2_ow1, w2_ow2, w2_ow3, w2_oh, w2_os, w2_oz, w2_od;
  // tcgen05 path (one hidden relu layer, scalar output): see pbnn_tc.cuh for the operand layouts
  int tc, tc_npad;                // tc_npad = rows rounded up to the 128-row MMA tile
  int off_mk, off_z, off_wb, off_w2s, off_tcm;
  // tcgen05 engine for two equal hidden relu layers and a scalar output (pbnn_tc2.cuh): bf16x3 operands, weights
  // streamed from a per-CTA global image block with cp.async.bulk, state arrays in the global workspace
  int tc2, t2_H, t2_Hp, t2_Hy, t2_Kx;      // hidden width, padded to 16, (H + 1) padded to 16, (d + 1) padded to 16
  int t2_nch, t2_nchy, t2_ncbx, t2_nmt;    // 64-wide chunks of Hp / Hy / Kx; 128-row tiles of Hp
  int t2_wstage, t2_ring, t2_nz, t2_help;  // bytes of one weight-ring stage / of the whole ring; noise-ring slots in use;
                                           // noise warps drain every other D2 step (all noise of a step fits the ring)
  int off_t2_vec, off_t2_bar, off_t2_big;  // float offsets: small vectors, mbarriers, 1024-byte aligned operand area
  int t2_img_w1, t2_img_w2, t2_img_z, t2_img_bytes;  // byte offsets inside the image block (X image at 0; z: noise ring)
  int off_act, off_scr, off_acc, off_idx, off_scal;
  int idx_cap;
  int smem_floats;
  int NT;
  float inv_s2, inv_pv, logprior_const;
};

static inline int pb_r4(int x) { return (x + 3) & ~3; }
static inline int pb_r32(int x) { return (x + 31) & ~31; }

// Engine-selection / tuning options.  The ONLY process-wide state of the library: an explicit table written through
// pbnn_set_option() (tests and experiments select engines with it); nothing is read from the environment.
enum {
  PB_OPT_DW_KS = 0, PB_OPT_NW, PB_OPT_NO_RESIDENT, PB_OPT_NO_H1, PB_OPT_NO_TC, PB_OPT_BT, PB_OPT_TWO_CTA,
  PB_OPT_ONE_CTA, PB_OPT_FORCE_GSTATE, PB_OPT_NO_SCHED, PB_OPT_SEG_LEN, PB_OPT_DEBUG, PB_OPT_PRED_TPC,
  PB_OPT_PRED_CHUNKS, PB_OPT_TC2_MIN_H, PB_OPT_TC2_MIN_ROWS, PB_OPT_NO_TC2, PB_OPT_TC2_CTAS, PB_OPT_TC2_NZ, PB_OPT_TC2_NOHELP, PB_OPT_NO_W1,
  PB_OPT_W1_MIN_CHAINS, PB_OPT_NO_W2, PB_OPT_NO_W1C, PB_OPT_COUNT
};
static const char* const pb_opt_names[PB_OPT_COUNT] = {
    "dw_ks", "nw", "no_resident", "no_h1", "no_tc", "bt", "two_cta", "one_cta", "force_gstate", "no_sched", "seg_len",
    "debug", "pred_tpc", "pred_chunks", "tc2_min_h", "tc2_min_rows", "no_tc2", "tc2_ctas", "tc2_nz", "tc2_nohelp",
    "no_w1", "w1_min_chains", "no_w2", "no_w1c"};
#define PB_OPT_UNSET (-2147483647 - 1)
inline int* pb_opt_table() {
  static int t[PB_OPT_COUNT];
  static bool init = false;
  if (!init) {
    for (int i = 0; i < PB_OPT_COUNT; ++i) t[i] = PB_OPT_UNSET;
    init = true;
  }
  return t;
}
static inline int pb_opt(int id, int dflt) {
  const int v = pb_opt_table()[id];
  return v == PB_OPT_UNSET ? dflt : v;
}
static inline int pb_opt_find(const char* name) {
  for (int i = 0; i < PB_OPT_COUNT; ++i)
    if (name && !strcmp(name, pb_opt_names[i])) return i;
  return -1;
}

static inline int pb_num_params(const pbnn_mlp_spec* sp) {
  if (!sp || sp->n_layers < 1 || sp->n_layers > PBNN_MAX_LAYERS) return -1;
  long P = 0;
  for (int l = 0; l <= sp->n_layers; ++l)
    if (sp->dims[l] < 1 || sp->dims[l] > 4096) return -1;
  for (int l = 0; l < sp->n_layers; ++l) P += (long)sp->dims[l + 1] * (sp->dims[l] + 1);
  return P > (1L << 26) ? -1 : (int)P;
}

// choose the lane layout / split-K of every weight-gradient GEMM for `nw` warps and batch tile `bt`
static inline void pb_plan_dw(PbPlan* pl, int nw, int bt, int kcap) {
  pl->ksmax = 1;
  const int ks_env = pb_opt(PB_OPT_DW_KS, 0);
  for (int l = 0; l < pl->nl; ++l) {
    PbLayer& Ly = pl->L[l];
    int best_ni = 8, best_ks = 1;
    long best = -1;
    for (int ni = 8; ni >= 4; ni -= 4) {
      const int ti = ni * 4, to = (32 / ni) * 4;
      const int tiles = ((Ly.inA + ti - 1) / ti) * ((Ly.out + to - 1) / to);
      for (int ks = 1; ks <= 8; ks *= 2) {
        if (bt % (4 * ks)) continue;
        if (ks > 1 && (size_t)(ks - 1) * pl->Ps * 4 > PB_GPART_BUDGET_BYTES) continue;
        if (ks > kcap || (ks_env > 0 && ks > ks_env)) continue;
        const long rounds = (tiles * ks + nw - 1) / nw;
        const long cost = rounds * (bt / ks + 6) + (ks > 1 ? 4 : 0);
        if (best < 0 || cost < best) {
          best = cost;
          best_ni = ni;
          best_ks = ks;
        }
      }
    }
    Ly.dw_ni = best_ni;
    Ly.dw_ks = best_ks;
    if (best_ks > pl->ksmax) pl->ksmax = best_ks;
  }
}

#define PB_PLAN_RES 1
#define PB_PLAN_ACC 2
#define PB_PLAN_TC 4  // the caller has a tcgen05 kernel for this op (HMC): allow the tensor-core layout
#define PB_PLAN_TC2 8  // the caller can run the two-hidden-layer tcgen05 engine (SG samplers, Adam, grad hook)
#define PB_PLAN_W1 16  // the caller can run the warp-per-chain engine (SG samplers, Adam)
#define PB_PLAN_W1C 128  // the caller is SGLD: with few chains a CTA of four warps per chain may run (pbnn_warp.cuh, NWC = 4)
#define PB_PLAN_W2 64  // the caller is the Adam trainer: small two-hidden-layer relu nets may run one warp per member
#define PB_PLAN_PRED 32  // posterior predictive: the MMA tile rows are test points (always full), narrower nets still pay
#define PB_W1_WARPS 1  // warps (= chains) per CTA of the warp-per-chain engine (finest balance over the SMs)

#define PB_T2_NT 576       // 8 worker warps + 1 TMA warp + 1 MMA-issue warp + 8 noise warps
#define PB_T2_NZ 16        // most slots of the noise ring (one slot = 8 normals for each of the 256 worker threads)
#define PB_T2_ZSLOT 2048   // floats per slot
#define PB_T2_NS 3         // weight-ring stages
#define PB_T2_TILE 16384   // bytes of one [128 rows][64 bf16] operand tile (SWIZZLE_128B)
#define PB_T2_VEC_FLOATS 1856
static inline int pb_r16(int x) { return (x + 15) & ~15; }

// two-hidden-layer tcgen05 plan: shared memory = scratch | small vectors | mbarriers | aligned operand area
// (activation chunk x 3 terms | relu mask of layer 2 | weight ring); everything parameter-sized lives in global memory
static inline void pb_layout_tc2(PbPlan* pl, const pbnn_mlp_spec* sp, int rows) {
  const int H = sp->dims[1];
  pl->tc2 = 1;
  pl->t2_H = H;
  pl->t2_Hp = pb_r16(H);
  pl->t2_Hy = pb_r16(H + 1);
  pl->t2_Kx = pb_r16(pl->d + 1);
  pl->t2_nch = (pl->t2_Hp + 63) / 64;
  pl->t2_nchy = (pl->t2_Hy + 63) / 64;
  pl->t2_ncbx = (pl->t2_Kx + 63) / 64;
  pl->t2_nmt = (pl->t2_Hp + 127) / 128;
  const int st_a = pl->t2_nch * 64 * 128, st_b = pl->t2_Hp * 128;
  pl->t2_wstage = st_a > st_b ? st_a : st_b;
  const int xi = 3 * pl->t2_ncbx * PB_T2_TILE;
  pl->t2_ring = PB_T2_NS * pl->t2_wstage > xi ? PB_T2_NS * pl->t2_wstage : xi;
  const int d1 = (((pl->d + 1) * (pl->t2_Hp + 4) * 4) + 1023) & ~1023;  // first-layer gradient dump [d + 1][Hp + 4]
  if (d1 > pl->t2_ring) pl->t2_ring = d1;
  pl->BT = 128;
  pl->ap = 132;
  pl->h1 = 0;
  pl->tc = 0;
  pl->ksmax = 1;
  pl->gstate = 1;
  pl->NT = PB_T2_NT;
  pl->res_rows = rows;
  pl->xp = 0;
  int off = 0;
  pl->off_gpart = pl->off_act = off;
  for (int l = 0; l <= pl->nl + 1; ++l) pl->aoff[l] = off;
  pl->off_xres = pl->off_yres = pl->off_xr = pl->off_yr = pl->off_hpart = off;
  pl->off_scr = off;
  off += PB_SCR_FLOATS;
  pl->off_acc = off;
  pl->off_idx = off;
  off += pl->idx_cap;
  pl->off_scal = off;
  off += 32;
  pl->off_t2_vec = off;
  off += PB_T2_VEC_FLOATS;
  pl->off_t2_bar = off;
  off += 160;  // up to 64 mbarriers + the TMEM base slot
  pl->off_t2_big = off;
  off += 256;  // slack for the run-time 1024-byte alignment
  off += (3 * PB_T2_TILE + pl->t2_nch * PB_T2_TILE + pl->t2_ring) / 4;
  pl->smem_floats = off;
  pl->t2_img_w1 = xi;
  pl->t2_img_w2 = pl->t2_img_w1 + 3 * pl->t2_nch * pl->t2_Kx * 128;
  pl->t2_img_z = pl->t2_img_w2 + 3 * pl->t2_nch * pl->t2_Hp * 128;
  // noise ring: deep (16 slots = 128 normals per worker ahead) for small nets, where the whole step's noise is drawn while
  // the forward / backward GEMMs run; shallow for wide nets, whose per-CTA working set (operand images + ring) must stay
  // inside the L2 for 148 concurrent chains (measured, 90-256-256-1: 162 us per chain-step at 113 MB, 139 us at 83 MB)
  pl->t2_nz = pb_opt(PB_OPT_TC2_NZ, pl->t2_Hp > 128 ? 4 : PB_T2_NZ);
  if (pl->t2_nz < 2) pl->t2_nz = 2;
  if (pl->t2_nz > PB_T2_NZ) pl->t2_nz = PB_T2_NZ;
  {  // 8-weight blocks of one evaluation, as the drains and the noise warps count them
    const int items = (pl->d + 1) * ((H + 7) / 8);
    int blocks = (items + 255) / 256;
    for (int ic = 0; ic < pl->t2_nchy; ++ic) {
      const int ncols = pl->t2_Hy - 64 * ic < 64 ? pl->t2_Hy - 64 * ic : 64;
      blocks += pl->t2_nmt > 1 ? ncols / 8 : (ncols / 8 < 4 ? ncols / 8 : 4);
    }
    pl->t2_help = blocks <= pl->t2_nz && !pb_opt(PB_OPT_TC2_NOHELP, 0);
  }
  pl->t2_img_bytes = pl->t2_img_z + pl->t2_nz * PB_T2_ZSLOT * 4;
}

// tcgen05 operand geometry (floats).  K-major, no swizzle: element (row x, k) at (k/4)*LBO + x*16 B + (k%4)*4 B.
#define PB_TC_MK_LBO 516   // mask^T  [k = row in half / 4][2 halves x 64 hidden units (+1 pad row: conflict-free stores)][4]
#define PB_TC_Z_LBO 260    // Z       [k = row in half / 4][2 halves x (raw, lo) x 16 inputs (+1 pad row)][4]
#define PB_TC_MK_FLOATS (16 * PB_TC_MK_LBO)  // one 128-row tile = 2 halves x 64 rows = 16 k-groups
#define PB_TC_Z_FLOATS (16 * PB_TC_Z_LBO)
#define PB_TC_MAX_TILES 32  // y stays in shared memory (N <= 4096); X (hi, lo) lives in TMEM, 4 tiles at a time

// shared-memory layout of the tcgen05 plan: state | y | mask | Z | W1 (raw, lo) | w2 | ... (X lives in TMEM)
static inline void pb_layout_tc(PbPlan* pl, int rows, int nstate) {
  pl->BT = 128;
  pl->ap = 132;
  pl->h1 = 1;
  pl->tc = 1;
  pl->ksmax = 1;
  pl->gstate = 0;
  pl->NT = 160;  // 4 epilogue warps + 1 MMA-issue warp; two CTAs per SM
  pl->h1_dp = pb_r4(pl->d + 1) / 4;
  pl->tc_npad = ((rows + 127) / 128) * 128;
  int off = nstate * pl->Ps;
  pl->off_gpart = off;
  pl->off_act = off;
  for (int l = 0; l <= pl->nl + 1; ++l) pl->aoff[l] = off;
  pl->res_rows = rows;
  pl->xp = 0;
  pl->off_xres = pl->off_yres = pl->off_xr = pl->off_hpart = off;
  pl->off_yr = off;
  off += pl->tc_npad;
  pl->off_mk = off;
  off += PB_TC_MK_FLOATS;
  pl->off_z = off;
  off += PB_TC_Z_FLOATS;
  pl->off_wb = off;
  off += 2 * 16 * 64;
  pl->off_w2s = off;
  off += 128;  // output weights, double-buffered
  pl->off_tcm = off;         // 5 mbarriers (8 B each) [0..9], TMEM base [12], per-warp residual sums [16..23]
  off += 32;
  pl->off_scr = off;
  off += PB_SCR_FLOATS;
  pl->off_acc = off;
  pl->off_idx = off;
  off += pl->idx_cap;
  pl->off_scal = off;
  off += 32;
  pl->smem_floats = off;
}

// shared-memory layout for one candidate configuration
static inline void pb_layout(PbPlan* pl, const pbnn_mlp_spec* sp, int rows, int nstate, int bt, int res, int kcap,
                             int want_acc, int h1, int gstate = 0) {
  const int ap = bt + 4;
  pl->BT = bt;
  pl->ap = ap;
  pl->h1 = h1;
  pb_plan_dw(pl, pl->NT / 32, bt, kcap);
  if (h1) pl->ksmax = 1;
  pl->gstate = gstate;
  int off = gstate ? 0 : nstate * pl->Ps;
  pl->off_gpart = off;
  off += (pl->ksmax - 1) * pl->Ps;
  pl->off_act = off;
  if (h1) {
    // fused path: no activation buffers; X row-major with the ones column, y, one partial gradient per warp
    if (pl->h1_nw > 0) pl->NT = 32 * pl->h1_nw;
    else if (pl->NT > 256) pl->NT = 256;
    for (int l = 0; l <= pl->nl + 1; ++l) pl->aoff[l] = off;
    pl->res_rows = rows;
    pl->h1_dp = pb_r4(pl->d + 1) / 4;
    pl->xp = 0;
    pl->off_xres = pl->off_yres = off;
    pl->off_xr = off;
    off += rows * 4 * pl->h1_dp;
    pl->off_yr = off;
    off += pb_r4(rows);
    pl->off_hpart = off;
    off += (pl->NT / 32) * pl->Ps;
    pl->off_scr = off;
    off += PB_SCR_FLOATS;
    pl->off_acc = off;
    pl->off_idx = off;
    off += pl->idx_cap;
    pl->off_scal = off;
    off += 32;
    pl->smem_floats = off;
    return;
  }
  for (int l = 0; l <= pl->nl + 1; ++l) {
    int r;
    if (l < pl->nl) r = pb_r4(sp->dims[l] + 1);
    else if (l == pl->nl) r = pl->s == 1 ? 1 : pb_r4(pl->s);  // A_nl: output / dZ_last (rank-1 path: one row)
    else r = pl->sy;                                           // yT
    if (res && (l == 0 || l == pl->nl + 1)) r = 0;
    pl->aoff[l] = off;
    off += r * ap;
  }
  pl->res_rows = 0;
  if (res) {
    pl->res_rows = rows;
    pl->xp = ((rows + bt - 1) / bt) * bt + 4;
    pl->off_xres = off;
    off += pb_r4(pl->d + 1) * pl->xp;
    pl->off_yres = off;
    off += pl->sy * pl->xp;
  }
  pl->off_scr = off;
  off += PB_SCR_FLOATS;
  pl->off_acc = off;
  if (want_acc) off += 2 * pb_r4(pl->s) * ap;  // predictive accumulators
  pl->off_idx = off;
  off += pl->idx_cap;
  pl->off_scal = off;
  off += 32;
  pl->smem_floats = off;
}

// rows: rows per gradient evaluation (minibatch B, N for HMC, M for predict); nstate: resident
// parameter-sized arrays; idx_cap: ints reserved for the minibatch index list; flags: PB_PLAN_RES = try to keep
// all `rows` input rows resident in shared memory (feature-major) instead of streaming tiles, PB_PLAN_ACC =
// reserve the predictive accumulators; chains_hint: number of CTAs of the launch (occupancy target).
static inline int pb_build_plan(const pbnn_mlp_spec* sp, int rows, int nstate, int idx_cap, int flags,
                                int chains_hint, PbPlan* pl, char* err, size_t errn) {
  memset(pl, 0, sizeof(*pl));
  const int P = pb_num_params(sp);
  if (P < 0) {
    snprintf(err, errn, "invalid mlp spec (1..%d layers, widths 1..4096)", PBNN_MAX_LAYERS);
    return PBNN_EINVAL;
  }
  if (sp->activation != PBNN_ACT_RELU && sp->activation != PBNN_ACT_TANH) {
    snprintf(err, errn, "unknown activation %d", sp->activation);
    return PBNN_EINVAL;
  }
  const int hetero = sp->likelihood == PBNN_LIK_HETEROSCEDASTIC;
  if (sp->likelihood != PBNN_LIK_HOMOSCEDASTIC && !hetero) {
    snprintf(err, errn, "unknown likelihood %d", sp->likelihood);
    return PBNN_EINVAL;
  }
  if (hetero && sp->dims[sp->n_layers] != 2) {
    snprintf(err, errn, "the heteroscedastic likelihood needs a network with 2 outputs (mu, log sigma^2)");
    return PBNN_EINVAL;
  }
  if ((!hetero && !(sp->noise_level > 0.f)) || !(sp->prior_scale > 0.f)) {
    snprintf(err, errn, "noise_level and prior_scale must be > 0");
    return PBNN_EINVAL;
  }
  if (rows < 1) {
    snprintf(err, errn, "rows must be >= 1");
    return PBNN_EINVAL;
  }
  pl->nl = sp->n_layers;
  pl->act = sp->activation;
  pl->P = P;
  pl->d = sp->dims[0];
  pl->s = sp->dims[sp->n_layers];
  pl->lik = sp->likelihood;
  pl->sy = hetero ? 1 : pl->s;
  pl->inv_s2 = hetero ? 1.0f : 1.0f / (sp->noise_level * sp->noise_level);
  pl->lik_scale = pl->inv_s2;
  pl->inv_pv = 1.0f / (sp->prior_scale * sp->prior_scale);
  pl->logprior_const = 0.918938533204672742f + logf(sp->prior_scale);
  int soff = 0, foff = 0;
  for (int l = 0; l < pl->nl; ++l) {
    PbLayer& Ly = pl->L[l];
    Ly.in = sp->dims[l];
    Ly.out = sp->dims[l + 1];
    Ly.inA = Ly.in + 1;
    Ly.rows = pb_r4(Ly.in + 1);
    int wp = pb_r4(Ly.out);
    if ((wp & 7) != 4) wp += 4;
    Ly.wp = wp;
    Ly.soff = soff;
    soff += Ly.rows * wp;
    Ly.div_out = ((1ull << 40) + (unsigned long long)Ly.out - 1) / (unsigned long long)Ly.out;
    Ly.fb = foff;  // ravel_pytree: "bias" < "kernel"
    Ly.fw = foff + Ly.out;
    foff += Ly.out * (Ly.in + 1);
  }
  pl->Ps = soff;
  pl->nstate = nstate;
  pl->idx_cap = pb_r4(idx_cap);
  pl->NT = 32 * pb_opt(PB_OPT_NW, 8);
  if (pl->NT < 32 || pl->NT > PB_MAX_NT || (pl->NT & (pl->NT - 1))) {
    snprintf(err, errn, "PBNN_NW out of range");
    return PBNN_EINVAL;
  }
  const int want_res = (flags & PB_PLAN_RES) && !pb_opt(PB_OPT_NO_RESIDENT, 0);
  int maxbt = pb_r32(rows);
  if (maxbt > 128) maxbt = 128;
  // fused one-hidden-layer path (scalar output, H <= 64, d+1 <= 16, all rows resident); CUDA only (warp shuffles)
  int h1_ok = want_res && pl->nl == 2 && pl->s == 1 && sp->dims[1] <= 64 && pl->d + 1 <= 16 &&
              !(flags & PB_PLAN_ACC) && !pb_opt(PB_OPT_NO_H1, 0);
#if defined(PB_EMU)
  h1_ok = 0;
#endif
  const int nw_env = pb_opt(PB_OPT_NW, 0);
#if !defined(PB_EMU)
  // four-warps-per-member Adam (pbnn_warp2.cuh): d-H-H-1 relu, H <= 64, d + 1 <= 16, B <= 32 (BASELINE configs[3])
  if ((flags & PB_PLAN_W2) && pl->nl == 3 && pl->s == 1 && !hetero && sp->activation == PBNN_ACT_RELU &&
      sp->dims[1] == sp->dims[2] && sp->dims[1] <= 64 && pl->d + 1 <= 16 && rows <= 32 && !pb_opt(PB_OPT_NO_W2, 0)) {
    const int H = sp->dims[1], kd = pb_r4(pl->d + 1);
    pl->w2 = 1;
    pl->w2_dp = kd / 4;
    pl->w2_jr = pb_r4(H + 1);
    pl->w2_pa = (pl->w2_jr % 8) == 4 ? pl->w2_jr : pl->w2_jr + 4;
    int off = 0;
    pl->w2_ox = off;
    off += 2 * 32 * kd + 64;  // rows of this step and of the next one (cp.async), then y of both
    pl->w2_ow1 = off;
    off += kd * 64;
    pl->w2_ow2 = off;
    off += pl->w2_jr * 64;
    pl->w2_ow3 = off;
    off += 64 + 4 * 32 + 36;  // w3, per-warp partial outputs, residuals (+ the broadcast slot of b3)
    pl->w2_oh = off;
    off += 32 * pl->w2_pa;
    pl->w2_os = off;
    off += 32 * pl->w2_pa;
    pl->w2_oz = off;
    off += 32 * pl->w2_pa;
    pl->w2_od = off;
    off += 32 * pl->w2_pa + 64;  // (+ slack: lanes of the last row read a pair beyond the padded width)
    pl->w2_warp_floats = off;
    pl->NT = 128;
    pl->BT = 32;
    pl->smem_floats = off;
    return PBNN_OK;
  }
  // two-hidden-layer tcgen05 engine: d-H-H-1 relu, one 128-row tile per gradient evaluation
  if ((flags & PB_PLAN_TC2) && pl->nl == 3 && pl->s == 1 && !hetero && sp->activation == PBNN_ACT_RELU &&
      sp->dims[1] == sp->dims[2] && sp->dims[1] <= 256 &&
      sp->dims[1] >= pb_opt(PB_OPT_TC2_MIN_H, (flags & PB_PLAN_PRED) ? 32 : 64) &&
      pl->d + 1 <= 128 && rows <= 128 && rows >= pb_opt(PB_OPT_TC2_MIN_ROWS, 64) && !(flags & PB_PLAN_ACC) &&
      !pb_opt(PB_OPT_NO_TC, 0) && !pb_opt(PB_OPT_NO_TC2, 0)) {
    pb_layout_tc2(pl, sp, rows);
    if ((size_t)pl->smem_floats * 4 <= (size_t)PB_SMEM_LIMIT_BYTES) return PBNN_OK;
    pl->tc2 = 0;
    pl->gstate = 0;
  }
#endif
  // tcgen05 path: the same nets, relu only (the weight gradient uses the 0/1 activation mask as an exact operand)
  if (h1_ok && (flags & PB_PLAN_TC) && sp->activation == PBNN_ACT_RELU && !hetero && !pb_opt(PB_OPT_NO_TC, 0) &&
      rows <= 128 * PB_TC_MAX_TILES) {
    pb_layout_tc(pl, rows, nstate);
    if ((size_t)pl->smem_floats * 4 <= (size_t)PB_SMEM_LIMIT_BYTES) return PBNN_OK;
    pl->tc = 0;
  }
#if !defined(PB_EMU)
  // Warp-per-chain engine: one-hidden-layer nets, B <= 256 rows resident per warp, more chains than the multi-warp
  // fused path can keep co-resident (two CTAs per SM).  Measured, 8-50-1 B = 32 (M chain-steps/s, fused | warp):
  // 148 chains 39 | 30, 296 chains 62 | 59, 444 chains 55 | 88, 2048 chains 91-101 | 186: a lone warp advances a chain in
  // ~5 us per step against ~3.8 us for eight co-operating warps, but sixteen of them share an SM.
  // (fewer chains, SGLD on a relu net: the same engine with a CTA of FOUR warps per chain -- w1 = 2.  Measured against
  //  one warp per chain, 8-50-1 B = 32, M chain-steps/s: 296 chains 76 | 60, 444 chains 102 | 89, 592 chains 108 | 118)
  const int w1c_ok = (flags & PB_PLAN_W1C) && sp->activation == PBNN_ACT_RELU && !pb_opt(PB_OPT_NO_W1C, 0);
  const int w1_many = chains_hint >= pb_opt(PB_OPT_W1_MIN_CHAINS, w1c_ok ? 520 : 2 * 148 + 1);
  const int w1_few = !w1_many && w1c_ok;
  if (h1_ok && (flags & PB_PLAN_W1) && !pb_opt(PB_OPT_NO_W1, 0) && rows <= 256 && (w1_many || w1_few)) {
    pl->NT = 256;
    pb_layout(pl, sp, rows, nstate, maxbt, 1, 1, 0, 1);
    const int kd = 4 * pl->h1_dp, ns = 2 * kd + 3, nwc = w1_few ? 4 : 1;
    int off = 0;
    pl->w1_ox = off;
    off += rows * kd;
    pl->w1_oy = off;
    off += pb_r4(rows);
    pl->w1_oi = off;
    off += pb_r4(rows);
    pl->w1_ot = off;
    off += (((P + 128 * nwc - 1) / (128 * nwc)) * (128 * nwc)) / 2;  // flat index -> [slot][lane] table (uint16)
    pl->w1_oz = off;
    off += (ns + 1) * 32;     // noise / trace staging in slot order (+ a dump row)
    pl->w1_ow = off;
    off += kd * 64 + 64 + 64 + 32;  // relu: published first-layer rows, output weights, ballots, residuals
    if (nwc > 1) off += nwc * 32 + nwc * (kd / 2) * 128;  // partial outputs and partial gradients of the warps
    pl->w1_oa = off;
    off += nstate > 2 ? 2 * ns * 32 : 0;
    pl->w1_warp_floats = off;
    pl->w1 = w1_few ? 2 : 1;
    pl->NT = w1_few ? 128 : 32 * PB_W1_WARPS;
    pl->smem_floats = w1_few ? off : PB_W1_WARPS * off;
    if ((size_t)off * 4 <= (size_t)((w1_few ? 56 : 24) * 1024)) return PBNN_OK;  // >= 9 warps / >= 4 four-warp CTAs per SM
    pl->w1 = 0;
  }
#endif
  // Fused path, many chains: the warp-autonomous gradient wants >= 16 rows per warp (fewer partial-gradient arrays to
  // reduce, cheaper barriers) and the SM is filled by co-resident chains instead (measured, 8-50-1 B=32 x 4096 chains:
  // 2 warps per CTA 98 M steps/s, 4 warps 83 M, 8 warps 65 M).  Few chains: 8 warps for the latency of each one.
  pl->h1_nw = 0;
  if (h1_ok && nw_env == 0 && chains_hint >= 4 * 148) {
    const int w = rows / 16;
    pl->h1_nw = w >= 8 ? 8 : w >= 4 ? 4 : 2;
  }
  const int forced = pb_opt(PB_OPT_BT, 0);
  if (forced > 0) maxbt = pb_r32(forced) > 128 ? 128 : pb_r32(forced);
  // Two passes: first try to stay under half an SM's shared memory so that two CTAs are co-resident (more warps to
  // hide the latency of the short dependent phases) when there are more chains than SMs; then the full SM.
  const size_t lim2 = (size_t)(228 * 1024) / 2 - 1024, lim1 = PB_SMEM_LIMIT_BYTES;
  const int two = (chains_hint > 148 || pb_opt(PB_OPT_TWO_CTA, 0)) && !pb_opt(PB_OPT_ONE_CTA, 0);
  // Plans that end up with one CTA per SM anyway (large resident state) run 16 warps instead of 8: the phases are
  // short dependent sequences, 4 warps per scheduler hide their latency much better than 2.
  for (int pass = pb_opt(PB_OPT_FORCE_GSTATE, 0) ? 2 : (two ? 0 : 1); pass < 2; ++pass) {
    pl->NT = nw_env > 0 ? 32 * nw_env : (pass == 0 ? 256 : 512);
    const size_t limit = pass == 0 ? lim2 : lim1;
    for (int bt = maxbt; bt >= 32; bt -= 32) {
      if (pass == 0 && bt < (maxbt < 64 ? maxbt : 64)) break;
      if (h1_ok && bt == maxbt) {
        pb_layout(pl, sp, rows, nstate, bt, 1, 1, 0, 1);
        if ((size_t)pl->smem_floats * 4 <= limit) return PBNN_OK;
      }
      for (int res = want_res ? 1 : 0; res >= 0; --res)
        for (int kcap = 8; kcap >= 1; kcap /= 2) {
          pb_layout(pl, sp, rows, nstate, bt, res, kcap, (flags & PB_PLAN_ACC) != 0, 0);
          if ((size_t)pl->smem_floats * 4 <= limit) return PBNN_OK;
        }
      if (forced > 0) break;
    }
  }
  // last resort: parameter-sized arrays in a global-memory workspace (L2 resident), activations in shared memory
  pl->NT = nw_env > 0 ? 32 * nw_env : 512;
  for (int bt = maxbt; bt >= 32; bt -= 32) {
    for (int res = want_res ? 1 : 0; res >= 0; --res) {
      pb_layout(pl, sp, rows, nstate, bt, res, 1, (flags & PB_PLAN_ACC) != 0, 0, 1);
      if ((size_t)pl->smem_floats * 4 <= lim1) return PBNN_OK;
    }
    if (forced > 0) break;
  }
  snprintf(err, errn,
           "model does not fit the shared-memory resident engine (P=%d, %d resident arrays need %zu bytes > %d)", P,
           nstate, (size_t)pl->smem_floats * 4, PB_SMEM_LIMIT_BYTES);
  return PBNN_ETOOBIG;
}

enum { PB_ALGO_SGLD = 0, PB_ALGO_PSGLD = 1, PB_ALGO_SGHMC = 2, PB_ALGO_ADAM = 3, PB_ALGO_GRAD = 4 };

struct PbSgArgs {
  const float* X;
  const float* y;
  int N;
  const int32_t* idx;
  int idx_steps;
  int mb_mode, which_draw, init_draw;
  int contig;      // minibatch = rows 0..B-1 of X (BlackJAX-style explicit batch)
  int direct_key;  // keys[c] IS the step key (kernel-level API) instead of split(keys[c], T)[t]
  int data_size;   // likelihood scale numerator if > 0 (else N)
  int no_theta_store;
  float const_step;  // step size when step_sizes == NULL
  int B;
  uint32_t randint_mult;
  const uint32_t* keys;
  float* theta;
  float* aux1;  // pSGLD g | Adam m | GRAD: grad out
  float* aux2;  // pSGLD V | Adam v | GRAD: loglik out [C]
  int has_state;
  const float* step_sizes;
  int n_step_sizes;
  long long t0;
  int T, thin, burnin, T_out;
  float* trace;
  uint32_t* flags;
  float alpha, beta;
  float mom_mu0, mom_mu1, mom_sig0, mom_sig1;  // SGHMC momentum resampling: p = (mu0 + mu1 sqrt(eps)) + (sig0 + sig1 sqrt(eps)) z
  int L;
  float lr, b1, b2, eps;
  int count0;
  float scale;  // (N/B) / sigma^2
  const float* cv_center;      // control variates: gradient of the full-data log-posterior at the centring point [C][P]
  const float* cv_center_est;  //                   minibatch estimate at the centring point [C][P]
  float temperature;           // SGLD / SGHMC temperature (1 in pbnn; 0 = plain gradient step)
  float* ws;    // global workspace [C][nstate][Ps] when plan.gstate
  unsigned char* ws_img;  // plan.tc2: one operand-image block of plan.t2_img_bytes per CTA (1024-byte aligned)
  int n_chains;           // plan.tc2: the CTAs are persistent and loop over the chains
};

struct PbHmcArgs {
  const float* X;
  const float* y;
  int N;
  const uint32_t* keys;
  float* theta;
  const float* inv_mass;
  float step_size;
  int L;
  long long t0;
  int S, thin, burnin, S_out;
  float* trace;
  float* logp;
  int32_t* accept_count;
  float* info_delta;
  int32_t* info_accept;
  uint32_t* flags;
  float* ws;
  float* grad_out;  // optional: gradient of the log-posterior at the initial position [C][P]
  int no_store;     // do not write theta back (value_and_grad only)
  // balanced schedule of the tcgen05 kernel (C > number of SMs): the chains are cut into segments of seg_len
  // iterations, persistent CTAs take (segment, chain) items round-robin and hand a chain over through global memory
  int seg_len, n_chains;
  float* sched_g0;        // [C][P] gradient at the current position
  float* sched_logp;      // [C]
  int32_t* sched_acc;     // [C]
  uint32_t* sched_flags;  // [C]
  int* sched_done;        // [C] segments completed (release / acquire hand-over); [C] = next item to hand out
};

struct PbPredArgs {
  const float* samples;
  int S;
  const float* Xtest;
  int M;
  float* preds;
  float* part;  // [nchunk][2][M][s] partial sums, or NULL
  int nchunk, chunk;
  float* ws;
};
