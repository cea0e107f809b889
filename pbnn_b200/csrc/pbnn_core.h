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
#define PB_MAX_NT 256

// One Dense layer in the resident (shared-memory) layout.
//   W_aug : rows x wp floats, rows = roundup4(in+1); row r<in is kernel row r, row `in` is the bias,
//           remaining rows and the columns >= out are zero pads that are never updated.
//   wp    : pitch, multiple of 4 with wp % 8 == 4 so that 8 consecutive rows hit 8 distinct 16-byte
//           bank groups (conflict-free LDS.128 in the k-contiguous GEMMs).
struct PbLayer {
  int in, out;
  int inA;   // in + 1
  int rows;  // roundup4(in + 1)
  int wp;
  int soff;  // float offset of W_aug inside a resident parameter array
  int fb, fw;  // flat (ravel_pytree) offsets of bias and kernel
};

struct PbPlan {
  int nl, act, P, Ps;
  int d, s;
  PbLayer L[PBNN_MAX_LAYERS];
  int BT, ap;                      // batch-tile rows (multiple of 32) and activation pitch (BT + 4)
  int aoff[PBNN_MAX_LAYERS + 2];   // A_0 (input tile, feature-major) .. A_nl (output / dZ_last), then yT
  int arows[PBNN_MAX_LAYERS + 2];
  int nstate;                      // resident parameter-sized arrays
  int off_act, off_scr, off_idx, off_scal;
  int idx_cap;
  int smem_floats;
  int NT;
  float inv_s2, inv_pv, logprior_const;
};

static inline int pb_r4(int x) { return (x + 3) & ~3; }
static inline int pb_r32(int x) { return (x + 31) & ~31; }

static inline int pb_env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

static inline int pb_num_params(const pbnn_mlp_spec* sp) {
  if (!sp || sp->n_layers < 1 || sp->n_layers > PBNN_MAX_LAYERS) return -1;
  long P = 0;
  for (int l = 0; l <= sp->n_layers; ++l)
    if (sp->dims[l] < 1 || sp->dims[l] > 4096) return -1;
  for (int l = 0; l < sp->n_layers; ++l) P += (long)sp->dims[l + 1] * (sp->dims[l] + 1);
  return P > (1L << 26) ? -1 : (int)P;
}

// rows: rows per gradient evaluation (minibatch B, N for HMC, M for predict); nstate: resident
// parameter-sized arrays; idx_cap: ints reserved for the minibatch index list.
static inline int pb_build_plan(const pbnn_mlp_spec* sp, int rows, int nstate, int idx_cap, PbPlan* pl, char* err,
                                size_t errn) {
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
  if (!(sp->noise_level > 0.f) || !(sp->prior_scale > 0.f)) {
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
  pl->inv_s2 = 1.0f / (sp->noise_level * sp->noise_level);
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
    Ly.fb = foff;  // ravel_pytree: "bias" < "kernel"
    Ly.fw = foff + Ly.out;
    foff += Ly.out * (Ly.in + 1);
  }
  pl->Ps = soff;
  pl->nstate = nstate;
  pl->idx_cap = pb_r4(idx_cap);
  pl->NT = 32 * pb_env_int("PBNN_NW", 8);
  if (pl->NT < 32 || pl->NT > PB_MAX_NT) {
    snprintf(err, errn, "PBNN_NW out of range");
    return PBNN_EINVAL;
  }
  int bt = pb_r32(rows);
  if (bt > 128) bt = 128;
  const int forced = pb_env_int("PBNN_BT", 0);
  if (forced > 0) bt = pb_r32(forced) > 128 ? 128 : pb_r32(forced);
  for (; bt >= 32; bt -= 32) {
    const int ap = bt + 4;
    int off = nstate * pl->Ps;
    pl->off_act = off;
    for (int l = 0; l <= pl->nl + 1; ++l) {
      int r;
      if (l < pl->nl) r = pb_r4(sp->dims[l] + 1);
      else r = pb_r4(pl->s);  // A_nl (output / dZ_last) and yT
      pl->aoff[l] = off;
      pl->arows[l] = r;
      off += r * ap;
    }
    pl->off_scr = off;
    off += 1024 + 64 + 128 + 128 + 2 * pb_r4(pl->s) * ap;  // reduction scratch, llbuf, fbuf, predictive accumulators
    pl->off_idx = off;
    off += pl->idx_cap;
    pl->off_scal = off;
    off += 32;
    pl->smem_floats = off;
    pl->BT = bt;
    pl->ap = ap;
    if ((size_t)off * 4 <= PB_SMEM_LIMIT_BYTES) return PBNN_OK;
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
  int L;
  float lr, b1, b2, eps;
  int count0;
  float scale;  // (N/B) / sigma^2
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
};

struct PbPredArgs {
  const float* samples;
  int S;
  const float* Xtest;
  int M;
  float* preds;
  float* part;  // [nchunk][2][M][s] partial sums, or NULL
  int nchunk, chunk;
};
