/*
 * C restatement of pbnn's hot path -- TEST INFRASTRUCTURE ONLY (parity checker + timed CPU baseline).
 *
 * Parity status: "parity unpinned" by the reference (its test-suite is `assert True`,
 * /root/reference/tests/test_dummpy.py:1-2, and JAX/BlackJAX cannot run here).  This file is validated against
 * oracle/pbnn_oracle.py (tests/test_oracle.py), whose RNG layer is pinned to the Random123 KATs and to outputs
 * printed in the JAX documentation.
 *
 * One scalar float32 chain per call of the *_chain functions, OpenMP over chains in the *_run entry points --
 * the same decomposition JAX-on-CPU would get from running C independent reference calls.
 * Formulas: src/pbnn/mcmc/langevin.py:21-140, hamiltonian.py:18-140, mcmc/sgmcmc/diffusions.py:33-60,
 * mcmc/sgmcmc/psgld.py:25-48, utils/data.py:47-78, utils/misc.py:34-55, deep_ensembles.py:59-75,
 * models.py:16-35, benchmark/pipelines/logprobs.py:6-17, benchmark/pipelines/hmc.py:59-62; blackjax 1.2.5
 * sgld/sghmc/hmc and jax.random (threefry2x32, partitionable) restated from their published sources.
 * Flat parameter order = ravel_pytree (bias before kernel, kernel row-major [in][out]).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXL 8

typedef struct {
  int nl;
  int dims[MAXL + 1];
  int act; /* 0 relu, 1 tanh */
  float sigma, prior_scale;
} Spec;

/* ---- threefry2x32 (jax/_src/prng.py) ---- */
static inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
#define RND(r) x0 += x1; x1 = rotl(x1, r); x1 ^= x0;
static inline void tf(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* o0, uint32_t* o1) {
  uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += k0; x1 += k1;
  RND(13) RND(15) RND(26) RND(6) x0 += k1; x1 += k2 + 1u;
  RND(17) RND(29) RND(16) RND(24) x0 += k2; x1 += k0 + 2u;
  RND(13) RND(15) RND(26) RND(6) x0 += k0; x1 += k1 + 3u;
  RND(17) RND(29) RND(16) RND(24) x0 += k1; x1 += k2 + 4u;
  RND(13) RND(15) RND(26) RND(6) x0 += k2; x1 += k0 + 5u;
  *o0 = x0; *o1 = x1;
}
static inline uint32_t bits_at(const uint32_t k[2], uint32_t i) {
  uint32_t a, b;
  tf(k[0], k[1], 0u, i, &a, &b);
  return a ^ b;
}
static inline void split_at(const uint32_t k[2], uint32_t i, uint32_t out[2]) { tf(k[0], k[1], 0u, i, &out[0], &out[1]); }
static inline float u01(uint32_t bits) {
  uint32_t u = (bits >> 9) | 0x3F800000u;
  float f;
  memcpy(&f, &u, 4);
  return f - 1.0f;
}
/* XLA ErfInv f32 + jax.random.normal */
static inline float erfinv32(float x) {
  float w = -log1pf(-(x * x)), p;
  if (w < 5.0f) {
    w -= 2.5f;
    p = 2.81022636e-08f; p = 3.43273939e-07f + p * w; p = -3.5233877e-06f + p * w; p = -4.39150654e-06f + p * w;
    p = 0.00021858087f + p * w; p = -0.00125372503f + p * w; p = -0.00417768164f + p * w; p = 0.246640727f + p * w;
    p = 1.50140941f + p * w;
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f; p = 0.000100950558f + p * w; p = 0.00134934322f + p * w; p = -0.00367342844f + p * w;
    p = 0.00573950773f + p * w; p = -0.0076224613f + p * w; p = 0.00943887047f + p * w; p = 1.00167406f + p * w;
    p = 2.83297682f + p * w;
  }
  return p * x;
}
static inline float normal_at(const uint32_t k[2], uint32_t i) {
  const float lo = -0.99999994f;
  float v = u01(bits_at(k, i)) * 2.0f + lo;
  if (v < lo) v = lo;
  return 1.41421356237f * erfinv32(v);
}
static void randint_fill(const uint32_t key[2], int B, int N, int32_t* out) {
  uint32_t k1[2], k2[2];
  split_at(key, 0, k1);
  split_at(key, 1, k2);
  uint32_t span = (uint32_t)N, m = 65536u % span;
  uint32_t mult = (m * m) % span; /* uint32 wrap, as lax.mul does */
  for (int b = 0; b < B; ++b) {
    uint32_t hi = bits_at(k1, (uint32_t)b), lo = bits_at(k2, (uint32_t)b);
    uint32_t off = (hi % span) * mult + (lo % span);
    out[b] = (int32_t)(off % span);
  }
}
/* draw-th next(batch_labeled_data) (utils/data.py:69-78) */
static void reference_draw(const uint32_t key[2], int draw, int B, int N, int32_t* out) {
  uint32_t k[2] = {key[0], key[1]}, t[2];
  for (int i = 0; i < draw; ++i) { split_at(k, 1, t); k[0] = t[0]; k[1] = t[1]; }
  randint_fill(k, B, N, out);
}

/* ---- MLP forward/backward for a batch gathered by idx (idx == NULL: rows 0..B-1) ---- */
typedef struct {
  int P, maxw;
  int boff[MAXL], woff[MAXL];
  float* act[MAXL + 1]; /* act[l]: [B][dims[l]] */
  float* dz;            /* [B][maxw] */
  float* dh;
  int Bcap;
} Work;

static int num_params(const Spec* s) {
  int P = 0;
  for (int l = 0; l < s->nl; ++l) P += s->dims[l + 1] * (s->dims[l] + 1);
  return P;
}
static void work_init(Work* w, const Spec* s, int B) {
  int off = 0;
  w->maxw = 0;
  for (int l = 0; l < s->nl; ++l) {
    w->boff[l] = off;
    w->woff[l] = off + s->dims[l + 1];
    off += s->dims[l + 1] * (s->dims[l] + 1);
  }
  w->P = off;
  for (int l = 0; l <= s->nl; ++l) {
    if (s->dims[l] > w->maxw) w->maxw = s->dims[l];
    w->act[l] = (float*)malloc(sizeof(float) * (size_t)B * s->dims[l]);
  }
  w->dz = (float*)malloc(sizeof(float) * (size_t)B * w->maxw);
  w->dh = (float*)malloc(sizeof(float) * (size_t)B * w->maxw);
  w->Bcap = B;
}
static void work_free(Work* w, const Spec* s) {
  for (int l = 0; l <= s->nl; ++l) free(w->act[l]);
  free(w->dz);
  free(w->dh);
}

static void forward(const Spec* s, Work* w, const float* th, const float* X, const int32_t* idx, int row0, int B) {
  const int d = s->dims[0];
  for (int b = 0; b < B; ++b) {
    const float* x = X + (size_t)(idx ? idx[b] : row0 + b) * d;
    memcpy(w->act[0] + (size_t)b * d, x, sizeof(float) * d);
  }
  for (int l = 0; l < s->nl; ++l) {
    const int in = s->dims[l], out = s->dims[l + 1];
    const float* bias = th + w->boff[l];
    const float* W = th + w->woff[l];
    for (int b = 0; b < B; ++b) {
      const float* h = w->act[l] + (size_t)b * in;
      float* z = w->act[l + 1] + (size_t)b * out;
      for (int o = 0; o < out; ++o) z[o] = bias[o];
      for (int i = 0; i < in; ++i) {
        const float hi = h[i];
        const float* wr = W + (size_t)i * out;
        for (int o = 0; o < out; ++o) z[o] += hi * wr[o];
      }
      if (l < s->nl - 1) {
        if (s->act == 0) for (int o = 0; o < out; ++o) z[o] = z[o] > 0.f ? z[o] : 0.f;
        else for (int o = 0; o < out; ++o) z[o] = tanhf(z[o]);
      }
    }
  }
}

/* g (+)= scale * d/dtheta sum_b loglik_b ; returns sum_b loglik_b.  accumulate=0 zeroes g first */
static float loglik_grad(const Spec* s, Work* w, const float* th, const float* X, const float* y, const int32_t* idx,
                         int row0, int B, float scale, float* g, int accumulate) {
  const int nl = s->nl, so = s->dims[nl];
  const float inv_s2 = 1.0f / (s->sigma * s->sigma);
  if (!accumulate) memset(g, 0, sizeof(float) * w->P);
  forward(s, w, th, X, idx, row0, B);
  float ll = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* yy = y + (size_t)(idx ? idx[b] : row0 + b) * so;
    const float* f = w->act[nl] + (size_t)b * so;
    for (int o = 0; o < so; ++o) {
      const float r = yy[o] - f[o];
      ll += -0.5f * r * r * inv_s2;
      w->dz[(size_t)b * so + o] = r * inv_s2 * scale;
    }
  }
  for (int l = nl - 1; l >= 0; --l) {
    const int in = s->dims[l], out = s->dims[l + 1];
    float* gb = g + w->boff[l];
    float* gW = g + w->woff[l];
    const float* W = th + w->woff[l];
    for (int b = 0; b < B; ++b) {
      const float* dz = w->dz + (size_t)b * out;
      const float* h = w->act[l] + (size_t)b * in;
      for (int o = 0; o < out; ++o) gb[o] += dz[o];
      for (int i = 0; i < in; ++i) {
        const float hi = h[i];
        float* gr = gW + (size_t)i * out;
        for (int o = 0; o < out; ++o) gr[o] += hi * dz[o];
      }
      if (l > 0) {
        float* dh = w->dh + (size_t)b * in;
        for (int i = 0; i < in; ++i) {
          const float* wr = W + (size_t)i * out;
          float a = 0.f;
          for (int o = 0; o < out; ++o) a += wr[o] * dz[o];
          const float hv = h[i];
          dh[i] = a * (s->act == 0 ? (hv > 0.f ? 1.f : 0.f) : 1.f - hv * hv);
        }
      }
    }
    if (l > 0) {
      float* t = w->dz; w->dz = w->dh; w->dh = t;
    }
  }
  return ll;
}

/* gradient of logprior + N * mean loglik (blackjax grad_estimator == utils/misc.py:48-53) */
static void grad_estimator(const Spec* s, Work* w, const float* th, const float* X, const float* y, const int32_t* idx,
                           int B, int N, float* g) {
  loglik_grad(s, w, th, X, y, idx, 0, B, (float)N / (float)B, g, 0);
  const float ipv = 1.0f / (s->prior_scale * s->prior_scale);
  for (int j = 0; j < w->P; ++j) g[j] -= th[j] * ipv;
}

static float logpost_grad(const Spec* s, Work* w, const float* th, const float* X, const float* y, int N, int tile,
                          float* g) {
  float ll = 0.f;
  for (int r0 = 0; r0 < N; r0 += tile) {
    const int nb = N - r0 < tile ? N - r0 : tile;
    ll += loglik_grad(s, w, th, X, y, NULL, r0, nb, 1.0f, g, r0 > 0);
  }
  const float ipv = 1.0f / (s->prior_scale * s->prior_scale);
  float t2 = 0.f;
  for (int j = 0; j < w->P; ++j) {
    g[j] -= th[j] * ipv;
    t2 += th[j] * th[j];
  }
  return (-0.5f * ipv * t2 - (float)w->P * (0.918938533204672742f + logf(s->prior_scale))) + ll;
}

int pbo_num_params(const Spec* s) { return num_params(s); }
void pbo_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int pbo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void pbo_normal(const uint32_t* keys, int C, int n, float* out) {
  for (int c = 0; c < C; ++c)
    for (int i = 0; i < n; ++i) out[(size_t)c * n + i] = normal_at(keys + 2 * c, (uint32_t)i);
}
void pbo_reference_draw(const uint32_t* keys, int C, int draw, int B, int N, int32_t* out) {
  for (int c = 0; c < C; ++c) reference_draw(keys + 2 * c, draw, B, N, out + (size_t)c * B);
}

/* algo: 0 sgld, 1 psgld, 2 sghmc.  trace [C][T][P] or NULL.  Fixed minibatch = draw #2.
 * mom[4]: SGHMC momentum resampling p = (mom[0] + mom[1] sqrt(step)) + (mom[2] + mom[3] sqrt(step)) z -- the upstream
 * blackjax.sghmc body is not vendored (src/pbnn/mcmc/hamiltonian.py:118-128 is only the call site), so the variant is
 * a parameter; NULL = unit variance, zero mean. */
void pbo_sg_run(const Spec* s, int algo, const float* X, const float* y, int N, int B, const uint32_t* keys,
                float* theta, float step, int T, int C, int L, float alpha, float beta, const float* mom,
                float* trace) {
  const float mom_mu = mom ? mom[0] + mom[1] * sqrtf(step) : 0.0f;
  const float mom_sig = mom ? mom[2] + mom[3] * sqrtf(step) : 1.0f;
  const int P = num_params(s);
#pragma omp parallel for schedule(dynamic)
  for (int c = 0; c < C; ++c) {
    Work w;
    work_init(&w, s, B);
    float* th = theta + (size_t)c * P;
    float* g = (float*)malloc(sizeof(float) * P);
    float* a1 = (float*)calloc(P, sizeof(float));
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * B);
    const uint32_t* key = keys + 2 * c;
    if (algo == 1) {
      reference_draw(key, 1, B, N, idx);
      grad_estimator(s, &w, th, X, y, idx, B, N, g);
      for (int j = 0; j < P; ++j) a1[j] = g[j] * g[j];
    }
    reference_draw(key, 2, B, N, idx);
    for (int t = 0; t < T; ++t) {
      uint32_t kt[2];
      split_at(key, (uint32_t)t, kt);
      if (algo == 0) {
        grad_estimator(s, &w, th, X, y, idx, B, N, g);
        const float ns = sqrtf(2.0f * step);
        for (int j = 0; j < P; ++j) th[j] = th[j] + step * g[j] + ns * normal_at(kt, (uint32_t)j);
      } else if (algo == 1) {
        for (int j = 0; j < P; ++j) {
          const float pre = 1.0f / (sqrtf(a1[j]) + 1e-6f);
          th[j] = th[j] + step * pre * g[j] + sqrtf(2.0f * pre * step) * normal_at(kt, (uint32_t)j);
        }
        grad_estimator(s, &w, th, X, y, idx, B, N, g);
        for (int j = 0; j < P; ++j) a1[j] = alpha * a1[j] + (1.0f - alpha) * (g[j] * g[j]);
      } else {
        for (int j = 0; j < P; ++j) a1[j] = mom_mu + mom_sig * normal_at(kt, (uint32_t)j);
        const float ns = sqrtf(2.0f * step * (alpha - beta)), fr = 1.0f - alpha;
        for (int l = 0; l < L; ++l) {
          uint32_t kl[2];
          split_at(kt, (uint32_t)l, kl);
          grad_estimator(s, &w, th, X, y, idx, B, N, g);
          for (int j = 0; j < P; ++j) {
            const float p = a1[j];
            th[j] = th[j] + p;
            a1[j] = fr * p + step * g[j] + ns * normal_at(kl, (uint32_t)j);
          }
        }
      }
      if (trace) memcpy(trace + ((size_t)c * T + t) * P, th, sizeof(float) * P);
    }
    free(g); free(a1); free(idx);
    work_free(&w, s);
  }
}

/* HMC (hamiltonian.py:18-77 + blackjax.hmc): trace [C][S][P] or NULL; accept [C][S] or NULL */
void pbo_hmc_run(const Spec* s, const float* X, const float* y, int N, const uint32_t* keys, float* theta,
                 const float* inv_mass, float step, int L, int S, int C, float* trace, int32_t* accept) {
  const int P = num_params(s);
  const int tile = N < 128 ? N : 128;
#pragma omp parallel for schedule(dynamic)
  for (int c = 0; c < C; ++c) {
    Work w;
    work_init(&w, s, tile);
    float* th = theta + (size_t)c * P;
    float* g = (float*)malloc(sizeof(float) * P);
    float* th1 = (float*)malloc(sizeof(float) * P);
    float* g1 = (float*)malloc(sizeof(float) * P);
    float* p = (float*)malloc(sizeof(float) * P);
    const uint32_t* key = keys + 2 * c;
    const float heps = step * 0.5f;
    float lp = logpost_grad(s, &w, th, X, y, N, tile, g);
    for (int it = 0; it < S; ++it) {
      uint32_t ks[2], km[2], ka[2];
      split_at(key, (uint32_t)it, ks);
      split_at(ks, 0, km);
      split_at(ks, 1, ka);
      float ke0 = 0.f;
      for (int j = 0; j < P; ++j) {
        p[j] = sqrtf(1.0f / inv_mass[j]) * normal_at(km, (uint32_t)j);
        ke0 += (inv_mass[j] * p[j]) * p[j];
        th1[j] = th[j];
        g1[j] = g[j];
      }
      float lp1 = lp;
      for (int l = 0; l < L; ++l) {
        for (int j = 0; j < P; ++j) {
          p[j] = p[j] + heps * g1[j];
          th1[j] = th1[j] + step * (inv_mass[j] * p[j]);
        }
        lp1 = logpost_grad(s, &w, th1, X, y, N, tile, g1);
        for (int j = 0; j < P; ++j) p[j] = p[j] + heps * g1[j];
      }
      float ke1 = 0.f;
      for (int j = 0; j < P; ++j) ke1 += (inv_mass[j] * p[j]) * p[j];
      float delta = (-lp + 0.5f * ke0) - (-lp1 + 0.5f * ke1);
      if (delta != delta) delta = -INFINITY;
      const float pacc = fminf(1.0f, expf(delta));
      const float u = u01(bits_at(ka, 0u));
      const int acc = u < pacc;
      if (acc) {
        memcpy(th, th1, sizeof(float) * P);
        memcpy(g, g1, sizeof(float) * P);
        lp = lp1;
      }
      if (accept) accept[(size_t)c * S + it] = acc;
      if (trace) memcpy(trace + ((size_t)c * S + it) * P, th, sizeof(float) * P);
    }
    free(g); free(th1); free(g1); free(p);
    work_free(&w, s);
  }
}

/* Adam members (deep_ensembles.py:59-75 + optax.adam): idx [C][n_steps][B] */
void pbo_adam_run(const Spec* s, const float* X, const float* y, int N, const int32_t* idx, int n_steps, int B,
                  float* theta, float lr, int C) {
  const int P = num_params(s);
#pragma omp parallel for schedule(dynamic)
  for (int c = 0; c < C; ++c) {
    Work w;
    work_init(&w, s, B);
    float* th = theta + (size_t)c * P;
    float* g = (float*)malloc(sizeof(float) * P);
    float* m = (float*)calloc(P, sizeof(float));
    float* v = (float*)calloc(P, sizeof(float));
    for (int t = 0; t < n_steps; ++t) {
      grad_estimator(s, &w, th, X, y, idx + ((size_t)c * n_steps + t) * B, B, N, g);
      const float cnt = (float)(t + 1);
      const float bc1 = 1.0f - powf(0.9f, cnt), bc2 = 1.0f - powf(0.999f, cnt);
      for (int j = 0; j < P; ++j) {
        const float gj = -g[j];
        m[j] = (1.0f - 0.9f) * gj + 0.9f * m[j];
        v[j] = (1.0f - 0.999f) * (gj * gj) + 0.999f * v[j];
        th[j] = th[j] + (-lr) * ((m[j] / bc1) / (sqrtf(v[j] / bc2) + 1e-8f));
      }
    }
    free(g); free(m); free(v);
    work_free(&w, s);
  }
}

/* predictive sums over samples: sum/sumsq [M][out] */
void pbo_predict_moments(const Spec* s, const float* samples, int S, const float* Xt, int M, float* sum, float* sumsq) {
  const int P = num_params(s), so = s->dims[s->nl];
  const int tile = 128;
  const int ntile = (M + tile - 1) / tile;
#pragma omp parallel for schedule(dynamic)
  for (int t = 0; t < ntile; ++t) {
    Work w;
    work_init(&w, s, tile);
    const int r0 = t * tile, nb = M - r0 < tile ? M - r0 : tile;
    for (int i = 0; i < nb * so; ++i) { sum[(size_t)r0 * so + i] = 0.f; sumsq[(size_t)r0 * so + i] = 0.f; }
    for (int k = 0; k < S; ++k) {
      forward(s, &w, samples + (size_t)k * P, Xt, NULL, r0, nb);
      for (int i = 0; i < nb * so; ++i) {
        const float f = w.act[s->nl][i];
        sum[(size_t)r0 * so + i] += f;
        sumsq[(size_t)r0 * so + i] += f * f;
      }
    }
    work_free(&w, s);
  }
}
