// Kernel-logic emulator -- TEST INFRASTRUCTURE ONLY (never loaded by the pbnn_b200 package).
//
// Compiles the *same* phase-structured kernel bodies as the CUDA build (pbnn_engine.cuh) with g++ and
// runs every phase as a loop over tid, one "CTA" after the other, with host pointers.  It exports the same
// C ABI as libpbnn_b200.so so the parity tests can exercise the index logic of the kernels on a machine
// without a GPU (`-m "not gpu"`).  It says nothing about the GPU's arithmetic or timing.
#define PB_EMU 1
#include <algorithm>
#include <vector>

#include "../../pbnn_b200/csrc/pbnn_engine.cuh"

static int pb_backend_num_sms() { return 1 << 30; }  // no tcgen05 plans here

static int pb_backend_sg(int algo, const PbPlan& pl, const PbSgArgs& a, int C, void*) {
  const int act = pl.act;
#pragma omp parallel for schedule(dynamic)
  for (int c = 0; c < C; ++c) {
    std::vector<float> sm(pl.smem_floats + 16, std::nanf(""));
    float* p = sm.data();
#define PB_CASE(ALGO)                                                        \
  case ALGO:                                                                 \
    if (act == PBNN_ACT_RELU) {                                              \
      if (pl.gstate) pb_sg_chain<ALGO, PBNN_ACT_RELU, true, 0>(pl, a, p, c, pl.NT);  \
      else pb_sg_chain<ALGO, PBNN_ACT_RELU, false, 0>(pl, a, p, c, pl.NT);      \
    } else {                                                                 \
      if (pl.gstate) pb_sg_chain<ALGO, PBNN_ACT_TANH, true, 0>(pl, a, p, c, pl.NT);  \
      else pb_sg_chain<ALGO, PBNN_ACT_TANH, false, 0>(pl, a, p, c, pl.NT);      \
    }                                                                        \
    break;
    switch (algo) {
      PB_CASE(PB_ALGO_SGLD)
      PB_CASE(PB_ALGO_PSGLD)
      PB_CASE(PB_ALGO_SGHMC)
      PB_CASE(PB_ALGO_ADAM)
      PB_CASE(PB_ALGO_GRAD)
    }
#undef PB_CASE
  }
  return PBNN_OK;
}

static int pb_backend_hmc(const PbPlan& pl, const PbHmcArgs& a, int C, void*) {
#pragma omp parallel for schedule(dynamic)
  for (int c = 0; c < C; ++c) {
    std::vector<float> sm(pl.smem_floats + 16, std::nanf(""));
    if (pl.act == PBNN_ACT_RELU) {
      if (pl.gstate) pb_hmc_chain<PBNN_ACT_RELU, true, 0>(pl, a, sm.data(), c, pl.NT);
      else pb_hmc_chain<PBNN_ACT_RELU, false, 0>(pl, a, sm.data(), c, pl.NT);
    } else {
      if (pl.gstate) pb_hmc_chain<PBNN_ACT_TANH, true, 0>(pl, a, sm.data(), c, pl.NT);
      else pb_hmc_chain<PBNN_ACT_TANH, false, 0>(pl, a, sm.data(), c, pl.NT);
    }
  }
  return PBNN_OK;
}

static int pb_backend_predict(const PbPlan& pl, const PbPredArgs& a, int mtiles, float* sum, float* sumsq, void*) {
#pragma omp parallel for schedule(dynamic) collapse(2)
  for (int mt = 0; mt < mtiles; ++mt)
    for (int ch = 0; ch < a.nchunk; ++ch) {
      std::vector<float> sm(pl.smem_floats + 16, std::nanf(""));
      if (pl.act == PBNN_ACT_RELU) {
        if (pl.gstate) pb_predict_block<PBNN_ACT_RELU, true>(pl, a, sm.data(), mt, ch, pl.NT);
        else pb_predict_block<PBNN_ACT_RELU, false>(pl, a, sm.data(), mt, ch, pl.NT);
      } else {
        if (pl.gstate) pb_predict_block<PBNN_ACT_TANH, true>(pl, a, sm.data(), mt, ch, pl.NT);
        else pb_predict_block<PBNN_ACT_TANH, false>(pl, a, sm.data(), mt, ch, pl.NT);
      }
    }
  if (sum) {
    const int n = a.M * pl.s;
    for (int i = 0; i < n; ++i) {
      float s = 0.f, q = 0.f;
      for (int ch = 0; ch < a.nchunk; ++ch) {
        s += a.part[(size_t)ch * 2 * n + i];
        q += a.part[(size_t)ch * 2 * n + n + i];
      }
      sum[i] = s;
      sumsq[i] = q;
    }
  }
  return PBNN_OK;
}

static int pb_backend_rng(int kind, const uint32_t* keys, int C, int n, void* out, void*) {
  for (int c = 0; c < C; ++c) {
    PbKey k{keys[2 * c], keys[2 * c + 1]};
    if (kind == 4) {
      const PbKey o = pb_split_at(k, (uint32_t)n);
      ((uint32_t*)out)[2 * c] = o.k0;
      ((uint32_t*)out)[2 * c + 1] = o.k1;
      continue;
    }
    for (int i = 0; i < n; ++i) {
      const size_t e = (size_t)c * n + i;
      if (kind == 0) {
        const PbKey o = pb_split_at(k, (uint32_t)i);
        ((uint32_t*)out)[2 * e] = o.k0;
        ((uint32_t*)out)[2 * e + 1] = o.k1;
      } else if (kind == 1) {
        ((uint32_t*)out)[e] = pb_bits_at(k, (uint32_t)i);
      } else if (kind == 2) {
        ((float*)out)[e] = fmaxf(0.f, pb_bits_to_uniform(pb_bits_at(k, (uint32_t)i)));
      } else {
        ((float*)out)[e] = pb_normal_at(k, (uint32_t)i);
      }
    }
  }
  return PBNN_OK;
}

static int pb_backend_mbidx(const uint32_t* keys, int C, int draw, int B, int N, uint32_t mult, int32_t* out, void*) {
  for (int c = 0; c < C; ++c) {
    PbKey k{keys[2 * c], keys[2 * c + 1]};
    for (int i = 0; i < draw; ++i) k = pb_split_at(k, 1u);
    const PbKey k1 = pb_split_at(k, 0u), k2 = pb_split_at(k, 1u);
    for (int b = 0; b < B; ++b) out[(size_t)c * B + b] = pb_randint_at(k1, k2, (uint32_t)b, (uint32_t)N, mult);
  }
  return PBNN_OK;
}

static int pb_backend_perm(const uint32_t* keys, int C, int N, int32_t* out, void*, void*) {
  const int rounds = (int)ceil(3.0 * log((double)(N > 1 ? N : 1)) / log(4294967295.0));
  for (int c = 0; c < C; ++c) {
    PbKey key{keys[2 * c], keys[2 * c + 1]};
    std::vector<int32_t> x(N), y(N);
    for (int i = 0; i < N; ++i) x[i] = i;
    for (int r = 0; r < rounds; ++r) {
      const PbKey sub = pb_split_at(key, 1u);
      key = pb_split_at(key, 0u);
      std::vector<unsigned long long> comp(N);
      for (int i = 0; i < N; ++i) comp[i] = ((unsigned long long)pb_bits_at(sub, (uint32_t)i) << 32) | (unsigned)i;
      std::sort(comp.begin(), comp.end());
      for (int j = 0; j < N; ++j) y[j] = x[(int)(comp[j] & 0xFFFFFFFFull)];
      x.swap(y);
    }
    for (int i = 0; i < N; ++i) out[(size_t)c * N + i] = x[i];
  }
  return PBNN_OK;
}

#include "../../pbnn_b200/csrc/pbnn_api.inc"

extern "C" int pbnn_thinning(const float* X, int n, int D, int size, int32_t* out, float* ws, void*) {
  if (!X || !out || !ws || n < 1 || D < 1 || size < 1 || size > n) return PBNN_EINVAL;
  float* xx = ws;
  float* k0 = ws + n;
  float* obj = ws + 2 * (size_t)n;
  auto dot = [&](int i, int j) {
    float a = 0.f;
    for (int k = 0; k < D; ++k) a += X[(size_t)i * D + k] * X[(size_t)j * D + k];
    return a;
  };
  for (int i = 0; i < n; ++i) xx[i] = dot(i, i);
  for (int i = 0; i < n; ++i) {
    float s = 0.f;
    for (int j = 0; j < n; ++j) s += sqrtf(xx[i]) + sqrtf(xx[j]) - sqrtf(fmaxf(xx[i] + xx[j] - 2.0f * dot(i, j), 0.f));
    k0[i] = s;
  }
  int cur = 0;
  for (int step = 0; step < size; ++step) {
    for (int i = 0; i < n; ++i) {
      if (step == 0) obj[i] = 2.0f * sqrtf(xx[i]) - 2.0f * (k0[i] / n);
      else {
        const float ki = sqrtf(xx[cur]) + sqrtf(xx[i]) - sqrtf(fmaxf(xx[cur] + xx[i] - 2.0f * dot(cur, i), 0.f));
        obj[i] = obj[i] + 2.0f * ki - 2.0f * (k0[i] / n);
      }
    }
    int bi = 0;
    for (int i = 1; i < n; ++i)
      if (obj[i] < obj[bi]) bi = i;
    cur = bi;
    out[step == 0 ? size - 1 : step - 1] = bi;
  }
  return PBNN_OK;
}
