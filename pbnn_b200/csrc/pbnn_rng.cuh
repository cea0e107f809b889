// threefry2x32 + the jax.random float transforms, bit-compatible with jax >= 0.5
// (jax_threefry_partitionable=True).  Sources restated: jax/_src/prng.py (threefry2x32,
// _threefry_split_foldlike, _threefry_random_bits_partitionable), jax/_src/random.py (_uniform,
// _normal_real, _randint), xla/hlo/builder/lib/math.cc (ErfInv f32).  pbnn call sites:
// src/pbnn/mcmc/langevin.py:69, src/pbnn/utils/data.py:70-76, blackjax generate_gaussian_noise.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef PB_DEV
#if defined(__CUDACC__)
#define PB_DEV __device__ __forceinline__
#else
#define PB_DEV static inline
#endif
#endif

struct PbKey {
  uint32_t k0, k1;
};

PB_DEV uint32_t pb_rotl(uint32_t x, int r) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(x, x, r);
#else
  return (x << r) | (x >> (32 - r));
#endif
}

#define PB_TF_ROUND(r)  \
  x0 += x1;             \
  x1 = pb_rotl(x1, r);  \
  x1 ^= x0;

PB_DEV PbKey pb_threefry(PbKey k, uint32_t x0, uint32_t x1) {
  const uint32_t ks0 = k.k0, ks1 = k.k1, ks2 = k.k0 ^ k.k1 ^ 0x1BD11BDAu;
  x0 += ks0;
  x1 += ks1;
  PB_TF_ROUND(13) PB_TF_ROUND(15) PB_TF_ROUND(26) PB_TF_ROUND(6)
  x0 += ks1;
  x1 += ks2 + 1u;
  PB_TF_ROUND(17) PB_TF_ROUND(29) PB_TF_ROUND(16) PB_TF_ROUND(24)
  x0 += ks2;
  x1 += ks0 + 2u;
  PB_TF_ROUND(13) PB_TF_ROUND(15) PB_TF_ROUND(26) PB_TF_ROUND(6)
  x0 += ks0;
  x1 += ks1 + 3u;
  PB_TF_ROUND(17) PB_TF_ROUND(29) PB_TF_ROUND(16) PB_TF_ROUND(24)
  x0 += ks1;
  x1 += ks2 + 4u;
  PB_TF_ROUND(13) PB_TF_ROUND(15) PB_TF_ROUND(26) PB_TF_ROUND(6)
  x0 += ks2;
  x1 += ks0 + 5u;
  PbKey o;
  o.k0 = x0;
  o.k1 = x1;
  return o;
}

// split(key, n)[i]  (n does not enter the value)
PB_DEV PbKey pb_split_at(PbKey k, uint32_t i) { return pb_threefry(k, 0u, i); }
// bits(key, (n,))[i]
PB_DEV uint32_t pb_bits_at(PbKey k, uint32_t i) {
  const PbKey o = pb_threefry(k, 0u, i);
  return o.k0 ^ o.k1;
}

PB_DEV float pb_u32_as_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

PB_DEV float pb_bits_to_uniform(uint32_t bits) { return pb_u32_as_float((bits >> 9) | 0x3F800000u) - 1.0f; }

// XLA ErfInv for float32 (Giles 2010 single-precision polynomial)
PB_DEV float pb_erfinv(float x) {
  float w = -log1pf(-(x * x));
  float p;
  if (w < 5.0f) {
    w = w - 2.5f;
    p = 2.81022636e-08f;
    p = 3.43273939e-07f + p * w;
    p = -3.5233877e-06f + p * w;
    p = -4.39150654e-06f + p * w;
    p = 0.00021858087f + p * w;
    p = -0.00125372503f + p * w;
    p = -0.00417768164f + p * w;
    p = 0.246640727f + p * w;
    p = 1.50140941f + p * w;
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    p = 0.000100950558f + p * w;
    p = 0.00134934322f + p * w;
    p = -0.00367342844f + p * w;
    p = 0.00573950773f + p * w;
    p = -0.0076224613f + p * w;
    p = 0.00943887047f + p * w;
    p = 1.00167406f + p * w;
    p = 2.83297682f + p * w;
  }
  return p * x;
}

// jax.random.normal float32: sqrt(2) * erf_inv(u in (-1, 1))
PB_DEV float pb_bits_to_normal(uint32_t bits) {
  const float lo = -0.99999994f;  // nextafter(-1, 0)
  const float u = pb_bits_to_uniform(bits);
  const float v = fmaxf(lo, u * 2.0f + lo);
  return 1.41421356237f * pb_erfinv(v);
}

PB_DEV float pb_normal_at(PbKey k, uint32_t i) { return pb_bits_to_normal(pb_bits_at(k, i)); }

// jax.random.randint(key, (B,), 0, span)[b]; k1,k2 = split(key); mult = (2^16 % span)^2 % span
PB_DEV int pb_randint_at(PbKey k1, PbKey k2, uint32_t b, uint32_t span, uint32_t mult) {
  const uint32_t hi = pb_bits_at(k1, b), lo = pb_bits_at(k2, b);
  uint32_t off = (hi % span) * mult + (lo % span);  // uint32 wrap-around, as lax does
  return (int)(off % span);
}
