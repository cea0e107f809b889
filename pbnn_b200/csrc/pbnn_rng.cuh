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
PB_DEV float pb_erfinv_poly_bulk(float w) {  // w < 5 (after w -= 2.5)
  float p = 2.81022636e-08f;
  p = 3.43273939e-07f + p * w;
  p = -3.5233877e-06f + p * w;
  p = -4.39150654e-06f + p * w;
  p = 0.00021858087f + p * w;
  p = -0.00125372503f + p * w;
  p = -0.00417768164f + p * w;
  p = 0.246640727f + p * w;
  p = 1.50140941f + p * w;
  return p;
}
PB_DEV float pb_erfinv_poly_tail(float w) {  // w >= 5 (after w = sqrt(w) - 3)
  float p = -0.000200214257f;
  p = 0.000100950558f + p * w;
  p = 0.00134934322f + p * w;
  p = -0.00367342844f + p * w;
  p = 0.00573950773f + p * w;
  p = -0.0076224613f + p * w;
  p = 0.00943887047f + p * w;
  p = 1.00167406f + p * w;
  p = 2.83297682f + p * w;
  return p;
}

PB_DEV float pb_erfinv(float x) {
#if defined(__CUDA_ARCH__)
  // Branch-free (the chains of several parameters are interleaved by the compiler; a branch would serialise them).
  // w = -log1p(-x^2) as -log(1 + t), t = -(x*x): 1 + t is exact for |t| >= 0.5 and off by <= 6e-8 (absolute) below, and
  // lg2.approx (MUFU.LG2) * ln 2 is within 2^-21.4 absolute on [0.5, 2] / 3 ulp elsewhere; the result enters z only through
  // x * p(w) with |p'| <= 0.25, so z moves by < 3e-7 in the bulk -- inside the 1e-6 contract on normals (tests).  The
  // tail (w >= 5, 0.3 % of the draws) uses w * rsqrt(w): relative 2^-22, inside the 5e-6 relative contract there.
  float l2, rs;
  // x * x must be ROUNDED before the subtraction, as XLA's log1p(-(x * x)) sees it: contracted into an FMA the argument
  // is the exact 1 - x^2, which differs from the reference's by up to 2e-4 relative when |x| -> 1 (2e-5 on z at |z| ~ 4)
  const float xx = __fmul_rn(x, x);
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(__fadd_rn(1.0f, -xx)));  // argument in [2^-23, 1]: never denormal
  const float w = l2 * -0.693147180559945f;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(w));  // w == 0 -> inf -> NaN below: discarded by the select
  const float pb = pb_erfinv_poly_bulk(w - 2.5f);
  const float pt = pb_erfinv_poly_tail(w * rs - 3.0f);
  return (w < 5.0f ? pb : pt) * x;
#else
  float w = -log1pf(-(x * x));
  const float p = w < 5.0f ? pb_erfinv_poly_bulk(w - 2.5f) : pb_erfinv_poly_tail(sqrtf(w) - 3.0f);
  return p * x;
#endif
}

// jax.random.normal float32: sqrt(2) * erf_inv(u in (-1, 1))
PB_DEV float pb_bits_to_normal(uint32_t bits) {
  const float lo = -0.99999994f;  // nextafter(-1, 0)
  const float u = pb_bits_to_uniform(bits);
  const float v = fmaxf(lo, u * 2.0f + lo);
  return 1.41421356237f * pb_erfinv(v);
}

PB_DEV float pb_normal_at(PbKey k, uint32_t i) { return pb_bits_to_normal(pb_bits_at(k, i)); }

#if defined(__CUDA_ARCH__)
// Two draws at once with the two Giles polynomials evaluated as packed FP32 (fma.rn.f32x2: one issue slot for the same
// Horner step of both draws, coefficient broadcast) -- per component the arithmetic is exactly pb_erfinv's, so the
// results are bit-identical to pb_bits_to_normal; 16 FFMA2 instead of 32 FFMA per pair of draws.
PB_DEV void pb_bits_to_normal2(uint32_t bits0, uint32_t bits1, float& z0, float& z1) {
  const float lo = -0.99999994f;
  const float v0 = fmaxf(lo, pb_bits_to_uniform(bits0) * 2.0f + lo), v1 = fmaxf(lo, pb_bits_to_uniform(bits1) * 2.0f + lo);
  float l0, l1, r0, r1;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(__fadd_rn(1.0f, -__fmul_rn(v0, v0))));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l1) : "f"(__fadd_rn(1.0f, -__fmul_rn(v1, v1))));
  const float w0 = l0 * -0.693147180559945f, w1 = l1 * -0.693147180559945f;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(w0));
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(w1));
  const float2 wb = make_float2(w0 - 2.5f, w1 - 2.5f), wt = make_float2(w0 * r0 - 3.0f, w1 * r1 - 3.0f);
#define PB_H2(p, w, c) p = __ffma2_rn(p, w, make_float2(c, c))
  float2 pb = make_float2(2.81022636e-08f, 2.81022636e-08f);
  PB_H2(pb, wb, 3.43273939e-07f); PB_H2(pb, wb, -3.5233877e-06f); PB_H2(pb, wb, -4.39150654e-06f);
  PB_H2(pb, wb, 0.00021858087f); PB_H2(pb, wb, -0.00125372503f); PB_H2(pb, wb, -0.00417768164f);
  PB_H2(pb, wb, 0.246640727f); PB_H2(pb, wb, 1.50140941f);
  float2 pt = make_float2(-0.000200214257f, -0.000200214257f);
  PB_H2(pt, wt, 0.000100950558f); PB_H2(pt, wt, 0.00134934322f); PB_H2(pt, wt, -0.00367342844f);
  PB_H2(pt, wt, 0.00573950773f); PB_H2(pt, wt, -0.0076224613f); PB_H2(pt, wt, 0.00943887047f);
  PB_H2(pt, wt, 1.00167406f); PB_H2(pt, wt, 2.83297682f);
#undef PB_H2
  z0 = 1.41421356237f * ((w0 < 5.0f ? pb.x : pt.x) * v0);
  z1 = 1.41421356237f * ((w1 < 5.0f ? pb.y : pt.y) * v1);
}
#endif

// N draws normal(k, .)[idx[j]] with the N threefry chains interleaved IN SOURCE ORDER, round by round: one chain is
// ~70 dependent integer instructions, so a thread needs several in flight; left to itself the compiler serialises the
// chains as soon as the surrounding code is short of registers (measured: 3x slower drains in pbnn_tc2.cuh).
template <int N, bool PACK = false>
PB_DEV void pb_normal_n(PbKey k, const uint32_t* idx, float* z) {
  const uint32_t ks0 = k.k0, ks1 = k.k1, ks2 = k.k0 ^ k.k1 ^ 0x1BD11BDAu;
  uint32_t x0[N], x1[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    x0[j] = ks0;
    x1[j] = idx[j] + ks1;
  }
#define PB_TF_ROUND_N(r)                   \
  _Pragma("unroll") for (int j = 0; j < N; ++j) x0[j] += x1[j];                  \
  _Pragma("unroll") for (int j = 0; j < N; ++j) x1[j] = pb_rotl(x1[j], r) ^ x0[j];
#define PB_TF_INJECT_N(a, b, c)            \
  _Pragma("unroll") for (int j = 0; j < N; ++j) {                                \
    x0[j] += a;                            \
    x1[j] += b + c;                        \
  }
  PB_TF_ROUND_N(13) PB_TF_ROUND_N(15) PB_TF_ROUND_N(26) PB_TF_ROUND_N(6)
  PB_TF_INJECT_N(ks1, ks2, 1u)
  PB_TF_ROUND_N(17) PB_TF_ROUND_N(29) PB_TF_ROUND_N(16) PB_TF_ROUND_N(24)
  PB_TF_INJECT_N(ks2, ks0, 2u)
  PB_TF_ROUND_N(13) PB_TF_ROUND_N(15) PB_TF_ROUND_N(26) PB_TF_ROUND_N(6)
  PB_TF_INJECT_N(ks0, ks1, 3u)
  PB_TF_ROUND_N(17) PB_TF_ROUND_N(29) PB_TF_ROUND_N(16) PB_TF_ROUND_N(24)
  PB_TF_INJECT_N(ks1, ks2, 4u)
  PB_TF_ROUND_N(13) PB_TF_ROUND_N(15) PB_TF_ROUND_N(26) PB_TF_ROUND_N(6)
  PB_TF_INJECT_N(ks2, ks0, 5u)
#undef PB_TF_ROUND_N
#undef PB_TF_INJECT_N
#if defined(__CUDA_ARCH__)
  if constexpr (PACK && (N % 2) == 0) {
#pragma unroll
    for (int j = 0; j < N; j += 2) pb_bits_to_normal2(x0[j] ^ x1[j], x0[j + 1] ^ x1[j + 1], z[j], z[j + 1]);
  } else
#endif
  {
#pragma unroll
    for (int j = 0; j < N; ++j) z[j] = pb_bits_to_normal(x0[j] ^ x1[j]);
  }
}

// jax.random.randint(key, (B,), 0, span)[b]; k1,k2 = split(key); mult = (2^16 % span)^2 % span
PB_DEV int pb_randint_at(PbKey k1, PbKey k2, uint32_t b, uint32_t span, uint32_t mult) {
  const uint32_t hi = pb_bits_at(k1, b), lo = pb_bits_at(k2, b);
  uint32_t off = (hi % span) * mult + (lo % span);  // uint32 wrap-around, as lax does
  return (int)(off % span);
}
