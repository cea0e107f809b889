// Microbenchmark: throughput of jax.random.normal draws (threefry2x32 + erf_inv) per SM as a function of resident warps
// and of the number of interleaved chains per thread.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rng_bench rng_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../pbnn_b200/csrc/pbnn_rng.cuh"

template <int ILP>
__global__ void k(uint32_t k0, uint32_t k1, int iters, float* out) {
  PbKey key; key.k0 = k0; key.k1 = k1;
  float acc = 0.f;
  uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) * (uint32_t)(iters * ILP);
  for (int it = 0; it < iters; ++it) {
    float z[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) z[j] = pb_normal_at(key, base + it * ILP + j);
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc += z[j];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int ILP>
void run(int threads, float* out) {
  const int iters = 512 / ILP * 8;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<ILP><<<148, threads>>>(1, 2, iters, out);
  cudaEventRecord(a);
  k<ILP><<<148, threads>>>(1, 2, iters, out);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double normals_per_sm = (double)threads * iters * ILP;
  double cycles = ms * 1e-3 * 1.965e9;
  printf("ILP %d threads/SM %4d: %.3f ms  %.3f cycles per normal per SM (%.2f G normals/s chip)\n", ILP, threads, ms,
         cycles / normals_per_sm, 148 * normals_per_sm / (ms * 1e-3) / 1e9);
}

int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  for (int t : {128, 256, 320, 512, 1024}) { run<1>(t, out); run<4>(t, out); run<8>(t, out); }
  return 0;
}
