// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM (sizing the attention softmax, profiles/README.md round 2).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../semivl_b200/csrc -o ldtm ldtm.cu && ./ldtm
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace svl;
template <int MODE>
__global__ void k(float* out, int iters) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&tptr), 256); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t v[32];
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) ptx::tmem_ld_32x32(base + c * 32, v);
      ptx::tmem_ld_wait();
      acc += v[0] + v[31];
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) ptx::tmem_st_32x32(base + c * 32, v);
      ptx::tmem_st_wait();
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tptr, 256); }
}
template <int MODE>
void run(const char* name, int threads) {
  float* out; cudaMalloc(&out, 148 * 2 * 1024 * 4);
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 2, threads>>>(out, 16);
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, threads>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double bytes = 148.0 * 2 * threads * 4.0 * 32 * 4 * iters;        // per SM: 2 CTAs x threads x 4 x (32 columns x 4 B)
  printf("%-6s %4d threads/CTA x 2 CTAs/SM: %8.3f ms  %.1f B/clk/SM at nominal %d MHz (err %s)\n", name, threads, ms, bytes / (ms * 1e-3) / (clk * 1e3) / 148.0,
         clk / 1000, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  run<0>("ld", 128); run<0>("ld", 256); run<0>("ld", 512);
  run<1>("st", 128); run<1>("st", 256);
  return 0;
}
