// Micro-benchmark: MUFU.EX2 / FFMA / FMNMX issue rates per SM on this part (sizing the attention softmax, profiles/README.md round 2).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = a * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
      if (MODE == 2) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));
      if (MODE == 3) asm volatile("add.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));
      if (MODE == 4) { unsigned u; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(x[i]), "f"(a)); x[i] = __uint_as_float(u); }
    }
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(out, 16, 0.5f, 0.25f);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(out, iters, 0.5f, 0.25f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double ops = 148.0 * 8 * 256 * 16.0 * iters;
  printf("%-8s %8.3f ms  %.2f ops/clk/SM at nominal %d MHz (thread-level ops)\n", name, ms, ops / (ms * 1e-3) / (clk * 1e3) / 148.0, clk / 1000);
}
int main() { run<0>("ex2"); run<1>("ffma"); run<2>("fmnmx"); run<3>("fadd"); run<4>("cvt.bf16x2"); return 0; }
