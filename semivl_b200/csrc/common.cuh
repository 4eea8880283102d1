// Shared helpers for the semivl_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/semivl_b200.h"

namespace svl {

void set_error(const char* fmt, ...);

#define SVL_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      svl::set_error(__VA_ARGS__);        \
      return SVL_ERR_INVALID;             \
    }                                     \
  } while (0)

#define SVL_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      svl::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return SVL_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define SVL_LAUNCH_CHECK()                                                         \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      svl::set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return SVL_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int pow2ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// ---- dtype helpers -----------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Activation tensors come in three storage types (include/semivl_b200.h, svl_dtype):
//   F32, BF16, and BF16X2 = split pair: value = hi + lo, hi at [row*ld + col], lo at [row*ld + ld/2 + col].
// BF16X2 is the "precise" mode operand format: three bf16 tensor-core products (hi*hi + hi*lo + lo*hi)
// reproduce an fp32 contraction to ~2^-17 relative.
__device__ __forceinline__ float load_as_f32(const void* p, int dtype, int64_t i, int64_t lo_off = 0) {
  if (dtype == SVL_F32) return ((const float*)p)[i];
  float v = __bfloat162float(((const __nv_bfloat16*)p)[i]);
  if (dtype == SVL_BF16X2) v += __bfloat162float(((const __nv_bfloat16*)p)[i + lo_off]);
  return v;
}
__device__ __forceinline__ void store_from_f32(void* p, int dtype, int64_t i, float v, int64_t lo_off = 0) {
  if (dtype == SVL_F32) { ((float*)p)[i] = v; return; }
  __nv_bfloat16 hi = __float2bfloat16_rn(v);
  ((__nv_bfloat16*)p)[i] = hi;
  if (dtype == SVL_BF16X2) ((__nv_bfloat16*)p)[i + lo_off] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__device__ __forceinline__ void bf16x8_to_f32(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 f32_to_bf16x8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
// 8 consecutive elements starting at element offset `off`; `cnt` (<= 8) of them are valid.  Vector path when aligned.
__device__ __forceinline__ void ld8(const void* p, int dtype, int64_t off, int64_t lo_off, int cnt, float* f) {
  if (dtype == SVL_F32) {
    const float* q = (const float*)p + off;
    if (cnt == 8 && ((uintptr_t)q & 15) == 0) {
      float4 a = ((const float4*)q)[0], b = ((const float4*)q)[1];
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = i < cnt ? q[i] : 0.f;
    }
    return;
  }
  const __nv_bfloat16* q = (const __nv_bfloat16*)p + off;
  if (cnt == 8 && ((uintptr_t)q & 15) == 0 && (dtype == SVL_BF16 || (lo_off & 7) == 0)) {
    bf16x8_to_f32(*(const uint4*)q, f);
    if (dtype == SVL_BF16X2) {
      float g[8];
      bf16x8_to_f32(*(const uint4*)(q + lo_off), g);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += g[i];
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = 0.f;
      if (i < cnt) {
        v = __bfloat162float(q[i]);
        if (dtype == SVL_BF16X2) v += __bfloat162float(q[i + lo_off]);
      }
      f[i] = v;
    }
  }
}
__device__ __forceinline__ void st8(void* p, int dtype, int64_t off, int64_t lo_off, int cnt, const float* f) {
  if (dtype == SVL_F32) {
    float* q = (float*)p + off;
    if (cnt == 8 && ((uintptr_t)q & 15) == 0) {
      ((float4*)q)[0] = make_float4(f[0], f[1], f[2], f[3]);
      ((float4*)q)[1] = make_float4(f[4], f[5], f[6], f[7]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) if (i < cnt) q[i] = f[i];
    }
    return;
  }
  __nv_bfloat16* q = (__nv_bfloat16*)p + off;
  float lo[8];
  uint4 hi = f32_to_bf16x8(f);
  if (dtype == SVL_BF16X2) {
    float h[8];
    bf16x8_to_f32(hi, h);
#pragma unroll
    for (int i = 0; i < 8; ++i) lo[i] = f[i] - h[i];
  }
  if (cnt == 8 && ((uintptr_t)q & 15) == 0 && (dtype == SVL_BF16 || (lo_off & 7) == 0)) {
    *(uint4*)q = hi;
    if (dtype == SVL_BF16X2) *(uint4*)(q + lo_off) = f32_to_bf16x8(lo);
  } else {
    const __nv_bfloat16* hs = reinterpret_cast<const __nv_bfloat16*>(&hi);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < cnt) {
        q[i] = hs[i];
        if (dtype == SVL_BF16X2) q[i + lo_off] = __float2bfloat16_rn(lo[i]);
      }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum for blockDim.x <= 1024 (multiple of 32); `red` is >= 32 floats of shared memory
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  return red[0];
}

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// 256-bit global stores (sm_100): one full 32-byte sector per thread per instruction
__device__ __forceinline__ void st_global_256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]),
               "r"(r[6]), "r"(r[7])
               : "memory");
}
// 16 consecutive elements (fast path of the GEMM epilogue): BF16 -> one 32-byte store, F32 -> two; otherwise two st8
__device__ __forceinline__ void st16(void* p, int dtype, int64_t off, int64_t lo_off, int cnt, const float* f) {
  if (cnt == 16 && dtype == SVL_BF16) {
    __nv_bfloat16* q = (__nv_bfloat16*)p + off;
    if (((uintptr_t)q & 31) == 0) {
      uint32_t r[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        r[i] = *(uint32_t*)&h;
      }
      st_global_256(q, r);
      return;
    }
  }
  if (cnt == 16 && dtype == SVL_F32) {
    float* q = (float*)p + off;
    if (((uintptr_t)q & 31) == 0) {
      st_global_256(q, (const uint32_t*)f);
      st_global_256(q + 8, (const uint32_t*)(f + 8));
      return;
    }
  }
  st8(p, dtype, off, lo_off, cnt < 8 ? cnt : 8, f);
  if (cnt > 8) st8(p, dtype, off + 8, lo_off, cnt - 8, f + 8);
}
__device__ __forceinline__ void ld16(const void* p, int dtype, int64_t off, int64_t lo_off, int cnt, float* f) {
  ld8(p, dtype, off, lo_off, cnt < 8 ? cnt : 8, f);
  if (cnt > 8) ld8(p, dtype, off + 8, lo_off, cnt - 8, f + 8);
  else {
#pragma unroll
    for (int i = 8; i < 16; ++i) f[i] = 0.f;
  }
}

// Fast GELU for the bf16 throughput-mode epilogues (the epilogue of the FFN1 GEMM is bound by its ALU work: every instruction counts).
// Phi(x) = 1 - erfc(x / sqrt 2) / 2 with Abramowitz-Stegun 7.1.25: erfc(u) = (a1 t + a2 t^2 + a3 t^3) exp(-u^2), t = 1 / (1 + p u), u >= 0,
// |error| <= 2.5e-5 (1.25e-5 on Phi: 1/300 of a bf16 rounding step); one MUFU.RCP + one MUFU.EX2 (ftz forms: no range fix-ups), constants
// folded so that everything starts from x; the same exponential is the Gaussian of gelu'.  The precise mode uses erff (gelu_exact).
__device__ __forceinline__ void phi_exp_fast(float x, float& cdf, float& gauss) {
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(fabsf(x), 0.47047f * 0.70710678118654752f, 1.f)));
  const float w = x * 0.84932180028801904f;                       // sqrt(log2(e) / 2): 2^(-w^2) = exp(-x^2 / 2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(w * -w));
  float pl = fmaf(0.5f * 0.7478556f, t, 0.5f * -0.0958798f);
  pl = fmaf(pl, t, 0.5f * 0.3480242f);
  const float h = pl * t * e;                                     // erfc(|x| / sqrt 2) / 2
  cdf = x >= 0.f ? 1.f - h : h;
  gauss = e;
}
__device__ __forceinline__ float gelu_fast(float x) {
  float cdf, e;
  phi_exp_fast(x, cdf, e);
  return x * cdf;
}
__device__ __forceinline__ void gelu_and_grad_fast(float x, float& g, float& dg) {
  float cdf, e;
  phi_exp_fast(x, cdf, e);
  g = x * cdf;
  dg = fmaf(x * 0.3989422804014327f, e, cdf);
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  float cdf, e;
  phi_exp_fast(x, cdf, e);
  return fmaf(x * 0.3989422804014327f, e, cdf);
}

}  // namespace svl
