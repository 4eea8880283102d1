// HBM-bound token / normalisation kernels of the ViT path: patchify, token assembly, LayerNorm fwd/bwd,
// L2 normalisation fwd/bwd and small utilities (cast, column sums, batch sums).  One warp per row, 128-bit loads,
// warp-shuffle reductions, fp32 statistics.
#include "common.cuh"

namespace svl {
namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kMaxVec = 8;   // up to 8 float4 per lane -> rows up to 1024 wide

inline int grid_for_rows(int64_t rows) {
  int64_t b = cdiv(rows, kWarpsPerBlock);
  int64_t cap = 148 * 16;
  return (int)(b < cap ? b : cap);
}

// ---------------------------------------------------------------------------------------------- patchify
// one thread per 8 output columns (8 consecutive px of one (c, py) row): 2 float4 loads, one 16-byte store
__global__ void patchify_kernel(const float* __restrict__ img, void* __restrict__ out, int out_dtype, int b, int H, int W, int p,
                                int hp, int wp) {
  const int K = 3 * p * p;
  const int groups = K / 8;
  const int64_t total = (int64_t)b * hp * wp * groups;
  const int64_t ld = out_dtype == SVL_BF16X2 ? 2 * K : K;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(idx % groups);
    const int64_t row = idx / groups;
    const int col = g * 8;
    const int c = col / (p * p), py = (col / p) % p, px = col % p;
    const int pw = (int)(row % wp), ph = (int)((row / wp) % hp), bi = (int)(row / ((int64_t)wp * hp));
    const int y = ph * p + py, x0 = pw * p + px;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int x = x0 + i;
      f[i] = (y < H && x < W) ? __ldg(img + (((int64_t)bi * 3 + c) * H + y) * W + x) : 0.f;
    }
    st8(out, out_dtype, row * ld + col, ld / 2, 8, f);
  }
}

// ---------------------------------------------------------------------------------------------- token assembly
__global__ void assemble_tokens_kernel(const float* __restrict__ patches, const float* __restrict__ cls, const float* __restrict__ pos,
                                       float* __restrict__ x, int b, int hw, int c) {
  const int L = hw + 1;
  const int c4 = c / 4;
  const int64_t total = (int64_t)b * L * c4;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % c4);
    const int64_t t = idx / c4;
    const int l = (int)(t % L), bi = (int)(t / L);
    float4 pe = __ldg((const float4*)(pos + (int64_t)l * c) + j);
    float4 v = l == 0 ? __ldg((const float4*)cls + j) : __ldg((const float4*)(patches + ((int64_t)bi * hw + l - 1) * c) + j);
    v.x += pe.x; v.y += pe.y; v.z += pe.z; v.w += pe.w;
    ((float4*)(x + t * c))[j] = v;
  }
}

// ---------------------------------------------------------------------------------------------- position-table resize
// MaskClipVisionTransformer.resize_pos_embed (maskclip_vit.py:462-490): the cls row is copied, the [gh, gw, c] grid rows are resized to
// [oh, ow] with torch's bicubic kernel (A = -0.75, align_corners=False: src = (dst + 0.5) * in/out - 0.5 WITHOUT clamping, taps at
// floor(src) - 1 .. + 2 with the tap index clamped to the grid).  Forward gathers 16 taps per output row; backward scatters them
// with red.global.add (the position table is a parameter: its gradient tolerates the atomics' order like every weight gradient).
__device__ __forceinline__ void cubic_taps(int o, float scale, int in_size, int idx[4], float w[4]) {
  const float A = -0.75f;
  const float real = scale * ((float)o + 0.5f) - 0.5f;
  const float fl = floorf(real);
  const float t = real - fl;
  const int i0 = (int)fl;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  w[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int v = i0 - 1 + k;
    idx[k] = v < 0 ? 0 : (v > in_size - 1 ? in_size - 1 : v);
  }
}
// one thread = one float4 of one output row; row 0 is the cls entry
template <bool kBackward>
__global__ void pos_resize_kernel(const float* __restrict__ src, float* __restrict__ dst, int gh, int gw, int oh, int ow, int c) {
  const int c4 = c / 4;
  const int64_t total = (int64_t)(oh * ow + 1) * c4;
  const float sy = (float)gh / (float)oh, sx = (float)gw / (float)ow;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % c4);
    const int row = (int)(idx / c4);
    if (row == 0) {
      const float4 v = __ldg((const float4*)src + j);
      if (kBackward) {
        float* d = dst + 4 * j;
        atomicAdd(d, v.x); atomicAdd(d + 1, v.y); atomicAdd(d + 2, v.z); atomicAdd(d + 3, v.w);
      } else {
        ((float4*)dst)[j] = v;
      }
      continue;
    }
    const int oy = (row - 1) / ow, ox = (row - 1) % ow;
    int iy[4], ix[4];
    float wy[4], wx[4];
    cubic_taps(oy, sy, gh, iy, wy);
    cubic_taps(ox, sx, gw, ix, wx);
    if (kBackward) {
      const float4 g = __ldg((const float4*)(src + (int64_t)row * c) + j);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const float w = wy[a] * wx[b];
          float* d = dst + (int64_t)(1 + iy[a] * gw + ix[b]) * c + 4 * j;
          atomicAdd(d, w * g.x); atomicAdd(d + 1, w * g.y); atomicAdd(d + 2, w * g.z); atomicAdd(d + 3, w * g.w);
        }
    } else {
      // torch accumulates row by row: out = sum_a wy[a] * (sum_b wx[b] * in[iy[a], ix[b]])
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const float4 v = __ldg((const float4*)(src + (int64_t)(1 + iy[a] * gw + ix[b]) * c) + j);
          r.x += wx[b] * v.x; r.y += wx[b] * v.y; r.z += wx[b] * v.z; r.w += wx[b] * v.w;
        }
        acc.x += wy[a] * r.x; acc.y += wy[a] * r.y; acc.z += wy[a] * r.z; acc.w += wy[a] * r.w;
      }
      ((float4*)(dst + (int64_t)row * c))[j] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------- LayerNorm
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
layernorm_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                     void* __restrict__ y, int y_dtype, int64_t ldy, float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, int c,
                     float eps) {
  const int lane = threadIdx.x & 31;
  const int nv = c / 128;                      // float4 per lane
  for (int64_t row = blockIdx.x * (int64_t)kWarpsPerBlock + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * kWarpsPerBlock) {
    const float4* xr = (const float4*)(x + row * ldx);
    float4 v[kMaxVec];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nv) {
        v[i] = xr[i * 32 + lane];
        s += v[i].x + v[i].y + v[i].z + v[i].w;
      }
    const float mu = warp_sum(s) / c;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nv) {
        float a = v[i].x - mu, b = v[i].y - mu, cc = v[i].z - mu, d = v[i].w - mu;
        q += a * a + b * b + cc * cc + d * d;
      }
    const float rs = rsqrtf(warp_sum(q) / c + eps);
    if (lane == 0) {
      if (mean) mean[row] = mu;
      if (rstd) rstd[row] = rs;
    }
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nv) {
        const int col = (i * 32 + lane) * 4;
        float4 g = __ldg((const float4*)gamma + i * 32 + lane), bt = __ldg((const float4*)beta + i * 32 + lane);
        float o0 = (v[i].x - mu) * rs * g.x + bt.x, o1 = (v[i].y - mu) * rs * g.y + bt.y, o2 = (v[i].z - mu) * rs * g.z + bt.z,
              o3 = (v[i].w - mu) * rs * g.w + bt.w;
        if (y_dtype == SVL_F32) {
          ((float4*)((float*)y + row * ldy))[i * 32 + lane] = make_float4(o0, o1, o2, o3);
        } else {
          __nv_bfloat16* yr = (__nv_bfloat16*)y + row * ldy + col;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(o0, o1), h1 = __floats2bfloat162_rn(o2, o3);
          uint2 u;
          u.x = *(uint32_t*)&h0; u.y = *(uint32_t*)&h1;
          *(uint2*)yr = u;
          if (y_dtype == SVL_BF16X2) {
            float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
            __nv_bfloat162 l0 = __floats2bfloat162_rn(o0 - f0.x, o1 - f0.y), l1 = __floats2bfloat162_rn(o2 - f1.x, o3 - f1.y);
            u.x = *(uint32_t*)&l0; u.y = *(uint32_t*)&l1;
            *(uint2*)(yr + ldy / 2) = u;
          }
        }
      }
  }
}

// dx = dres1 + dres2 + LN'(dy); optional activation-format copy of dx; optional dgamma/dbeta accumulation.
// NV = float4 vectors per lane (row width / 128), WG = accumulate weight gradients.  Everything is 8/16-byte vector traffic.
__device__ __forceinline__ void load4_any(const void* p, int dtype, int64_t off, int64_t lo_off, float* d) {
  if (dtype == SVL_F32) {
    const float4 t = *(const float4*)((const float*)p + off);
    d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w;
  } else {
    const __nv_bfloat16* q = (const __nv_bfloat16*)p + off;
    uint2 u = *(const uint2*)q;
    float2 a = __bfloat1622float2(*(__nv_bfloat162*)&u.x), b = __bfloat1622float2(*(__nv_bfloat162*)&u.y);
    d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y;
    if (dtype == SVL_BF16X2) {
      u = *(const uint2*)(q + lo_off);
      a = __bfloat1622float2(*(__nv_bfloat162*)&u.x); b = __bfloat1622float2(*(__nv_bfloat162*)&u.y);
      d[0] += a.x; d[1] += a.y; d[2] += b.x; d[3] += b.y;
    }
  }
}
__device__ __forceinline__ void store4_act(void* p, int dtype, int64_t off, int64_t lo_off, const float* o) {
  if (dtype == SVL_F32) {
    *(float4*)((float*)p + off) = make_float4(o[0], o[1], o[2], o[3]);
    return;
  }
  __nv_bfloat16* q = (__nv_bfloat16*)p + off;
  __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
  uint2 u;
  u.x = *(uint32_t*)&h0; u.y = *(uint32_t*)&h1;
  *(uint2*)q = u;
  if (dtype == SVL_BF16X2) {
    float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    __nv_bfloat162 l0 = __floats2bfloat162_rn(o[0] - f0.x, o[1] - f0.y), l1 = __floats2bfloat162_rn(o[2] - f1.x, o[3] - f1.y);
    u.x = *(uint32_t*)&l0; u.y = *(uint32_t*)&l1;
    *(uint2*)(q + lo_off) = u;
  }
}

template <int NV, bool WG>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
layernorm_bwd_kernel(const void* __restrict__ dy, int dy_dtype, int64_t lddy, const float* __restrict__ x, int64_t ldx,
                     const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ dres1, const float* __restrict__ dres2, float* __restrict__ dx, void* __restrict__ dx_act,
                     int act_dtype, int64_t ld_act, float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows) {
  constexpr int c = NV * 128;
  __shared__ float sg[WG ? c : 1], sb[WG ? c : 1];
  const int lane = threadIdx.x & 31;
  float ag[WG ? NV * 4 : 1], ab[WG ? NV * 4 : 1];
  float gm[NV * 4];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 t = __ldg((const float4*)(gamma + (i * 32 + lane) * 4));
    gm[i * 4] = t.x; gm[i * 4 + 1] = t.y; gm[i * 4 + 2] = t.z; gm[i * 4 + 3] = t.w;
  }
  if (WG) {
#pragma unroll
    for (int i = 0; i < NV * 4; ++i) ag[i] = ab[i] = 0.f;
    for (int i = threadIdx.x; i < c; i += blockDim.x) sg[i] = sb[i] = 0.f;
    __syncthreads();
  }
  for (int64_t row = blockIdx.x * (int64_t)kWarpsPerBlock + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * kWarpsPerBlock) {
    const float mu = mean[row], rs = rstd[row];
    float xh[NV * 4], g[NV * 4], rsd[NV * 4];
    float s1 = 0.f, s2 = 0.f;
    // all of the row's global loads are issued before the first reduction (memory-level parallelism, not occupancy, feeds HBM here)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 4;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dres1) t = *(const float4*)(dres1 + row * (int64_t)c + col);
      if (dres2) {
        const float4 u = *(const float4*)(dres2 + row * (int64_t)c + col);
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      rsd[i * 4] = t.x; rsd[i * 4 + 1] = t.y; rsd[i * 4 + 2] = t.z; rsd[i * 4 + 3] = t.w;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 4;
      const float4 xv = *(const float4*)(x + row * ldx + col);
      float d[4];
      load4_any(dy, dy_dtype, row * lddy + col, lddy / 2, d);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float h = (xs[j] - mu) * rs;
        xh[i * 4 + j] = h;
        if (WG) { ag[i * 4 + j] += d[j] * h; ab[i * 4 + j] += d[j]; }
        const float gg = d[j] * gm[i * 4 + j];
        g[i * 4 + j] = gg;
        s1 += gg;
        s2 += gg * h;
      }
    }
    s1 = warp_sum(s1) * (1.f / c);
    s2 = warp_sum(s2) * (1.f / c);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 4;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = rs * (g[i * 4 + j] - s1 - xh[i * 4 + j] * s2) + rsd[i * 4 + j];
      if (dx) *(float4*)(dx + row * (int64_t)c + col) = make_float4(o[0], o[1], o[2], o[3]);
      if (dx_act) store4_act(dx_act, act_dtype, row * ld_act + col, ld_act / 2, o);
    }
  }
  if (WG) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&sg[col + j], ag[i * 4 + j]);
        atomicAdd(&sb[col + j], ab[i * 4 + j]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
      atomicAdd(dgamma + i, sg[i]);
      atomicAdd(dbeta + i, sb[i]);
    }
  }
}

// Lean variant for the common case (no weight gradients -- the encoder's LayerNorms are frozen, vlm.py:80-88 -- dy in f32 or bf16,
// bf16 / f32 / no operand copy): the generic kernel above keeps xhat, g, the residual and gamma of the whole row in registers
// (128 regs, 13 warps per SM, 86 % of issue slots without an eligible warp, 2.7 TB/s: profiles/r01_ncu_targets.md).  Here a warp
// keeps only the raw x, dy and residual vectors, recomputes xhat after the reduction, reads gamma from shared memory and runs a
// persistent grid-stride loop over rows, so ~24 warps per SM each have their whole row (and residual) in flight.
template <int NV, bool DY_BF16>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 3)
layernorm_bwd_lean_kernel(const void* __restrict__ dy, int64_t lddy, const float* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                          const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dres1,
                          const float* __restrict__ dres2, float* __restrict__ dx, void* __restrict__ dx_act, int act_dtype, int64_t ld_act,
                          int64_t rows) {
  constexpr int c = NV * 128;
  __shared__ float4 sgm[NV * 32];
  for (int i = threadIdx.x; i < NV * 32; i += blockDim.x) sgm[i] = __ldg((const float4*)gamma + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int64_t row = blockIdx.x * (int64_t)kWarpsPerBlock + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * kWarpsPerBlock) {
    float4 xv[NV];
    uint2 db[DY_BF16 ? NV : 1];
    float4 df[DY_BF16 ? 1 : NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 4;
      xv[i] = *(const float4*)(x + row * ldx + col);
      if (DY_BF16) db[i] = *(const uint2*)((const __nv_bfloat16*)dy + row * lddy + col);
      else df[i] = *(const float4*)((const float*)dy + row * lddy + col);
    }
    // the residual gradients are only needed after the reduction: pull their lines into L2 now (no registers), load them at use
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (dres1) asm volatile("prefetch.global.L2 [%0];" ::"l"(dres1 + row * (int64_t)c + (i * 32 + lane) * 4));
      if (dres2) asm volatile("prefetch.global.L2 [%0];" ::"l"(dres2 + row * (int64_t)c + (i * 32 + lane) * 4));
    }
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 gm = sgm[i * 32 + lane];
      float d[4];
      if (DY_BF16) {
        const float2 a = __bfloat1622float2(*(const __nv_bfloat162*)&db[i].x), b = __bfloat1622float2(*(const __nv_bfloat162*)&db[i].y);
        d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y;
      } else {
        d[0] = df[i].x; d[1] = df[i].y; d[2] = df[i].z; d[3] = df[i].w;
      }
      const float g0 = d[0] * gm.x, g1 = d[1] * gm.y, g2 = d[2] * gm.z, g3 = d[3] * gm.w;
      s1 += (g0 + g1) + (g2 + g3);
      s2 += g0 * ((xv[i].x - mu) * rs) + g1 * ((xv[i].y - mu) * rs) + g2 * ((xv[i].z - mu) * rs) + g3 * ((xv[i].w - mu) * rs);
    }
    s1 = warp_sum(s1) * (1.f / c);
    s2 = warp_sum(s2) * (1.f / c);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 4;
      const float4 gm = sgm[i * 32 + lane];
      float d[4];
      if (DY_BF16) {
        const float2 a = __bfloat1622float2(*(const __nv_bfloat162*)&db[i].x), b = __bfloat1622float2(*(const __nv_bfloat162*)&db[i].y);
        d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y;
      } else {
        d[0] = df[i].x; d[1] = df[i].y; d[2] = df[i].z; d[3] = df[i].w;
      }
      float o[4];
      o[0] = rs * (d[0] * gm.x - s1 - (xv[i].x - mu) * rs * s2);
      o[1] = rs * (d[1] * gm.y - s1 - (xv[i].y - mu) * rs * s2);
      o[2] = rs * (d[2] * gm.z - s1 - (xv[i].z - mu) * rs * s2);
      o[3] = rs * (d[3] * gm.w - s1 - (xv[i].w - mu) * rs * s2);
      if (dres1) {
        const float4 u = *(const float4*)(dres1 + row * (int64_t)c + col);
        o[0] += u.x; o[1] += u.y; o[2] += u.z; o[3] += u.w;
      }
      if (dres2) {
        const float4 u = *(const float4*)(dres2 + row * (int64_t)c + col);
        o[0] += u.x; o[1] += u.y; o[2] += u.z; o[3] += u.w;
      }
      if (dx) *(float4*)(dx + row * (int64_t)c + col) = make_float4(o[0], o[1], o[2], o[3]);
      if (dx_act) store4_act(dx_act, act_dtype, row * ld_act + col, ld_act / 2, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------- L2 normalisation
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l2norm_fwd_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ y, void* __restrict__ y_act, int act_dtype, int64_t ld_act,
                  float* __restrict__ inv_norm, int64_t rows, int c, float eps) {
  const int lane = threadIdx.x & 31;
  for (int64_t row = blockIdx.x * (int64_t)kWarpsPerBlock + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * kWarpsPerBlock) {
    const float* xr = x + row * ldx;
    float s = 0.f;
    for (int i = lane; i < c; i += 32) s += xr[i] * xr[i];
    const float nrm = sqrtf(warp_sum(s));
    const float inv = 1.f / fmaxf(nrm, eps);
    if (lane == 0 && inv_norm) inv_norm[row] = inv;
    for (int i = lane; i < c; i += 32) {
      const float v = xr[i] * inv;
      if (y) y[row * (int64_t)c + i] = v;
      if (y_act) store_from_f32(y_act, act_dtype, row * ld_act + i, v, ld_act / 2);
    }
  }
}

// y = x * inv  ->  dx = inv * (dy - y * <dy, y>)      (rows whose norm was clamped by eps are treated the same way; never hit in practice)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l2norm_bwd_kernel(const void* __restrict__ dy, int dy_dtype, int64_t lddy, const float* __restrict__ y, const float* __restrict__ inv_norm,
                  float* __restrict__ dx, int64_t lddx, int accumulate, int64_t rows, int c) {
  const int lane = threadIdx.x & 31;
  for (int64_t row = blockIdx.x * (int64_t)kWarpsPerBlock + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * kWarpsPerBlock) {
    float s = 0.f;
    for (int i = lane; i < c; i += 32) s += load_as_f32(dy, dy_dtype, row * lddy + i, lddy / 2) * y[row * (int64_t)c + i];
    s = warp_sum(s);
    const float inv = inv_norm[row];
    for (int i = lane; i < c; i += 32) {
      float v = inv * (load_as_f32(dy, dy_dtype, row * lddy + i, lddy / 2) - y[row * (int64_t)c + i] * s);
      if (accumulate) v += dx[row * lddx + i];
      dx[row * lddx + i] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------- utilities
__global__ void cast_kernel(const void* __restrict__ src, int src_dtype, int64_t ld_src, int64_t sbs, void* __restrict__ dst, int dst_dtype,
                            int64_t ld_dst, int64_t dbs, int64_t rows_per_batch, int64_t rows, int cols, float scale) {
  const int groups = (cols + 7) / 8;
  const int64_t total = rows * groups;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(idx % groups);
    const int64_t row = idx / groups;
    const int64_t bi = row / rows_per_batch, ri = row % rows_per_batch;
    const int cnt = min(8, cols - g * 8);
    float f[8];
    ld8(src, src_dtype, bi * sbs + ri * ld_src + g * 8, ld_src / 2, cnt, f);
    if (scale != 1.f) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] *= scale;
    }
    st8(dst, dst_dtype, bi * dbs + ri * ld_dst + g * 8, ld_dst / 2, cnt, f);
  }
}

// out[col] += sum_rows x[row, col]; a thread owns 8 consecutive columns (one 16-byte load per row), blockDim = (32 column groups, 8 row lanes),
// grid = (column tiles of 256, row splits); block partials through shared memory, one atomic per column per block
__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ x, int x_dtype, int64_t ld, int64_t rows, int cols, float* __restrict__ out) {
  __shared__ float red[8][257];
  const int col = (blockIdx.x * 32 + threadIdx.x) * 8;
  const int cnt = min(8, cols - col);
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  if (cnt > 0) {
    for (int64_t r = blockIdx.y * 8 + threadIdx.y; r < rows; r += (int64_t)gridDim.y * 8) {
      float f[8];
      ld8(x, x_dtype, r * ld + col, ld / 2, cnt, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += f[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.y][threadIdx.x * 8 + i] = s[i];
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;            // 0..255: one column of the tile each
  const int c = blockIdx.x * 256 + t;
  if (c < cols) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += red[i][t];
    atomicAdd(out + c, v);
  }
}

// Streaming variant (16-byte-aligned rows of bf16 / f32, cols % 8 == 0): the thread block is shaped to the matrix -- vt column
// lanes (8 columns each) x 256 / vt row lanes, so a 32-column gradient still uses all 256 threads -- and every thread keeps four
// rows (64-128 bytes) in flight.  The generic kernel above had 1 load in flight and idled 7/8 of the block on narrow matrices.
template <bool BF16>
__global__ void __launch_bounds__(256)
colsum_stream_kernel(const void* __restrict__ x, int64_t ld, int64_t rows, int cols, float* __restrict__ out, int vt_shift) {
  __shared__ float red[256 * 8 + 256];           // [nrl][vt * 8 + 1], nrl * vt = 256, nrl <= 256
  const int vt = 1 << vt_shift, cl = threadIdx.x & (vt - 1), rl = threadIdx.x >> vt_shift, nrl = 256 >> vt_shift;
  const int col = (blockIdx.x * vt + cl) * 8;
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  if (col < cols) {
    const int64_t stride = (int64_t)gridDim.y * nrl;
    for (int64_t r = (int64_t)blockIdx.y * nrl + rl; r < rows; r += 4 * stride) {
      float f[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t rr = r + u * stride < rows ? r + u * stride : r;
        if (BF16) {
          const uint4 v = __ldg((const uint4*)((const __nv_bfloat16*)x + rr * ld + col));
          bf16x8_to_f32(v, f[u]);
        } else {
          const float4 a = __ldg((const float4*)((const float*)x + rr * ld + col)), b = __ldg((const float4*)((const float*)x + rr * ld + col) + 1);
          f[u][0] = a.x; f[u][1] = a.y; f[u][2] = a.z; f[u][3] = a.w; f[u][4] = b.x; f[u][5] = b.y; f[u][6] = b.z; f[u][7] = b.w;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (r + u * stride < rows) {
#pragma unroll
          for (int i = 0; i < 8; ++i) s[i] += f[u][i];
        }
      }
    }
  }
  // red[rl][cl * 8 + i], row pitch vt * 8 + 1
  const int pitch = vt * 8 + 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) red[rl * pitch + cl * 8 + i] = s[i];
  __syncthreads();
  if (threadIdx.x < vt * 8) {
    const int c = blockIdx.x * vt * 8 + threadIdx.x;
    if (c < cols) {
      float v = 0.f;
      for (int j = 0; j < nrl; ++j) v += red[j * pitch + threadIdx.x];
      atomicAdd(out + c, v);
    }
  }
}

// out[i] (+)= sum_b x[b, i]
__global__ void batch_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int b, int64_t inner, int accumulate) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < inner; i += (int64_t)gridDim.x * blockDim.x) {
    float s = accumulate ? out[i] : 0.f;
    for (int j = 0; j < b; ++j) s += x[j * inner + i];
    out[i] = s;
  }
}

__global__ void axpy_kernel(float* __restrict__ dst, const float* __restrict__ src, float alpha, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] += alpha * src[i];
}

inline int ew_grid(int64_t total, int block = 256) {
  int64_t b = cdiv(total, block);
  int64_t cap = 148 * 32;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace
}  // namespace svl

using namespace svl;

extern "C" int svl_patchify(const float* img, void* out, int out_dtype, int b, int H, int W, int p, int hp, int wp, void* stream) {
  SVL_CHECK_ARG(img && out && b > 0 && p % 8 == 0, "svl_patchify: bad arguments");
  SVL_CHECK_ARG(hp * p >= H && wp * p >= W, "svl_patchify: patch grid does not cover the image");
  const int64_t total = (int64_t)b * hp * wp * (3 * p * p / 8);
  patchify_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(img, out, out_dtype, b, H, W, p, hp, wp);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_assemble_tokens(const float* patches, const float* cls, const float* pos, float* x, int b, int hw, int c, void* stream) {
  SVL_CHECK_ARG(patches && cls && pos && x && c % 4 == 0, "svl_assemble_tokens: bad arguments");
  const int64_t total = (int64_t)b * (hw + 1) * (c / 4);
  assemble_tokens_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(patches, cls, pos, x, b, hw, c);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_pos_resize_fwd(const float* pos, float* out, int gh, int gw, int oh, int ow, int c, void* stream) {
  SVL_CHECK_ARG(pos && out && c % 4 == 0 && gh > 0 && gw > 0 && oh > 0 && ow > 0, "svl_pos_resize_fwd: bad arguments");
  pos_resize_kernel<false><<<ew_grid((int64_t)(oh * ow + 1) * (c / 4)), 256, 0, (cudaStream_t)stream>>>(pos, out, gh, gw, oh, ow, c);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_pos_resize_bwd(const float* dout, float* dpos, int gh, int gw, int oh, int ow, int c, void* stream) {
  SVL_CHECK_ARG(dout && dpos && c % 4 == 0 && gh > 0 && gw > 0 && oh > 0 && ow > 0, "svl_pos_resize_bwd: bad arguments");
  pos_resize_kernel<true><<<ew_grid((int64_t)(oh * ow + 1) * (c / 4)), 256, 0, (cudaStream_t)stream>>>(dout, dpos, gh, gw, oh, ow, c);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, void* y, int y_dtype, int64_t ldy,
                                 float* mean, float* rstd, int64_t rows, int c, float eps, void* stream) {
  SVL_CHECK_ARG(x && gamma && beta && y, "svl_layernorm_fwd: null pointer");
  SVL_CHECK_ARG(c % 128 == 0 && c <= 128 * kMaxVec, "svl_layernorm_fwd: c=%d must be a multiple of 128 and <= %d", c, 128 * kMaxVec);
  SVL_CHECK_ARG(ldx % 4 == 0 && ldy % 4 == 0, "svl_layernorm_fwd: strides must be multiples of 4");
  if (rows == 0) return SVL_OK;
  layernorm_fwd_kernel<<<grid_for_rows(rows), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(x, ldx, gamma, beta, y, y_dtype, ldy, mean, rstd,
                                                                                           rows, c, eps);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_layernorm_bwd(const void* dy, int dy_dtype, int64_t lddy, const float* x, int64_t ldx, const float* gamma,
                                 const float* mean, const float* rstd, const float* dres1, const float* dres2, float* dx, void* dx_act,
                                 int act_dtype, int64_t ld_act, float* dgamma, float* dbeta, int64_t rows, int c, void* stream) {
  SVL_CHECK_ARG(dy && x && gamma && mean && rstd && (dx || dx_act), "svl_layernorm_bwd: null pointer");
  SVL_CHECK_ARG(c % 128 == 0 && c <= 128 * kMaxVec, "svl_layernorm_bwd: c=%d must be a multiple of 128 and <= %d", c, 128 * kMaxVec);
  if (rows == 0) return SVL_OK;
  int grid = grid_for_rows(rows);
  if (dgamma && grid > 296) grid = 296;     // fewer, longer-lived blocks: cheaper dgamma/dbeta reduction
  SVL_CHECK_ARG(c == 768 || c == 256, "svl_layernorm_bwd: rows of %d elements are not instantiated (768: ViT, 256: class attention)", c);
  SVL_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr), "svl_layernorm_bwd: dgamma and dbeta go together");
  if (!dgamma && c == 768 && (dy_dtype == SVL_F32 || dy_dtype == SVL_BF16) && lddy % 4 == 0 && ldx % 4 == 0 && ld_act % 4 == 0) {
    const int lean_grid = (int)(cdiv(rows, kWarpsPerBlock) < 148 * 3 ? cdiv(rows, kWarpsPerBlock) : 148 * 3);     // persistent: 3 CTAs per SM
    if (dy_dtype == SVL_BF16)
      layernorm_bwd_lean_kernel<6, true><<<lean_grid, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(dy, lddy, x, ldx, gamma, mean, rstd, dres1, dres2, dx,
                                                                                                   dx_act, act_dtype, ld_act, rows);
    else
      layernorm_bwd_lean_kernel<6, false><<<lean_grid, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(dy, lddy, x, ldx, gamma, mean, rstd, dres1, dres2, dx,
                                                                                                    dx_act, act_dtype, ld_act, rows);
    SVL_LAUNCH_CHECK();
    return SVL_OK;
  }
#define SVL_LN_BWD(NV, WG)                                                                                                                       \
  layernorm_bwd_kernel<NV, WG><<<grid, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(dy, dy_dtype, lddy, x, ldx, gamma, mean, rstd, dres1, dres2, \
                                                                                        dx, dx_act, act_dtype, ld_act, dgamma, dbeta, rows)
  if (c == 768) { if (dgamma) SVL_LN_BWD(6, true); else SVL_LN_BWD(6, false); }
  else { if (dgamma) SVL_LN_BWD(2, true); else SVL_LN_BWD(2, false); }
#undef SVL_LN_BWD
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_l2norm_fwd(const float* x, int64_t ldx, float* y, void* y_act, int act_dtype, int64_t ld_act, float* inv_norm, int64_t rows,
                              int c, float eps, void* stream) {
  SVL_CHECK_ARG(x && (y || y_act), "svl_l2norm_fwd: null pointer");
  if (rows == 0) return SVL_OK;
  l2norm_fwd_kernel<<<grid_for_rows(rows), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(x, ldx, y, y_act, act_dtype, ld_act, inv_norm, rows, c, eps);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_l2norm_bwd(const void* dy, int dy_dtype, int64_t lddy, const float* y, const float* inv_norm, float* dx, int64_t lddx,
                              int accumulate, int64_t rows, int c, void* stream) {
  SVL_CHECK_ARG(dy && y && inv_norm && dx, "svl_l2norm_bwd: null pointer");
  if (rows == 0) return SVL_OK;
  l2norm_bwd_kernel<<<grid_for_rows(rows), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(dy, dy_dtype, lddy, y, inv_norm, dx, lddx, accumulate,
                                                                                        rows, c);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

// Batched parameter jobs: ONE launch refreshes every GEMM-operand copy of the trainable weights after the optimizer step (dst[i] =
// cast(src[idx[i]]), idx < 0 = zero padding: any operand layout is a gather of the parameter) or scatters every staged weight gradient
// back into the parameter layout (dst[idx[i]] += src[i]; src[i] = 0: the index maps are injective, no atomics).  A job row is
// {src, dst, idx, n, flags} (flags bit 0: dst is f32 instead of bf16; bit 1: the layout is the 2-D transpose of the parameter viewed as
// [flags >> 8, n / (flags >> 8)], handled in 32 x 32 tiles); block_start[j] = first block of job j (1024 elements or one tile each).
__global__ void __launch_bounds__(256) param_jobs_kernel(const long long* __restrict__ jobs, const int* __restrict__ block_start, int njobs, int mode) {
  int lo = 0, hi = njobs;                                   // last job whose first block is <= blockIdx.x
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (block_start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
  }
  const long long* J = jobs + 5 * lo;
  const long long n = J[3];
  const int* __restrict__ idx = (const int*)J[2];
  const long long base = (long long)((int)blockIdx.x - block_start[lo]) * 1024;
  if (mode == 0 && (J[4] & 2)) {
    // the operand is the plain 2-D transpose of the parameter viewed as [R, C] (data-gradient copies of the linear layers: half of all
    // refreshed bytes): 32 x 32 tiles through shared memory, both sides coalesced, instead of a gather with a stride of C elements
    __shared__ float tile[32][33];
    const int R = (int)(J[4] >> 8), C = (int)(n / R);
    const int tiles_c = (C + 31) / 32;
    const int b = (int)blockIdx.x - block_start[lo], tr = b / tiles_c, tc = b - tr * tiles_c;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* __restrict__ src = (const float*)J[0];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = tr * 32 + ty + 8 * k, c = tc * 32 + tx;
      tile[ty + 8 * k][tx] = (r < R && c < C) ? src[(long long)r * C + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = tc * 32 + ty + 8 * k, r = tr * 32 + tx;      // dst is [C, R]
      if (c < C && r < R) {
        const float v = tile[tx][ty + 8 * k];
        if (J[4] & 1) ((float*)J[1])[(long long)c * R + r] = v;
        else ((__nv_bfloat16*)J[1])[(long long)c * R + r] = __float2bfloat16(v);
      }
    }
    return;
  }
  if (mode == 0) {
    const float* __restrict__ src = (const float*)J[0];
    const bool f32 = J[4] & 1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = base + k * 256 + threadIdx.x;
      if (i < n) {
        const int s = idx[i];
        const float v = s >= 0 ? src[s] : 0.f;
        if (f32) ((float*)J[1])[i] = v;
        else ((__nv_bfloat16*)J[1])[i] = __float2bfloat16(v);
      }
    }
  } else {
    float* __restrict__ src = (float*)J[0];
    float* __restrict__ dst = (float*)J[1];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = base + k * 256 + threadIdx.x;
      if (i < n) {
        const int s = idx[i];
        if (s >= 0) dst[s] += src[i];
        src[i] = 0.f;
      }
    }
  }
}

extern "C" int svl_param_jobs(const void* jobs, const void* block_start, int njobs, int total_blocks, int mode, void* stream) {
  SVL_CHECK_ARG(jobs && block_start && njobs >= 1 && total_blocks >= 1 && (mode == 0 || mode == 1), "svl_param_jobs: bad arguments");
  param_jobs_kernel<<<total_blocks, 256, 0, (cudaStream_t)stream>>>((const long long*)jobs, (const int*)block_start, njobs, mode);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_cast(const void* src, int src_dtype, int64_t ld_src, int64_t src_batch_stride, void* dst, int dst_dtype, int64_t ld_dst,
                        int64_t dst_batch_stride, int batch, int64_t rows, int cols, float scale, void* stream) {
  SVL_CHECK_ARG(src && dst && batch >= 1, "svl_cast: bad arguments");
  if (rows == 0 || cols == 0) return SVL_OK;
  cast_kernel<<<ew_grid(batch * rows * ((cols + 7) / 8)), 256, 0, (cudaStream_t)stream>>>(
      src, src_dtype, ld_src, src_batch_stride, dst, dst_dtype, ld_dst, dst_batch_stride, rows, batch * rows, cols, scale == 0.f ? 1.f : scale);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_colsum(const void* x, int x_dtype, int64_t ld, int64_t rows, int cols, float* out, void* stream) {
  SVL_CHECK_ARG(x && out, "svl_colsum: null pointer");
  if (rows == 0 || cols == 0) return SVL_OK;
  if ((x_dtype == SVL_F32 || x_dtype == SVL_BF16) && cols % 8 == 0 && ld % 8 == 0 && ((uintptr_t)x & 15) == 0) {
    int vt_shift = 0;
    while ((1 << vt_shift) < cols / 8 && vt_shift < 5) ++vt_shift;            // column lanes per block: next power of two of cols / 8, <= 32
    const int vt = 1 << vt_shift, nrl = 256 >> vt_shift;
    const int xt = (int)cdiv(cols / 8, vt);
    int64_t ysplit = cdiv(rows, (int64_t)nrl * 8);                            // >= 8 rows per thread
    const int64_t cap = (148 * 8 + xt - 1) / xt;
    if (ysplit > cap) ysplit = cap;
    if (ysplit < 1) ysplit = 1;
    dim3 grid(xt, (unsigned)ysplit);
    if (x_dtype == SVL_BF16) colsum_stream_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, ld, rows, cols, out, vt_shift);
    else colsum_stream_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, ld, rows, cols, out, vt_shift);
    SVL_LAUNCH_CHECK();
    return SVL_OK;
  }
  const int xt = (cols + 255) / 256;
  int64_t ysplit = cdiv(rows, 8 * 16);
  const int64_t cap = (148 * 8 + xt - 1) / xt;
  if (ysplit > cap) ysplit = cap;
  dim3 grid(xt, (unsigned)ysplit);
  colsum_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, x_dtype, ld, rows, cols, out);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_batch_sum(const float* x, float* out, int b, int64_t inner, int accumulate, void* stream) {
  SVL_CHECK_ARG(x && out, "svl_batch_sum: null pointer");
  if (inner == 0) return SVL_OK;
  batch_sum_kernel<<<ew_grid(inner), 256, 0, (cudaStream_t)stream>>>(x, out, b, inner, accumulate);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_axpy(float* dst, const float* src, float alpha, int64_t n, void* stream) {
  SVL_CHECK_ARG(dst && src, "svl_axpy: null pointer");
  if (n == 0) return SVL_OK;
  axpy_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(dst, src, alpha, n);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
