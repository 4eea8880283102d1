// HBM-bound kernels of the VLG decode head (model/decode_heads/vlg_head.py): GroupNorm+ReLU fwd/bwd, the 7x7 im2col of the
// similarity maps and its transpose, global-average pooling, the class-token pooling / un-pooling around the
// SemanticTransformer, the skip-feature upsample + concat and its gradient, and the 32->1 output conv.
// Activations are NHWC; a "map" is one (image, class) pair, maps are ordered (image, class).
#include <stdlib.h>

#include "common.cuh"

namespace svl {
namespace {

inline int ew_grid(int64_t total, int block = 256) {
  int64_t b = cdiv(total, block);
  int64_t cap = 148 * 32;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// ---------------------------------------------------------------------------------------------- GroupNorm + ReLU
// Three phases, each a full grid of (map, split) CTAs so that small batches of large maps still fill the machine:
//   stats    : per-(map, group) sum / sum of squares (block partials -> global atomics)
//   finalize : sums -> mean / rstd (in place)
//   apply    : y = relu((x - mean) * rstd * gamma + beta) (+ residual); a thread keeps the same 8 channels for all its pixels
constexpr int kGnThreads = 256;
constexpr int kMaxGroups = 16;

__device__ __forceinline__ void gn_range(int hw, int splits, int split, int& p0, int& p1) {
  const int per = (hw + splits - 1) / splits;
  p0 = split * per;
  p1 = p0 + per < hw ? p0 + per : hw;
}

// Fixed-order reduction of the per-thread (a, b) partials of a statistics CTA to per-group sums.  Thread t owns channel vector
// t % vpp (8 channels, two vectors per 16-channel group) of pixel slot t / vpp.  Lanes of one group inside a warp are combined
// by an xor-shuffle tree (bit 0 = the vector pair, bits >= log2(vpp) = pixel slots), the warps by one thread per group in
// warp order: the result does not depend on scheduling, so the statistics -- and with them every bf16 rounding downstream --
// are bit-reproducible from run to run (no fp32 atomics; profiles/r01_determinism.md).
__device__ __forceinline__ void gn_block_reduce2(float a, float b, int vpp, int G, float* s_part /* [2][8][kMaxGroups] */, float& ra, float& rb,
                                                 bool pair = true) {
  // pair == false: thread t owns all 16 channels of group t % vpp (vpp = G lanes per pixel slot), no vector pair to combine
  if (pair) {
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    b += __shfl_xor_sync(0xffffffffu, b, 1);
  }
  for (int o = vpp; o < 32; o <<= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // vpp <= 32: a warp covers groups [0, vpp/2) (vpp == 32: all 16) -- with vpp > 32 excluded by the C <= 256 check
  if (lane < vpp && (!pair || (lane & 1) == 0)) {
    const int gi = pair ? lane >> 1 : lane;
    s_part[(0 * 8 + warp) * kMaxGroups + gi] = a;
    s_part[(1 * 8 + warp) * kMaxGroups + gi] = b;
  }
  __syncthreads();
  ra = rb = 0.f;
  if (threadIdx.x < G) {
#pragma unroll
    for (int w = 0; w < kGnThreads / 32; ++w) {
      ra += s_part[(0 * 8 + w) * kMaxGroups + threadIdx.x];
      rb += s_part[(1 * 8 + w) * kMaxGroups + threadIdx.x];
    }
  }
}

__global__ void __launch_bounds__(kGnThreads)
gn_stats_kernel(const void* __restrict__ x, int x_dtype, int64_t ldx, float* __restrict__ part, int hw, int C, int G, int splits) {
  __shared__ float s_part[2 * 8 * kMaxGroups];
  const int map = blockIdx.x / splits, split = blockIdx.x % splits;
  const int vpp = C / 8, ppi = kGnThreads / vpp;      // pixels per iteration
  int p0, p1;
  gn_range(hw, splits, split, p0, p1);
  const int c8 = (threadIdx.x % vpp) * 8;
  const int64_t xbase = (int64_t)map * hw * ldx + c8;
  float pa = 0.f, pb = 0.f;
  for (int pix = p0 + threadIdx.x / vpp; pix < p1; pix += 4 * ppi) {
    float f[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (pix + u * ppi < p1) ld8(x, x_dtype, xbase + (int64_t)(pix + u * ppi) * ldx, ldx / 2, 8, f[u]);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[u][i] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { pa += f[u][i]; pb += f[u][i] * f[u][i]; }
    }
  }
  float ra, rb;
  gn_block_reduce2(pa, pb, vpp, G, s_part, ra, rb);
  if (threadIdx.x < G) {                                 // partial of this (map, split): reduced in split order by gn_finalize_kernel
    part[((int64_t)blockIdx.x * G + threadIdx.x) * 2] = ra;
    part[((int64_t)blockIdx.x * G + threadIdx.x) * 2 + 1] = rb;
  }
}
// second stage: the `splits` partials of a (map, group) are summed in split order
__global__ void gn_finalize_kernel(const float* __restrict__ part, float* __restrict__ mean, float* __restrict__ rstd, int64_t maps, int G, int splits,
                                   float count, float eps) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= maps * G) return;
  const int64_t map = i / G;
  const int g = (int)(i % G);
  float a = 0.f, b = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float2 v = *(const float2*)(part + (((map * splits + s) * G) + g) * 2);
    a += v.x;
    b += v.y;
  }
  const float mu = a / count;
  const float var = fmaxf(b / count - mu * mu, 0.f);
  mean[i] = mu;
  rstd[i] = rsqrtf(var + eps);
}
__global__ void gn_bwd_finalize_kernel(const float* __restrict__ part, float* __restrict__ ws, int64_t maps, int G, int splits) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= maps * G) return;
  const int64_t map = i / G;
  const int g = (int)(i % G);
  float a = 0.f, b = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float2 v = *(const float2*)(part + (((map * splits + s) * G) + g) * 2);
    a += v.x;
    b += v.y;
  }
  ws[i * 2] = a;
  ws[i * 2 + 1] = b;
}
// apply: one thread = the 16 channels of one (pixel, group) -- every GroupNorm of the head has 16 channels per group -- i.e. two
// 16-byte vectors in flight per thread at ~40 registers, flat grid-stride indexing for full occupancy.
__global__ void __launch_bounds__(256, 6)
gn_apply_kernel(const void* __restrict__ x, int x_dtype, int64_t ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                void* __restrict__ out, int out_dtype, int64_t ldo, const void* __restrict__ res, int res_dtype, int64_t ldres,
                const float* __restrict__ mean, const float* __restrict__ rstd, int64_t maps, int hw, int C, int G) {
  const int64_t total = maps * hw * G;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(idx % G);
    const int64_t pix = idx / G;                 // global pixel index (map * hw + p)
    const int64_t map = pix / hw;
    const int c0 = g * 16;
    float f[16];
    ld8(x, x_dtype, pix * ldx + c0, ldx / 2, 8, f);
    ld8(x, x_dtype, pix * ldx + c0 + 8, ldx / 2, 8, f + 8);
    const float mu = __ldg(mean + map * G + g), rs = __ldg(rstd + map * G + g);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 ga = __ldg((const float4*)(gamma + c0) + j), be = __ldg((const float4*)(beta + c0) + j);
      f[4 * j] = fmaxf((f[4 * j] - mu) * rs * ga.x + be.x, 0.f);
      f[4 * j + 1] = fmaxf((f[4 * j + 1] - mu) * rs * ga.y + be.y, 0.f);
      f[4 * j + 2] = fmaxf((f[4 * j + 2] - mu) * rs * ga.z + be.z, 0.f);
      f[4 * j + 3] = fmaxf((f[4 * j + 3] - mu) * rs * ga.w + be.w, 0.f);
    }
    if (res) {
      float r[16];
      ld8(res, res_dtype, pix * ldres + c0, ldres / 2, 8, r);
      ld8(res, res_dtype, pix * ldres + c0 + 8, ldres / 2, 8, r + 8);
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] += r[i];
    }
    st8(out, out_dtype, pix * ldo + c0, ldo / 2, 8, f);
    st8(out, out_dtype, pix * ldo + c0 + 8, ldo / 2, 8, f + 8);
  }
}

// backward: stats (s1 = sum g, s2 = sum g*xhat per (map, group), dgamma/dbeta) then apply  dx = rstd * (g - s1/n - xhat * s2/n).
// One thread = the 16 channels of one (pixel, group): four 16-byte loads (x, dy) in flight per thread, raw bf16 kept packed.
__global__ void __launch_bounds__(kGnThreads, 3)
gn_bwd_stats_kernel(const void* __restrict__ dy, int dy_dtype, int64_t lddy, const void* __restrict__ x, int x_dtype, int64_t ldx,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                    const float* __restrict__ rstd, float* __restrict__ ws, float* __restrict__ dgamma, float* __restrict__ dbeta, int hw,
                    int C, int G, int splits) {
  __shared__ float s_part[2 * 8 * kMaxGroups];
  __shared__ float s_dg[256], s_db[256];
  const int map = blockIdx.x / splits, split = blockIdx.x % splits;
  const int ppi = kGnThreads / G;                     // pixels per iteration
  s_dg[threadIdx.x] = s_db[threadIdx.x] = 0.f;
  __syncthreads();
  int p0, p1;
  gn_range(hw, splits, split, p0, p1);
  const int g = threadIdx.x % G, c0 = g * 16;
  const float mu = mean[(int64_t)map * G + g], rs = rstd[(int64_t)map * G + g];
  const int64_t xbase = (int64_t)map * hw * ldx + c0, ybase = (int64_t)map * hw * lddy + c0;
  float s1 = 0.f, s2 = 0.f, dg[16], db[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) dg[i] = db[i] = 0.f;
#pragma unroll 1
  for (int pix = p0 + threadIdx.x / G; pix < p1; pix += ppi) {
    float xv[16], d[16];
    ld8(x, x_dtype, xbase + (int64_t)pix * ldx, ldx / 2, 8, xv);
    ld8(x, x_dtype, xbase + (int64_t)pix * ldx + 8, ldx / 2, 8, xv + 8);
    ld8(dy, dy_dtype, ybase + (int64_t)pix * lddy, lddy / 2, 8, d);
    ld8(dy, dy_dtype, ybase + (int64_t)pix * lddy + 8, lddy / 2, 8, d + 8);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float ga = __ldg(gamma + c0 + i), be = __ldg(beta + c0 + i);
      const float xh = (xv[i] - mu) * rs;
      const float dd = (xh * ga + be > 0.f) ? d[i] : 0.f;
      dg[i] += dd * xh;
      db[i] += dd;
      const float gg = dd * ga;
      s1 += gg;
      s2 += gg * xh;
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) { atomicAdd(&s_dg[c0 + i], dg[i]); atomicAdd(&s_db[c0 + i], db[i]); }
  float ra, rb;
  gn_block_reduce2(s1, s2, G, G, s_part, ra, rb, false);           // (contains the __syncthreads that also publishes s_dg / s_db)
  if (threadIdx.x < G) {                                           // partial of this (map, split), reduced in split order by gn_bwd_finalize_kernel
    ws[((int64_t)blockIdx.x * G + threadIdx.x) * 2] = ra;
    ws[((int64_t)blockIdx.x * G + threadIdx.x) * 2 + 1] = rb;
  }
  if (threadIdx.x < C && dgamma) {
    atomicAdd(dgamma + threadIdx.x, s_dg[threadIdx.x]);
    atomicAdd(dbeta + threadIdx.x, s_db[threadIdx.x]);
  }
}
__global__ void __launch_bounds__(256, 4)
gn_bwd_apply_kernel(const void* __restrict__ dy, int dy_dtype, int64_t lddy, const void* __restrict__ x, int x_dtype, int64_t ldx,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const float* __restrict__ ws, void* __restrict__ dx, int dx_dtype, int64_t lddx, int64_t maps,
                    int hw, int C, int G) {
  const int64_t total = maps * hw * G;
  const float n = (float)hw * 16.f;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(idx % G);
    const int64_t pix = idx / G;
    const int64_t map = pix / hw;
    const int c0 = g * 16;
    float xv[16], d[16];
    ld8(x, x_dtype, pix * ldx + c0, ldx / 2, 8, xv);
    ld8(x, x_dtype, pix * ldx + c0 + 8, ldx / 2, 8, xv + 8);
    ld8(dy, dy_dtype, pix * lddy + c0, lddy / 2, 8, d);
    ld8(dy, dy_dtype, pix * lddy + c0 + 8, lddy / 2, 8, d + 8);
    const float mu = __ldg(mean + map * G + g), rs = __ldg(rstd + map * G + g);
    const float m1 = __ldg(ws + (map * G + g) * 2) / n, m2 = __ldg(ws + (map * G + g) * 2 + 1) / n;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float ga = __ldg(gamma + c0 + i), be = __ldg(beta + c0 + i);
      const float xh = (xv[i] - mu) * rs;
      const float dd = (xh * ga + be > 0.f) ? d[i] : 0.f;
      d[i] = rs * (dd * ga - m1 - xh * m2);
    }
    st8(dx, dx_dtype, pix * lddx + c0, lddx / 2, 8, d);
    st8(dx, dx_dtype, pix * lddx + c0 + 8, lddx / 2, 8, d + 8);
  }
}

// ---------------------------------------------------------------------------------------------- all-bf16 GroupNorm fast paths
// Every large GroupNorm of the throughput mode reads a bf16 conv output and writes a bf16 operand.  The generic kernels above
// go through ld8/st8 (run-time dtype + alignment dispatch around every access), which keeps ptxas from batching the loads:
// they ran at 1.2 - 3.6 TB/s (profiles/r01_ncu_targets.md).  Here a thread owns ONE 8-channel vector (16 bytes) of U pixels and
// issues all its 16-byte loads back to back (streaming, no L1 allocation) before it uses any of them; the map is a grid
// dimension, so there is no per-element division and the statistics of the block are loaded once.
__device__ __forceinline__ uint4 ld_stream16(const __nv_bfloat16* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void ldg8f(const float* p, float* f) {
  const float4 a = __ldg((const float4*)p), b = __ldg((const float4*)p + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

__global__ void __launch_bounds__(kGnThreads)
gn_stats_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, float* __restrict__ part, int hw, int C, int G, int splits) {
  __shared__ float s_part[2 * 8 * kMaxGroups];
  const int map = blockIdx.x / splits, split = blockIdx.x % splits;
  const int vpp = C / 8, ppi = kGnThreads / vpp;
  int p0, p1;
  gn_range(hw, splits, split, p0, p1);
  const int c8 = (threadIdx.x % vpp) * 8;
  const __nv_bfloat16* xb = x + (int64_t)map * hw * ldx + c8;
  float pa = 0.f, pb = 0.f;
  for (int pix = p0 + threadIdx.x / vpp; pix < p1; pix += 4 * ppi) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = pix + u * ppi;
      v[u] = ld_stream16(xb + (int64_t)(q < p1 ? q : pix) * ldx);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (pix + u * ppi < p1) {
        float f[8];
        bf16x8_to_f32(v[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { pa += f[i]; pb += f[i] * f[i]; }
      }
    }
  }
  float ra, rb;
  gn_block_reduce2(pa, pb, vpp, G, s_part, ra, rb);
  if (threadIdx.x < G) {                                 // partial of this (map, split): reduced in split order by gn_finalize_kernel
    part[((int64_t)blockIdx.x * G + threadIdx.x) * 2] = ra;
    part[((int64_t)blockIdx.x * G + threadIdx.x) * 2 + 1] = rb;
  }
}

// grid = (ceil(hw * vpp / (256 * U)), maps); item = (pixel of the map, 8-channel vector); vshift = log2(vpp)
constexpr int kGnApplyU = 4;
__global__ void __launch_bounds__(256)
gn_apply_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                     __nv_bfloat16* __restrict__ out, int64_t ldo, const __nv_bfloat16* __restrict__ res, int64_t ldres,
                     const float* __restrict__ mean, const float* __restrict__ rstd, int hw, int C, int G, int vshift) {
  constexpr int U = kGnApplyU;
  const int vpp = 1 << vshift, v = threadIdx.x & (vpp - 1), c8 = v * 8, g = c8 / (C / G);
  const int map = blockIdx.y, items = hw << vshift;
  const int base = blockIdx.x * (256 * U) + threadIdx.x;
  if (base >= items) return;
  const int64_t pix0 = (int64_t)map * hw;
  int pix[U];
  uint4 xv[U], rv[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int idx = base + u * 256;
    pix[u] = (idx < items ? idx : base) >> vshift;
    xv[u] = ld_stream16(x + (pix0 + pix[u]) * ldx + c8);
  }
  if (res) {
#pragma unroll
    for (int u = 0; u < U; ++u) rv[u] = ld_stream16(res + (pix0 + pix[u]) * ldres + c8);
  }
  float ga[8], be[8];
  ldg8f(gamma + c8, ga);
  ldg8f(beta + c8, be);
  const float mu = __ldg(mean + (int64_t)map * G + g), rs = __ldg(rstd + (int64_t)map * G + g);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (base + u * 256 < items) {
      float f[8];
      bf16x8_to_f32(xv[u], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaxf((f[i] - mu) * rs * ga[i] + be[i], 0.f);       // same expression as the backward's ReLU mask
      if (res) {
        float r[8];
        bf16x8_to_f32(rv[u], r);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] += r[i];
      }
      *(uint4*)(out + (pix0 + pix[u]) * ldo + c8) = f32_to_bf16x8(f);
    }
  }
}

// backward statistics: thread = (pixel slot, 8-channel vector), four pixels (x, dy: eight 16-byte loads) in flight (two left the kernel at
// 3.2 TB/s: latency-bound); the 18 per-thread
// partials (dgamma[8], dbeta[8], s1, s2) are reduced through a padded shared-memory table instead of same-address atomics.
__global__ void __launch_bounds__(kGnThreads, 3)
gn_bwd_stats_bf16_kernel(const __nv_bfloat16* __restrict__ dy, int64_t lddy, const __nv_bfloat16* __restrict__ x, int64_t ldx,
                         const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                         const float* __restrict__ rstd, float* __restrict__ ws, float* __restrict__ dgamma, float* __restrict__ dbeta, int hw,
                         int C, int G, int splits) {
  __shared__ float part[18][kGnThreads + 1];
  const int map = blockIdx.x / splits, split = blockIdx.x % splits;
  const int vpp = C / 8, ppi = kGnThreads / vpp, cpg = C / G;
  int p0, p1;
  gn_range(hw, splits, split, p0, p1);
  const int v = threadIdx.x % vpp, c8 = v * 8, g = c8 / cpg;
  float ga[8], be[8];
  ldg8f(gamma + c8, ga);
  ldg8f(beta + c8, be);
  const float mu = mean[(int64_t)map * G + g], rs = rstd[(int64_t)map * G + g];
  const __nv_bfloat16* xb = x + (int64_t)map * hw * ldx + c8;
  const __nv_bfloat16* yb = dy + (int64_t)map * hw * lddy + c8;
  float s1 = 0.f, s2 = 0.f, dg[8], db[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) dg[i] = db[i] = 0.f;
  constexpr int U = 4;                                   // pixels (x, dy: two 16-byte loads each) in flight per thread: 128 bytes
  for (int pix = p0 + threadIdx.x / vpp; pix < p1; pix += U * ppi) {
    uint4 a[U], b[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      ok[u] = pix + u * ppi < p1;
      const int q = ok[u] ? pix + u * ppi : pix;
      a[u] = ld_stream16(xb + (int64_t)q * ldx);
      b[u] = ld_stream16(yb + (int64_t)q * lddy);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (ok[u]) {
        float xv[8], d[8];
        bf16x8_to_f32(a[u], xv);
        bf16x8_to_f32(b[u], d);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (xv[i] - mu) * rs;
          const float dd = (xh * ga[i] + be[i] > 0.f) ? d[i] : 0.f;
          dg[i] += dd * xh;
          db[i] += dd;
        }
      }
    }
  }
  // s1 = sum dd * gamma and s2 = sum dd * gamma * xhat are linear in the per-channel sums just accumulated: no per-element work
#pragma unroll
  for (int i = 0; i < 8; ++i) { s1 += ga[i] * db[i]; s2 += ga[i] * dg[i]; }
#pragma unroll
  for (int i = 0; i < 8; ++i) { part[i][threadIdx.x] = dg[i]; part[8 + i][threadIdx.x] = db[i]; }
  part[16][threadIdx.x] = s1;
  part[17][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.x < C && dgamma) {                       // channel c = (vector c / 8, lane c % 8): sum over the pixel slots
    const int cv = threadIdx.x / 8, ci = threadIdx.x % 8;
    float a = 0.f, b = 0.f;
    for (int ps = 0; ps < ppi; ++ps) { a += part[ci][ps * vpp + cv]; b += part[8 + ci][ps * vpp + cv]; }
    atomicAdd(dgamma + threadIdx.x, a);
    atomicAdd(dbeta + threadIdx.x, b);
  }
  if (threadIdx.x >= kGnThreads - G) {                   // one (late) thread per group: sum over its vectors and the pixel slots
    const int gg = kGnThreads - 1 - threadIdx.x, v0 = gg * cpg / 8, v1 = (gg + 1) * cpg / 8;
    float a = 0.f, b = 0.f;
    for (int ps = 0; ps < ppi; ++ps)
      for (int vv = v0; vv < v1; ++vv) { a += part[16][ps * vpp + vv]; b += part[17][ps * vpp + vv]; }
    ws[((int64_t)blockIdx.x * G + gg) * 2] = a;               // partial of this (map, split): gn_bwd_finalize_kernel sums them in split order
    ws[((int64_t)blockIdx.x * G + gg) * 2 + 1] = b;
  }
}

constexpr int kGnBwdApplyU = 2;
__global__ void __launch_bounds__(256)
gn_bwd_apply_bf16_kernel(const __nv_bfloat16* __restrict__ dy, int64_t lddy, const __nv_bfloat16* __restrict__ x, int64_t ldx,
                         const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                         const float* __restrict__ rstd, const float* __restrict__ ws, __nv_bfloat16* __restrict__ dx, int64_t lddx, int hw,
                         int C, int G, int vshift) {
  constexpr int U = kGnBwdApplyU;
  const int vpp = 1 << vshift, v = threadIdx.x & (vpp - 1), c8 = v * 8, g = c8 / (C / G);
  const int map = blockIdx.y, items = hw << vshift;
  const int base = blockIdx.x * (256 * U) + threadIdx.x;
  if (base >= items) return;
  const int64_t pix0 = (int64_t)map * hw;
  int pix[U];
  uint4 xv[U], dv[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int idx = base + u * 256;
    pix[u] = (idx < items ? idx : base) >> vshift;
    xv[u] = ld_stream16(x + (pix0 + pix[u]) * ldx + c8);
    dv[u] = ld_stream16(dy + (pix0 + pix[u]) * lddy + c8);
  }
  float ga[8], be[8];
  ldg8f(gamma + c8, ga);
  ldg8f(beta + c8, be);
  const float n = (float)hw * (float)(C / G);
  const float mu = __ldg(mean + (int64_t)map * G + g), rs = __ldg(rstd + (int64_t)map * G + g);
  const float m1 = __ldg(ws + ((int64_t)map * G + g) * 2) / n, m2 = __ldg(ws + ((int64_t)map * G + g) * 2 + 1) / n;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (base + u * 256 < items) {
      float xf[8], d[8];
      bf16x8_to_f32(xv[u], xf);
      bf16x8_to_f32(dv[u], d);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = (xf[i] - mu) * rs;
        const float dd = (xh * ga[i] + be[i] > 0.f) ? d[i] : 0.f;
        d[i] = rs * (dd * ga[i] - m1 - xh * m2);
      }
      *(uint4*)(dx + (pix0 + pix[u]) * lddx + c8) = f32_to_bf16x8(d);
    }
  }
}

inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }
inline int log2_exact(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return (1 << s) == v ? s : -1;
}

// ---------------------------------------------------------------------------------------------- bf16 fast paths of the token / skip glue
// Same idea as the GroupNorm kernels: one 8-channel vector (16 bytes) per thread access instead of a 2-byte element, loads batched.
__device__ __forceinline__ void bilin_src(int o, float scale, int in_size, int& i0, int& i1, float& w1);      // defined with the generic kernels
// out[map, c] = scale * sum_pix x[map, pix, c];  grid = maps, 256 threads = (256 / vpp pixel lanes) x (vpp channel vectors)
__global__ void __launch_bounds__(256)
map_sum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, float* __restrict__ out, int hw, int C, float scale) {
  __shared__ float red[256 * 8 + 256];
  const int vpp = C / 8, lanes = 256 / vpp, v = threadIdx.x % vpp, pl = threadIdx.x / vpp;
  const __nv_bfloat16* xb = x + (int64_t)blockIdx.x * hw * ld + v * 8;
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  for (int p = pl; p < hw; p += 4 * lanes) {
    uint4 r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) r[u] = ld_stream16(xb + (int64_t)(p + u * lanes < hw ? p + u * lanes : p) * ld);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (p + u * lanes < hw) {
        float f[8];
        bf16x8_to_f32(r[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] += f[i];
      }
    }
  }
  const int pitch = C + 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) red[pl * pitch + v * 8 + i] = s[i];
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.f;
    for (int j = 0; j < lanes; ++j) t += red[j * pitch + threadIdx.x];
    out[(int64_t)blockIdx.x * C + threadIdx.x] = t * scale;
  }
}
// tok[(b, py, px), n, 0:C] = avgpool x[(b, n)];  tok[..., C:C+Ct] = text[n]: a thread owns 8 columns of one token
__global__ void __launch_bounds__(256)
pool_tokens_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const float* __restrict__ text, float* __restrict__ tok, int B, int N,
                        int h, int w, int C, int Ct, int pool) {
  const int hp = h / pool, wp = w / pool, vt = (C + Ct) / 8;
  const int64_t total = (int64_t)B * hp * wp * N * vt;
  const float inv = 1.f / (pool * pool);
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % vt) * 8;
    int64_t t = idx / vt;
    const int n = (int)(t % N); t /= N;
    const int px = (int)(t % wp), py = (int)((t / wp) % hp), b = (int)(t / ((int64_t)wp * hp));
    float s[8];
    if (c8 >= C) {
      ldg8f(text + n * Ct + c8 - C, s);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] = 0.f;
      const __nv_bfloat16* base = x + ((((int64_t)b * N + n) * h + py * pool) * w + px * pool) * ldx + c8;
      for (int i = 0; i < pool; ++i) {
#pragma unroll 4
        for (int j = 0; j < pool; ++j) {
          float f[8];
          bf16x8_to_f32(ld_stream16(base + ((int64_t)i * w + j) * ldx), f);
#pragma unroll
          for (int k = 0; k < 8; ++k) s[k] += f[k];
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] *= inv;
    }
    float4* o = (float4*)(tok + (idx / vt) * (int64_t)(C + Ct) + c8);
    o[0] = make_float4(s[0], s[1], s[2], s[3]);
    o[1] = make_float4(s[4], s[5], s[6], s[7]);
  }
}
// out = x + bilinear_ac(tok): grid = (chunks of 512 items, B * N maps), item = (pixel, 8-channel vector), two items per thread
__global__ void __launch_bounds__(256)
unpool_add_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const float* __restrict__ tok, int64_t ldt, __nv_bfloat16* __restrict__ out,
                       int64_t ldo, int N, int h, int w, int C, int hp, int wp) {
  const int vpp = C / 8, items = h * w * vpp;
  const int map = blockIdx.y, b = map / N, n = map - b * N;
  const float sy = hp > 1 && h > 1 ? (float)(hp - 1) / (h - 1) : 0.f, sx = wp > 1 && w > 1 ? (float)(wp - 1) / (w - 1) : 0.f;
  const int base = blockIdx.x * 512 + threadIdx.x;
  uint4 xv[2];
  int pix[2], c8[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int idx = base + u * 256 < items ? base + u * 256 : 0;
    pix[u] = idx / vpp;
    c8[u] = (idx - pix[u] * vpp) * 8;
    xv[u] = ld_stream16(x + ((int64_t)map * h * w + pix[u]) * ldx + c8[u]);
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (base + u * 256 >= items) break;
    const int yy = pix[u] / w, xx = pix[u] - yy * w;
    int y0, y1, x0, x1;
    float wy, wx;
    bilin_src(yy, sy, hp, y0, y1, wy);
    bilin_src(xx, sx, wp, x0, x1, wx);
    float f[8], t00[8], t01[8], t10[8], t11[8];
    ldg8f(tok + ((((int64_t)b * hp + y0) * wp + x0) * N + n) * ldt + c8[u], t00);
    ldg8f(tok + ((((int64_t)b * hp + y0) * wp + x1) * N + n) * ldt + c8[u], t01);
    ldg8f(tok + ((((int64_t)b * hp + y1) * wp + x0) * N + n) * ldt + c8[u], t10);
    ldg8f(tok + ((((int64_t)b * hp + y1) * wp + x1) * N + n) * ldt + c8[u], t11);
    bf16x8_to_f32(xv[u], f);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      f[i] += (1.f - wy) * ((1.f - wx) * t00[i] + wx * t01[i]) + wy * ((1.f - wx) * t10[i] + wx * t11[i]);       // same expression as the generic kernel
    *(uint4*)(out + ((int64_t)map * h * w + pix[u]) * ldo + c8[u]) = f32_to_bf16x8(f);
  }
}
// cat[(b,n), Y, X, c0 + c] = bilinear_ac(skip[b]) for every class n: a thread interpolates 8 channels once and stores them N times
__global__ void __launch_bounds__(256)
skip_fill_bf16_kernel(const float* __restrict__ skip, int64_t lds, __nv_bfloat16* __restrict__ cat, int64_t ldc, int c0, int B, int N, int h, int w,
                      int Cs, int H2, int W2) {
  const int vs = Cs / 8;
  const int64_t total = (int64_t)B * H2 * W2 * vs;
  const float sy = h > 1 && H2 > 1 ? (float)(h - 1) / (H2 - 1) : 0.f, sx = w > 1 && W2 > 1 ? (float)(w - 1) / (W2 - 1) : 0.f;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % vs) * 8;
    int64_t pix = idx / vs;
    const int X = (int)(pix % W2), Y = (int)((pix / W2) % H2), b = (int)(pix / ((int64_t)W2 * H2));
    int y0, y1, x0, x1;
    float wy, wx;
    bilin_src(Y, sy, h, y0, y1, wy);
    bilin_src(X, sx, w, x0, x1, wx);
    float s00[8], s01[8], s10[8], s11[8], v[8];
    ldg8f(skip + (((int64_t)b * h + y0) * w + x0) * lds + c, s00);
    ldg8f(skip + (((int64_t)b * h + y0) * w + x1) * lds + c, s01);
    ldg8f(skip + (((int64_t)b * h + y1) * w + x0) * lds + c, s10);
    ldg8f(skip + (((int64_t)b * h + y1) * w + x1) * lds + c, s11);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (1.f - wy) * ((1.f - wx) * s00[i] + wx * s01[i]) + wy * ((1.f - wx) * s10[i] + wx * s11[i]);
    const uint4 pk = f32_to_bf16x8(v);
    __nv_bfloat16* dst = cat + (((int64_t)b * N * H2 + Y) * W2 + X) * ldc + c0 + c;
    for (int n = 0; n < N; ++n) *(uint4*)(dst + (int64_t)n * H2 * W2 * ldc) = pk;
  }
}
// x[map, pix, c] += scale * v[map, c], four channels per thread (f32 x, f32 v)
__global__ void map_bcast_add_f32x4_kernel(float* __restrict__ x, const float* __restrict__ v, int64_t ldv, int64_t maps, int hw, int C, float scale) {
  const int c4n = C / 4;
  const int64_t total = maps * hw * c4n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c4n) * 4;
    const int64_t map = idx / ((int64_t)hw * c4n);
    const float4 a = __ldg((const float4*)(v + map * ldv + c));
    float4* d = (float4*)x + idx;
    float4 t = *d;
    t.x += scale * a.x; t.y += scale * a.y; t.z += scale * a.z; t.w += scale * a.w;
    *d = t;
  }
}

inline int gn_splits(int64_t maps, int hw, int C) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("SVL_GN_SPLITS"); forced = e ? atoi(e) : 0; }
  if (forced > 0) return forced;
  // ~16 pixel iterations per thread.  The split count depends on the map geometry only -- never on the number of maps -- so the
  // order of the partial sums, and with it every statistic, is independent of the batch an image is processed in.
  (void)maps;
  int64_t s = (int64_t)hw * (C / 16) / (kGnThreads * 16);
  return (int)(s < 1 ? 1 : (s > 256 ? 256 : s));
}

// ---------------------------------------------------------------------------------------------- 7x7 im2col of the similarity maps
// sim f32 [B*hw, ld_sim] (pixel-major, class = column)  ->  out operand [(b, n, y, x), kpad] with column t = (dy+r)*ks + (dx+r)
__global__ void sim_im2col_kernel(const float* __restrict__ sim, int64_t ld_sim, void* __restrict__ out, int out_dtype, int64_t ldo, int B, int N,
                                  int h, int w, int ks, int kpad) {
  const int r = ks / 2;
  const int groups = kpad / 8;
  const int64_t total = (int64_t)B * N * h * w * groups;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int gi = (int)(idx % groups);
    int64_t row = idx / groups;
    const int x = (int)(row % w), y = (int)((row / w) % h);
    const int n = (int)((row / ((int64_t)w * h)) % N), b = (int)(row / ((int64_t)w * h * N));
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = gi * 8 + i;
      float v = 0.f;
      if (t < ks * ks) {
        const int yy = y + t / ks - r, xx = x + t % ks - r;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) v = __ldg(sim + ((int64_t)(b * h + yy) * w + xx) * ld_sim + n);
      }
      f[i] = v;
    }
    st8(out, out_dtype, row * ldo + gi * 8, ldo / 2, 8, f);
  }
}
// transpose of the above: dsim[(b,y,x), n] = sum_t dcol[(b, n, y - dy_t, x - dx_t), t]; columns >= N of dsim are zeroed
__global__ void sim_col2im_kernel(const void* __restrict__ dcol, int dtype, int64_t ld, void* __restrict__ dsim, int ds_dtype, int64_t ld_ds,
                                  int ncols, int B, int N, int h, int w, int ks) {
  const int r = ks / 2;
  const int64_t total = (int64_t)B * h * w * ncols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx % ncols);
    const int64_t pix = idx / ncols;
    const int x = (int)(pix % w), y = (int)((pix / w) % h), b = (int)(pix / ((int64_t)w * h));
    float s = 0.f;
    if (n < N) {
      for (int t = 0; t < ks * ks; ++t) {
        const int yy = y - (t / ks - r), xx = x - (t % ks - r);
        if (yy >= 0 && yy < h && xx >= 0 && xx < w)
          s += load_as_f32(dcol, dtype, ((((int64_t)b * N + n) * h + yy) * w + xx) * ld + t, ld / 2);
      }
    }
    store_from_f32(dsim, ds_dtype, pix * ld_ds + n, s, ld_ds / 2);
  }
}

// ---------------------------------------------------------------------------------------------- per-map spatial sums / broadcasts
// out[map, c] = scale * sum_pix x[map, pix, c]
__global__ void map_sum_kernel(const void* __restrict__ x, int dtype, int64_t ld, float* __restrict__ out, int hw, int C, float scale) {
  __shared__ float red[8][129];
  const int map = blockIdx.x;
  const int c = threadIdx.x;           // blockDim = (C, 1024 / C ... ) handled by caller: blockDim.x = C (<=128), blockDim.y = rows
  float s = 0.f;
  for (int p = threadIdx.y; p < hw; p += blockDim.y) s += load_as_f32(x, dtype, ((int64_t)map * hw + p) * ld + c, ld / 2);
  red[threadIdx.y][c] = s;
  __syncthreads();
  if (threadIdx.y == 0) {
    float t = 0.f;
    for (int i = 0; i < blockDim.y; ++i) t += red[i][c];
    out[(int64_t)map * C + c] = t * scale;
  }
}
// x[map, pix, c] += scale * v[map, c]   (f32 x)
__global__ void map_bcast_add_kernel(float* __restrict__ x, const void* __restrict__ v, int v_dtype, int64_t ldv, int64_t maps, int hw, int C,
                                     float scale) {
  const int64_t total = maps * hw * C;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int64_t map = idx / ((int64_t)hw * C);
    x[idx] += scale * load_as_f32(v, v_dtype, map * ldv + c, ldv / 2);
  }
}

// ---------------------------------------------------------------------------------------------- SemanticTransformer pooling
// tok[(b, py, px), n, 0:C] = avgpool_{pool x pool} x[(b, n), :, :, 0:C];  tok[..., C:C+Ct] = text[n, :]      (vlg_head.py:41-51)
__global__ void pool_tokens_kernel(const void* __restrict__ x, int x_dtype, int64_t ldx, const float* __restrict__ text, float* __restrict__ tok,
                                   int B, int N, int h, int w, int C, int Ct, int pool) {
  const int hp = h / pool, wp = w / pool;
  const int Cd = C + Ct;
  const int64_t total = (int64_t)B * hp * wp * N * Cd;
  const float inv = 1.f / (pool * pool);
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Cd);
    int64_t t = idx / Cd;
    const int n = (int)(t % N); t /= N;
    const int px = (int)(t % wp), py = (int)((t / wp) % hp), b = (int)(t / ((int64_t)wp * hp));
    float v;
    if (c >= C) {
      v = text[n * Ct + c - C];
    } else {
      float s = 0.f;
      const int64_t base = (((int64_t)b * N + n) * h + py * pool) * w + px * pool;
      for (int i = 0; i < pool; ++i)
        for (int j = 0; j < pool; ++j) s += load_as_f32(x, x_dtype, (base + (int64_t)i * w + j) * ldx + c, ldx / 2);
      v = s * inv;
    }
    tok[idx] = v;
  }
}

__device__ __forceinline__ void bilin_src(int o, float scale, int in_size, int& i0, int& i1, float& w1) {
  // align_corners=True source coordinate
  const float s = o * scale;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + 1 < in_size ? i0 + 1 : i0;
  w1 = s - i0;
}

// out[(b,n), y, x, c] = x[(b,n), y, x, c] + bilinear_ac(tok[(b, :, :), n, c])   (vlg_head.py:60-66)
__global__ void unpool_add_kernel(const void* __restrict__ x, int x_dtype, int64_t ldx, const float* __restrict__ tok, int64_t ldt,
                                  void* __restrict__ out, int out_dtype, int64_t ldo, int B, int N, int h, int w, int C, int hp, int wp) {
  const int vpp = C / 8;
  const int64_t total = (int64_t)B * N * h * w * vpp;
  const float sy = hp > 1 && h > 1 ? (float)(hp - 1) / (h - 1) : 0.f, sx = wp > 1 && w > 1 ? (float)(wp - 1) / (w - 1) : 0.f;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % vpp) * 8;
    int64_t pix = idx / vpp;
    const int xx = (int)(pix % w), yy = (int)((pix / w) % h);
    const int n = (int)((pix / ((int64_t)w * h)) % N), b = (int)(pix / ((int64_t)w * h * N));
    int y0, y1, x0, x1;
    float wy, wx;
    bilin_src(yy, sy, hp, y0, y1, wy);
    bilin_src(xx, sx, wp, x0, x1, wx);
    float f[8];
    ld8(x, x_dtype, pix * ldx + c8, ldx / 2, 8, f);
    const float* t00 = tok + ((((int64_t)b * hp + y0) * wp + x0) * N + n) * ldt + c8;
    const float* t01 = tok + ((((int64_t)b * hp + y0) * wp + x1) * N + n) * ldt + c8;
    const float* t10 = tok + ((((int64_t)b * hp + y1) * wp + x0) * N + n) * ldt + c8;
    const float* t11 = tok + ((((int64_t)b * hp + y1) * wp + x1) * N + n) * ldt + c8;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      f[i] += (1.f - wy) * ((1.f - wx) * t00[i] + wx * t01[i]) + wy * ((1.f - wx) * t10[i] + wx * t11[i]);
    st8(out, out_dtype, pix * ldo + c8, ldo / 2, 8, f);
  }
}
// dtok[(b,py,px), n, c] = sum_{y,x} W(y,py) W(x,px) dout[(b,n), y, x, c]   for c < C;  columns [C, ldt) are zeroed
__global__ void unpool_bwd_kernel(const void* __restrict__ dout, int dtype, int64_t ld, float* __restrict__ dtok, int64_t ldt, int B, int N, int h,
                                  int w, int C, int hp, int wp) {
  const int64_t total = (int64_t)B * hp * wp * N * ldt;
  const float sy = hp > 1 && h > 1 ? (float)(hp - 1) / (h - 1) : 0.f, sx = wp > 1 && w > 1 ? (float)(wp - 1) / (w - 1) : 0.f;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % ldt);
    int64_t t = idx / ldt;
    const int n = (int)(t % N); t /= N;
    const int px = (int)(t % wp), py = (int)((t / wp) % hp), b = (int)(t / ((int64_t)wp * hp));
    float s = 0.f;
    if (c < C) {
      // output rows whose source interval touches py
      const int ylo = sy > 0.f ? max(0, (int)ceilf((py - 1) / sy)) : 0, yhi = sy > 0.f ? min(h - 1, (int)floorf((py + 1) / sy)) : h - 1;
      const int xlo = sx > 0.f ? max(0, (int)ceilf((px - 1) / sx)) : 0, xhi = sx > 0.f ? min(w - 1, (int)floorf((px + 1) / sx)) : w - 1;
      for (int yy = ylo; yy <= yhi; ++yy) {
        int y0, y1; float wy;
        bilin_src(yy, sy, hp, y0, y1, wy);
        const float ay = (y0 == py ? 1.f - wy : 0.f) + (y1 == py ? wy : 0.f);
        if (ay == 0.f) continue;
        for (int xx = xlo; xx <= xhi; ++xx) {
          int x0, x1; float wx;
          bilin_src(xx, sx, wp, x0, x1, wx);
          const float ax = (x0 == px ? 1.f - wx : 0.f) + (x1 == px ? wx : 0.f);
          if (ax == 0.f) continue;
          s += ay * ax * load_as_f32(dout, dtype, ((((int64_t)b * N + n) * h + yy) * w + xx) * ld + c, ld / 2);
        }
      }
    }
    dtok[idx] = s;
  }
}
// vectorised variant of the above (bf16 or f32 gradient): a thread owns 8 channels of one token (16-byte loads instead of single elements)
template <bool BF16>
__global__ void __launch_bounds__(256)
unpool_bwd_vec_kernel(const void* __restrict__ dout_, int64_t ld, float* __restrict__ dtok, int64_t ldt, int B, int N, int h, int w, int C,
                      int hp, int wp) {
  const __nv_bfloat16* dout = (const __nv_bfloat16*)dout_;
  const float* doutf = (const float*)dout_;
  const int vt = (int)(ldt / 8);
  const int64_t total = (int64_t)B * hp * wp * N * vt;
  const float sy = hp > 1 && h > 1 ? (float)(hp - 1) / (h - 1) : 0.f, sx = wp > 1 && w > 1 ? (float)(wp - 1) / (w - 1) : 0.f;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % vt) * 8;
    int64_t t = idx / vt;
    const int n = (int)(t % N); t /= N;
    const int px = (int)(t % wp), py = (int)((t / wp) % hp), b = (int)(t / ((int64_t)wp * hp));
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.f;
    if (c8 < C) {
      const int ylo = sy > 0.f ? max(0, (int)ceilf((py - 1) / sy)) : 0, yhi = sy > 0.f ? min(h - 1, (int)floorf((py + 1) / sy)) : h - 1;
      const int xlo = sx > 0.f ? max(0, (int)ceilf((px - 1) / sx)) : 0, xhi = sx > 0.f ? min(w - 1, (int)floorf((px + 1) / sx)) : w - 1;
      const int64_t base = ((int64_t)b * N + n) * h * w * ld + c8;
      for (int yy = ylo; yy <= yhi; ++yy) {
        int y0, y1; float wy;
        bilin_src(yy, sy, hp, y0, y1, wy);
        const float ay = (y0 == py ? 1.f - wy : 0.f) + (y1 == py ? wy : 0.f);
        if (ay == 0.f) continue;
        for (int xx = xlo; xx <= xhi; ++xx) {
          int x0, x1; float wx;
          bilin_src(xx, sx, wp, x0, x1, wx);
          const float ax = (x0 == px ? 1.f - wx : 0.f) + (x1 == px ? wx : 0.f);
          if (ax == 0.f) continue;
          float f[8];
          if (BF16) bf16x8_to_f32(__ldg((const uint4*)(dout + base + ((int64_t)yy * w + xx) * ld)), f);
          else ldg8f(doutf + base + ((int64_t)yy * w + xx) * ld, f);
          const float a = ay * ax;
#pragma unroll
          for (int i = 0; i < 8; ++i) s[i] += a * f[i];
        }
      }
    }
    float4* o = (float4*)(dtok + (idx / vt) * ldt + c8);
    o[0] = make_float4(s[0], s[1], s[2], s[3]);
    o[1] = make_float4(s[4], s[5], s[6], s[7]);
  }
}
// dx[(b,n), y, x, c] (+)= dtok[(b, y/pool, x/pool), n, c] / pool^2  for y < hp*pool, x < wp*pool        (f32 dx, float4 per thread)
__global__ void pool_tokens_bwd_kernel(const float* __restrict__ dtok, int64_t ldt, float* __restrict__ dx, int B, int N, int h, int w, int C,
                                       int pool) {
  const int hp = h / pool, wp = w / pool;
  const int c4 = C / 4;
  const int64_t total = (int64_t)B * N * h * w * c4;
  const float inv = 1.f / (pool * pool);
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c4) * 4;
    int64_t pix = idx / c4;
    const int xx = (int)(pix % w), yy = (int)((pix / w) % h);
    const int n = (int)((pix / ((int64_t)w * h)) % N), b = (int)(pix / ((int64_t)w * h * N));
    const int py = yy / pool, px = xx / pool;
    if (py < hp && px < wp) {
      const float4 t = *(const float4*)(dtok + ((((int64_t)b * hp + py) * wp + px) * N + n) * ldt + c);
      float4* d = (float4*)(dx + pix * C + c);
      float4 v = *d;
      v.x += inv * t.x; v.y += inv * t.y; v.z += inv * t.z; v.w += inv * t.w;
      *d = v;
    }
  }
}

// dx[(b,n), y, x, c] = src[(b,n), y, x, c] + dtok[(b, y/pool, x/pool), n, c] / pool^2  (f32 dx written once: the cast of the incoming gradient and the
// pooling gradient in one pass instead of a cast kernel followed by a read-modify-write); src bf16 or f32; 8 channels per thread
template <bool BF16>
__global__ void __launch_bounds__(256)
pool_tokens_bwd_from_kernel(const void* __restrict__ src_, int64_t lds, const float* __restrict__ dtok, int64_t ldt, float* __restrict__ dx, int B,
                            int N, int h, int w, int C, int pool) {
  const int hp = h / pool, wp = w / pool, vpp = C / 8;
  const int64_t total = (int64_t)B * N * h * w * vpp;
  const float inv = 1.f / (pool * pool);
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % vpp) * 8;
    const int64_t pix = idx / vpp;
    const int xx = (int)(pix % w), yy = (int)((pix / w) % h);
    const int n = (int)((pix / ((int64_t)w * h)) % N), b = (int)(pix / ((int64_t)w * h * N));
    float f[8];
    if (BF16) bf16x8_to_f32(ld_stream16((const __nv_bfloat16*)src_ + pix * lds + c), f);
    else ldg8f((const float*)src_ + pix * lds + c, f);
    const int py = yy / pool, px = xx / pool;
    if (py < hp && px < wp) {
      float t[8];
      ldg8f(dtok + ((((int64_t)b * hp + py) * wp + px) * N + n) * ldt + c, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += inv * t[i];
    }
    float4* d = (float4*)(dx + pix * C + c);
    d[0] = make_float4(f[0], f[1], f[2], f[3]);
    d[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
}

// ---------------------------------------------------------------------------------------------- skip features of `Up`
// cat[(b,n), Y, X, c0 + c] = bilinear_ac(skip[b, :, :, c]) for every class n     (vlg_head.py:127-131: resize + repeat + cat)
__global__ void skip_fill_kernel(const void* __restrict__ skip, int s_dtype, int64_t lds, void* __restrict__ cat, int c_dtype, int64_t ldc, int c0,
                                 int B, int N, int h, int w, int Cs, int H2, int W2) {
  const int64_t total = (int64_t)B * H2 * W2 * Cs;
  const float sy = h > 1 && H2 > 1 ? (float)(h - 1) / (H2 - 1) : 0.f, sx = w > 1 && W2 > 1 ? (float)(w - 1) / (W2 - 1) : 0.f;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Cs);
    int64_t pix = idx / Cs;
    const int X = (int)(pix % W2), Y = (int)((pix / W2) % H2), b = (int)(pix / ((int64_t)W2 * H2));
    int y0, y1, x0, x1;
    float wy, wx;
    bilin_src(Y, sy, h, y0, y1, wy);
    bilin_src(X, sx, w, x0, x1, wx);
    auto S = [&](int yy, int xx) { return load_as_f32(skip, s_dtype, (((int64_t)b * h + yy) * w + xx) * lds + c, lds / 2); };
    const float v = (1.f - wy) * ((1.f - wx) * S(y0, x0) + wx * S(y0, x1)) + wy * ((1.f - wx) * S(y1, x0) + wx * S(y1, x1));
    for (int n = 0; n < N; ++n)
      store_from_f32(cat, c_dtype, ((((int64_t)b * N + n) * H2 + Y) * W2 + X) * ldc + c0 + c, v, ldc / 2);
  }
}
// out[b, p, c] = sum_n x[(b, n), p, c0 + c]   (f32 out; the per-class copies of the skip channels collapse back onto the image:
// transpose of the `repeat` in vlg_head.py:129).  One thread per (b, pixel, 8 channels): N coalesced 16-byte loads.
__global__ void class_sum_kernel(const void* __restrict__ x, int dtype, int64_t ld, int c0, float* __restrict__ out, int B, int N, int64_t P, int Cs) {
  const int c8n = Cs / 8;
  const int64_t total = (int64_t)B * P * c8n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c8n) * 8;
    const int64_t pix = (idx / c8n) % P, b = idx / (c8n * P);
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.f;
    for (int n = 0; n < N; ++n) {
      float f[8];
      ld8(x, dtype, (((int64_t)b * N + n) * P + pix) * ld + c0 + c, ld / 2, 8, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += f[i];
    }
    st8(out, SVL_F32, ((int64_t)b * P + pix) * Cs + c, 0, 8, s);
  }
}

// bf16 fast path of class_sum: the N per-class loads of a thread are issued four at a time
__global__ void __launch_bounds__(256)
class_sum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, int c0, float* __restrict__ out, int B, int N, int64_t P, int Cs) {
  const int c8n = Cs / 8;
  const int64_t total = (int64_t)B * P * c8n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c8n) * 8;
    const int64_t pix = (idx / c8n) % P, b = idx / (c8n * P);
    const __nv_bfloat16* base = x + ((int64_t)b * N * P + pix) * ld + c0 + c;
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.f;
    for (int n = 0; n < N; n += 4) {
      uint4 r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) r[u] = ld_stream16(base + (int64_t)(n + u < N ? n + u : n) * P * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (n + u < N) {
          float f[8];
          bf16x8_to_f32(r[u], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) s[i] += f[i];
        }
      }
    }
    float4* o = (float4*)(out + ((int64_t)b * P + pix) * Cs + c);
    o[0] = make_float4(s[0], s[1], s[2], s[3]);
    o[1] = make_float4(s[4], s[5], s[6], s[7]);
  }
}

// dskip_pre[b, y, x, c] = relu'(skip) * sum_n sum_{Y,X} W(Y,y) W(X,x) dcat[(b,n), Y, X, c0 + c]      (8 channels per thread)
__global__ void skip_grad_kernel(const void* __restrict__ dcat, int d_dtype, int64_t ldd, int c0, const void* __restrict__ skip, int s_dtype,
                                 int64_t lds, void* __restrict__ dskip, int o_dtype, int64_t ldo, int B, int N, int h, int w, int Cs, int H2,
                                 int W2) {
  const int c8n = Cs / 8;
  const int64_t total = (int64_t)B * h * w * c8n;
  const float sy = h > 1 && H2 > 1 ? (float)(h - 1) / (H2 - 1) : 0.f, sx = w > 1 && W2 > 1 ? (float)(w - 1) / (W2 - 1) : 0.f;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c8n) * 8;
    int64_t pix = idx / c8n;
    const int x = (int)(pix % w), y = (int)((pix / w) % h), b = (int)(pix / ((int64_t)w * h));
    float s[8], sk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.f;
    ld8(skip, s_dtype, pix * lds + c, lds / 2, 8, sk);
    const int Ylo = sy > 0.f ? max(0, (int)ceilf((y - 1) / sy)) : 0, Yhi = sy > 0.f ? min(H2 - 1, (int)floorf((y + 1) / sy)) : H2 - 1;
    const int Xlo = sx > 0.f ? max(0, (int)ceilf((x - 1) / sx)) : 0, Xhi = sx > 0.f ? min(W2 - 1, (int)floorf((x + 1) / sx)) : W2 - 1;
    for (int Y = Ylo; Y <= Yhi; ++Y) {
      int y0, y1; float wy;
      bilin_src(Y, sy, h, y0, y1, wy);
      const float ay = (y0 == y ? 1.f - wy : 0.f) + (y1 == y ? wy : 0.f);
      if (ay == 0.f) continue;
      for (int X = Xlo; X <= Xhi; ++X) {
        int x0, x1; float wx;
        bilin_src(X, sx, w, x0, x1, wx);
        const float ax = (x0 == x ? 1.f - wx : 0.f) + (x1 == x ? wx : 0.f);
        if (ax == 0.f) continue;
        const float wgt = ay * ax;
        for (int n = 0; n < N; ++n) {
          float f[8];
          ld8(dcat, d_dtype, ((((int64_t)b * N + n) * H2 + Y) * W2 + X) * ldd + c0 + c, ldd / 2, 8, f);
#pragma unroll
          for (int i = 0; i < 8; ++i) s[i] += wgt * f[i];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = sk[i] > 0.f ? s[i] : 0.f;
    st8(dskip, o_dtype, pix * ldo + c, ldo / 2, 8, s);
  }
}

// ---------------------------------------------------------------------------------------------- output conv 3x3, C -> 1
// out[map, y, x] = bias + sum_{t, c} w[t*C + c] * x[map, y+dy_t, x+dx_t, c]            (vlg_head.py:190,239-240)
// Tile = 8 x 32 output pixels.  Every pixel of the (8+2) x (32+2) halo region is read ONCE: its 9 per-tap partial dot products
// z_t = <w[t, :], x[p, :]> go to shared memory, then each output pixel sums the 9 shifted entries.
constexpr int kO1TH = 8, kO1TW = 32, kO1HW = (kO1TH + 2) * (kO1TW + 2);
__global__ void __launch_bounds__(256)
conv_out1_fwd_kernel(const void* __restrict__ x, int dtype, int64_t ld, const float* __restrict__ wgt, const float* __restrict__ bias,
                     float* __restrict__ out, int h, int w, int C, int tiles_x, int tiles_y) {
  extern __shared__ float s_o1[];       // [9*C] weights, then [9][kO1HW] partial sums
  float* s_w = s_o1;
  float* s_z = s_o1 + 9 * C;
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) s_w[i] = wgt[i];
  const int tx = blockIdx.x % tiles_x, ty = (blockIdx.x / tiles_x) % tiles_y;
  const int64_t map = blockIdx.x / (tiles_x * tiles_y);
  const int x0 = tx * kO1TW - 1, y0 = ty * kO1TH - 1;
  __syncthreads();
  for (int i = threadIdx.x; i < kO1HW; i += blockDim.x) {
    const int yy = y0 + i / (kO1TW + 2), xx = x0 + i % (kO1TW + 2);
    float z[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) z[t] = 0.f;
    if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
      const int64_t base = ((map * h + yy) * w + xx) * ld;
      for (int c = 0; c < C; c += 8) {
        float f[8];
        ld8(x, dtype, base + c, ld / 2, 8, f);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
#pragma unroll
          for (int k = 0; k < 8; ++k) z[t] += f[k] * s_w[t * C + c + k];
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) s_z[t * kO1HW + i] = z[t];
  }
  __syncthreads();
  const int ly = threadIdx.x / kO1TW, lx = threadIdx.x % kO1TW;
  const int yy = ty * kO1TH + ly, xx = tx * kO1TW + lx;
  if (yy < h && xx < w) {
    float s = bias[0];
#pragma unroll
    for (int t = 0; t < 9; ++t) s += s_z[t * kO1HW + (ly + t / 3) * (kO1TW + 2) + lx + t % 3];
    out[(map * h + yy) * w + xx] = s;
  }
}
// dx[map, y, x, c] = sum_t w[t*C + c] * dout[map, y-dy_t, x-dx_t]
__global__ void conv_out1_dgrad_kernel(const float* __restrict__ dout, const float* __restrict__ wgt, void* __restrict__ dx, int dtype, int64_t ld,
                                       int64_t maps, int h, int w, int C) {
  extern __shared__ float s_w[];
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) s_w[i] = wgt[i];
  __syncthreads();
  const int vpp = C / 8;
  const int64_t total = maps * h * w * vpp;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % vpp) * 8;
    const int64_t pix = idx / vpp;
    const int xx = (int)(pix % w), yy = (int)((pix / w) % h);
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = 0.f;
    for (int t = 0; t < 9; ++t) {
      const int y2 = yy - (t / 3 - 1), x2 = xx - (t % 3 - 1);
      if (y2 < 0 || y2 >= h || x2 < 0 || x2 >= w) continue;
      const float d = dout[pix - (int64_t)(t / 3 - 1) * w - (t % 3 - 1)];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += d * s_w[t * C + c8 + i];
    }
    st8(dx, dtype, pix * ld + c8, ld / 2, 8, f);
  }
}
// dw[t*C + c] += sum_q x[q, c] * dout[q - tap_t];  dbias += sum dout
// Persistent CTAs over 8 x 32 pixel tiles: the (8+2) x (32+2) dout halo goes to shared memory, a warp walks the tile's pixels with
// lane = channel (one 64-byte x row per pixel, 9 broadcast smem reads), 9 accumulators per lane live across all tiles of the CTA.
__global__ void __launch_bounds__(256)
conv_out1_wgrad_kernel(const float* __restrict__ dout, const void* __restrict__ x, int dtype, int64_t ld, float* __restrict__ dw,
                       float* __restrict__ dbias, int h, int w, int C, int tiles_x, int tiles_y, int64_t num_tiles) {
  __shared__ float s_d[kO1HW];
  __shared__ float red[8][9][33];
  __shared__ float redb[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + lane;
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    float sb = 0.f;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tx = (int)(tile % tiles_x), ty = (int)((tile / tiles_x) % tiles_y);
      const int64_t map = tile / ((int64_t)tiles_x * tiles_y);
      const int x0 = tx * kO1TW - 1, y0 = ty * kO1TH - 1;
      __syncthreads();
      for (int i = threadIdx.x; i < kO1HW; i += blockDim.x) {
        const int yy = y0 + i / (kO1TW + 2), xx = x0 + i % (kO1TW + 2);
        s_d[i] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? dout[(map * h + yy) * w + xx] : 0.f;
      }
      __syncthreads();
      // warp `warp` takes tile row `warp`, 32 pixels
      const int yy = ty * kO1TH + warp;
      if (yy < h) {
#pragma unroll 4
        for (int lx = 0; lx < kO1TW; ++lx) {
          const int xx = tx * kO1TW + lx;
          if (xx >= w) break;
          const float xv = c < C ? load_as_f32(x, dtype, ((map * h + yy) * w + xx) * ld + c, ld / 2) : 0.f;
          // output pixel (yy - dy, xx - dx) reads this input pixel through tap (dy, dx): halo index (warp + 1 - dy, lx + 1 - dx)
#pragma unroll
          for (int t = 0; t < 9; ++t) acc[t] += xv * s_d[(warp + 1 - (t / 3 - 1)) * (kO1TW + 2) + lx + 1 - (t % 3 - 1)];
          if (c0 == 0) sb += s_d[(warp + 1) * (kO1TW + 2) + lx + 1];
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) red[warp][t][lane] = acc[t];
    if (lane == 0) redb[warp] = sb;
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * 32; i += blockDim.x) {
      const int t = i / 32, l = i % 32;
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) v += red[k][t][l];
      if (c0 + l < C && v != 0.f) atomicAdd(dw + t * C + c0 + l, v);
    }
    if (threadIdx.x == 0 && c0 == 0) {
      float v = 0.f;
      for (int k = 0; k < 8; ++k) v += redb[k];
      atomicAdd(dbias, v);
    }
    __syncthreads();
  }
}


// ---------------------------------------------------------------------------------------------- output conv, all-bf16 fast paths (C = 32)
// The generic kernels above are instruction-bound, not HBM-bound (0.9 TB/s on 352 MB: one scalar shared-memory weight read per FMA,
// 2-byte x loads, run-time dtype dispatch).  Here: 16-byte loads issued in batches, weights as float4 broadcasts (forward) or
// resident in registers (backward: a thread owns one 8-channel vector for all its pixels).
constexpr int kO1C = 32;
__global__ void __launch_bounds__(256)
conv_out1_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, const float* __restrict__ wgt, const float* __restrict__ bias,
                          float* __restrict__ out, int h, int w, int tiles_x, int tiles_y) {
  __shared__ float4 s_w4[9 * kO1C / 4];
  __shared__ float s_z[9 * kO1HW];
  for (int i = threadIdx.x; i < 9 * kO1C / 4; i += blockDim.x) s_w4[i] = __ldg((const float4*)wgt + i);
  const int tx = blockIdx.x % tiles_x, ty = (blockIdx.x / tiles_x) % tiles_y;
  const int64_t map = blockIdx.x / (tiles_x * tiles_y);
  const int x0 = tx * kO1TW - 1, y0 = ty * kO1TH - 1;
  // the (at most two) halo pixels of this thread: all eight 16-byte loads are issued before the first FMA
  uint4 v[2][4];
  bool ok[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int i = threadIdx.x + u * 256;
    const int yy = y0 + i / (kO1TW + 2), xx = x0 + i % (kO1TW + 2);
    ok[u] = i < kO1HW && yy >= 0 && yy < h && xx >= 0 && xx < w;
    const __nv_bfloat16* px = x + ((map * h + (ok[u] ? yy : 0)) * w + (ok[u] ? xx : 0)) * ld;
#pragma unroll
    for (int q = 0; q < 4; ++q) v[u][q] = ld_stream16(px + q * 8);
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int i = threadIdx.x + u * 256;
    if (i < kO1HW) {
      float z[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) z[t] = 0.f;
      if (ok[u]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float f[8];
          bf16x8_to_f32(v[u][q], f);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const float4 wa = s_w4[t * (kO1C / 4) + q * 2], wb = s_w4[t * (kO1C / 4) + q * 2 + 1];
            z[t] += f[0] * wa.x + f[1] * wa.y + f[2] * wa.z + f[3] * wa.w + f[4] * wb.x + f[5] * wb.y + f[6] * wb.z + f[7] * wb.w;
          }
        }
      }
#pragma unroll
      for (int t = 0; t < 9; ++t) s_z[t * kO1HW + i] = z[t];
    }
  }
  __syncthreads();
  const int ly = threadIdx.x / kO1TW, lx = threadIdx.x % kO1TW;
  const int yy = ty * kO1TH + ly, xx = tx * kO1TW + lx;
  if (yy < h && xx < w) {
    float s = __ldg(bias);
#pragma unroll
    for (int t = 0; t < 9; ++t) s += s_z[t * kO1HW + (ly + t / 3) * (kO1TW + 2) + lx + t % 3];
    out[(map * h + yy) * w + xx] = s;
  }
}

// Backward tiles of the output conv: a CTA owns kB1TH x kB1TW pixels of one map and keeps their dout halo ((TH + 2) x (TW + 2) floats) in
// shared memory; item = (pixel, 8-channel vector), 4 lanes per pixel read the same nine halo entries (broadcast).  The first version read the
// nine neighbours of every item with predicated global loads and their index arithmetic: 333 / 270 us (data / weight gradient) on a problem
// whose HBM time is 57 us each.
constexpr int kB1TH = 8, kB1TW = 64, kB1HW = (kB1TH + 2) * (kB1TW + 2);
__device__ __forceinline__ void out1_load_halo(float* s_d, const float* __restrict__ dmap, int y0, int x0, int h, int w) {
  for (int i = threadIdx.x; i < kB1HW; i += blockDim.x) {
    const int yy = y0 - 1 + i / (kB1TW + 2), xx = x0 - 1 + i % (kB1TW + 2);
    s_d[i] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(dmap + yy * w + xx) : 0.f;
  }
}
// d[t] = dout[yy - (t/3 - 1), xx - (t%3 - 1)] for the tile pixel (ly, lx): halo entry (ly + 2 - t/3, lx + 2 - t%3)
__device__ __forceinline__ void out1_halo_neighbours(const float* s_d, int ly, int lx, float* d) {
#pragma unroll
  for (int t = 0; t < 9; ++t) d[t] = s_d[(ly + 2 - t / 3) * (kB1TW + 2) + lx + 2 - t % 3];
}

// dx[map, y, x, c] = sum_t w[t, c] * d_t(y, x); the 72 weights of the thread's vector live in registers.  grid = (tiles, maps)
__global__ void __launch_bounds__(256)
conv_out1_dgrad_bf16_kernel(const float* __restrict__ dout, const float* __restrict__ wgt, __nv_bfloat16* __restrict__ dx, int64_t ld, int h, int w,
                            int tiles_x) {
  constexpr int VPP = kO1C / 8, U = kB1TH * kB1TW * VPP / 256;
  __shared__ float s_d[kB1HW];
  const int v = threadIdx.x % VPP, c8 = v * 8;
  float wr[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t) ldg8f(wgt + t * kO1C + c8, wr[t]);
  const int64_t map = blockIdx.y;
  const int y0 = (blockIdx.x / tiles_x) * kB1TH, x0 = (blockIdx.x % tiles_x) * kB1TW;
  out1_load_halo(s_d, dout + map * h * w, y0, x0, h, w);
  __syncthreads();
#pragma unroll 2
  for (int u = 0; u < U; ++u) {
    const int pl = (u * 256 + threadIdx.x) / VPP, ly = pl / kB1TW, lx = pl % kB1TW;
    const int yy = y0 + ly, xx = x0 + lx;
    if (yy >= h || xx >= w) continue;
    float d[9], f[8];
    out1_halo_neighbours(s_d, ly, lx, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += d[t] * wr[t][i];
    }
    *(uint4*)(dx + ((map * h + yy) * w + xx) * ld + c8) = f32_to_bf16x8(f);
  }
}
// dw[t, c] += sum_q x[q, c] * d_t(q);  dbias += sum dout.  Persistent CTAs over (map, tile) units; a thread keeps the 9 x 8
// accumulators of its channel vector in registers across all its units, then warp shuffles + shared memory + one atomic per
// (tap, channel) per CTA.
__global__ void __launch_bounds__(256, 2)
conv_out1_wgrad_bf16_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ x, int64_t ld, float* __restrict__ dw,
                            float* __restrict__ dbias, int h, int w, int tiles_x, int tiles_y, int64_t units) {
  constexpr int VPP = kO1C / 8, U = kB1TH * kB1TW * VPP / 256;
  __shared__ float s_d[kB1HW];
  __shared__ float red[8][9 * kO1C + 1];
  const int v = threadIdx.x % VPP, c8 = v * 8;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
  }
  float sb = 0.f;
  const int hw = h * w, tpm = tiles_x * tiles_y;
  for (int64_t unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const int64_t map = unit / tpm;
    const int tile = (int)(unit % tpm), y0 = (tile / tiles_x) * kB1TH, x0 = (tile % tiles_x) * kB1TW;
    __syncthreads();                                       // the previous unit's halo is no longer read
    out1_load_halo(s_d, dout + map * hw, y0, x0, h, w);
    __syncthreads();
#pragma unroll 2
    for (int u = 0; u < U; u += 2) {                       // two items (16-byte x loads) in flight
      uint4 xv[2];
      bool ok[2];
      int ly[2], lx[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int pl = ((u + k) * 256 + threadIdx.x) / VPP;
        ly[k] = pl / kB1TW; lx[k] = pl % kB1TW;
        ok[k] = y0 + ly[k] < h && x0 + lx[k] < w;
        xv[k] = ld_stream16(x + ((map * h + (ok[k] ? y0 + ly[k] : 0)) * w + (ok[k] ? x0 + lx[k] : 0)) * ld + c8);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (ok[k]) {
          float d[9], f[8];
          out1_halo_neighbours(s_d, ly[k], lx[k], d);
          bf16x8_to_f32(xv[k], f);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[t][i] += d[t] * f[i];
          }
          if (v == 0) sb += d[4];
        }
      }
    }
  }
  // lanes of a warp with the same channel vector: lane % VPP == v  ->  butterfly over the pixel-slot bits
#pragma unroll
  for (int t = 0; t < 9; ++t) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = acc[t][i];
#pragma unroll
      for (int o = VPP; o < 32; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      acc[t][i] = a;
    }
  }
#pragma unroll
  for (int o = VPP; o < 32; o <<= 1) sb += __shfl_xor_sync(0xffffffffu, sb, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane < VPP) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp][t * kO1C + c8 + i] = acc[t][i];
    }
    if (lane == 0) red[warp][9 * kO1C] = sb;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * kO1C + 1; i += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a += red[k][i];
    if (a != 0.f) atomicAdd(i < 9 * kO1C ? dw + i : dbias, a);
  }
}

}  // namespace
}  // namespace svl

using namespace svl;
#define ST (cudaStream_t) stream

// floats of scratch the GroupNorm entry points need: [maps, G, 2] totals + [maps * splits, G, 2] per-CTA partials
extern "C" size_t svl_gn_workspace(int64_t maps, int hw, int C, int G) {
  if (maps <= 0 || G <= 0) return 0;
  return (size_t)(2 * maps * G * (1 + (int64_t)gn_splits(maps, hw, C)));
}

extern "C" int svl_gn_relu_fwd(const void* x, int x_dtype, int64_t ldx, const float* gamma, const float* beta, void* out, int out_dtype,
                               int64_t ldo, const void* res, int res_dtype, int64_t ldres, float* mean, float* rstd, float* ws, int64_t maps,
                               int hw, int C, int G, float eps, int stats_splits, void* stream) {
  SVL_CHECK_ARG(x && gamma && beta && out && mean && rstd && ws, "svl_gn_relu_fwd: null pointer");
  SVL_CHECK_ARG(C % 8 == 0 && C <= 256 && G <= kMaxGroups && C % G == 0 && (C / G) % 8 == 0 && kGnThreads % (C / 8) == 0,
                "svl_gn_relu_fwd: unsupported C=%d G=%d", C, G);
  if (maps == 0) return SVL_OK;
  const int splits = stats_splits > 0 ? stats_splits : gn_splits(maps, hw, C);
  SVL_CHECK_ARG(maps * splits < (1ll << 31), "svl_gn_relu_fwd: grid too large");
  // all-bf16 fast path (every large GroupNorm of the throughput mode); anything else takes the generic kernels
  const int vshift = log2_exact(C / 8);
  const bool fast = x_dtype == SVL_BF16 && out_dtype == SVL_BF16 && (!res || res_dtype == SVL_BF16) && al16(x) && al16(out) && al16(res) &&
                    ldx % 8 == 0 && ldo % 8 == 0 && (!res || ldres % 8 == 0) && vshift >= 0 && maps <= 65535 && C / G == 16 &&
                    (int64_t)hw * (C / 8) < (1ll << 30);
  if (stats_splits > 0) {
    // first stage done by the producing convolution's epilogue (conv_roll.cu)
  } else if (fast)
    gn_stats_bf16_kernel<<<(unsigned)(maps * splits), kGnThreads, 0, ST>>>((const __nv_bfloat16*)x, ldx, ws, hw, C, G, splits);
  else
    gn_stats_kernel<<<(unsigned)(maps * splits), kGnThreads, 0, ST>>>(x, x_dtype, ldx, ws, hw, C, G, splits);
  SVL_LAUNCH_CHECK();
  gn_finalize_kernel<<<(unsigned)((maps * G + 255) / 256), 256, 0, ST>>>(ws, mean, rstd, maps, G, splits, (float)hw * (C / G), eps);
  SVL_LAUNCH_CHECK();
  SVL_CHECK_ARG(C / G == 16, "svl_gn_relu_fwd: the apply kernel is specialised for 16 channels per group");
  if (fast) {
    const dim3 grid((unsigned)cdiv((int64_t)hw * (C / 8), 256 * kGnApplyU), (unsigned)maps);
    gn_apply_bf16_kernel<<<grid, 256, 0, ST>>>((const __nv_bfloat16*)x, ldx, gamma, beta, (__nv_bfloat16*)out, ldo, (const __nv_bfloat16*)res, ldres,
                                              mean, rstd, hw, C, G, vshift);
  } else {
    gn_apply_kernel<<<ew_grid(maps * hw * G, 256), 256, 0, ST>>>(x, x_dtype, ldx, gamma, beta, out, out_dtype, ldo, res, res_dtype, ldres, mean, rstd,
                                                                maps, hw, C, G);
  }
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_gn_relu_bwd(const void* dy, int dy_dtype, int64_t lddy, const void* x, int x_dtype, int64_t ldx, const float* gamma,
                               const float* beta, const float* mean, const float* rstd, void* dx, int dx_dtype, int64_t lddx, float* dgamma,
                               float* dbeta, float* ws, int64_t maps, int hw, int C, int G, void* stream) {
  SVL_CHECK_ARG(dy && x && gamma && beta && mean && rstd && dx && ws, "svl_gn_relu_bwd: null pointer");
  SVL_CHECK_ARG(C / G == 16 && kGnThreads % G == 0, "svl_gn_relu_bwd: kernels are specialised for 16 channels per group");
  SVL_CHECK_ARG(C % 8 == 0 && C <= 256 && G <= kMaxGroups && C % G == 0 && (C / G) % 8 == 0 && kGnThreads % (C / 8) == 0,
                "svl_gn_relu_bwd: unsupported C=%d G=%d", C, G);
  if (maps == 0) return SVL_OK;
  const int splits = gn_splits(maps, hw, C);
  SVL_CHECK_ARG(maps * splits < (1ll << 31), "svl_gn_relu_bwd: grid too large");
  float* part = ws + maps * G * 2;                      // [maps * splits, G, 2] partials behind the [maps, G, 2] totals
  const int vshift = log2_exact(C / 8);
  const bool fast = dy_dtype == SVL_BF16 && x_dtype == SVL_BF16 && dx_dtype == SVL_BF16 && al16(dy) && al16(x) && al16(dx) && lddy % 8 == 0 &&
                    ldx % 8 == 0 && lddx % 8 == 0 && vshift >= 0 && maps <= 65535 && (int64_t)hw * (C / 8) < (1ll << 30);
  if (fast) {
    gn_bwd_stats_bf16_kernel<<<(unsigned)(maps * splits), kGnThreads, 0, ST>>>((const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)x, ldx, gamma,
                                                                              beta, mean, rstd, part, dgamma, dbeta, hw, C, G, splits);
    SVL_LAUNCH_CHECK();
    gn_bwd_finalize_kernel<<<(unsigned)((maps * G + 255) / 256), 256, 0, ST>>>(part, ws, maps, G, splits);
    SVL_LAUNCH_CHECK();
    const dim3 grid((unsigned)cdiv((int64_t)hw * (C / 8), 256 * kGnBwdApplyU), (unsigned)maps);
    gn_bwd_apply_bf16_kernel<<<grid, 256, 0, ST>>>((const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)x, ldx, gamma, beta, mean, rstd, ws,
                                                  (__nv_bfloat16*)dx, lddx, hw, C, G, vshift);
    SVL_LAUNCH_CHECK();
    return SVL_OK;
  }
  gn_bwd_stats_kernel<<<(unsigned)(maps * splits), kGnThreads, 0, ST>>>(dy, dy_dtype, lddy, x, x_dtype, ldx, gamma, beta, mean, rstd, part, dgamma,
                                                                       dbeta, hw, C, G, splits);
  SVL_LAUNCH_CHECK();
  gn_bwd_finalize_kernel<<<(unsigned)((maps * G + 255) / 256), 256, 0, ST>>>(part, ws, maps, G, splits);
  SVL_LAUNCH_CHECK();
  gn_bwd_apply_kernel<<<ew_grid(maps * hw * G, 256), 256, 0, ST>>>(dy, dy_dtype, lddy, x, x_dtype, ldx, gamma, beta, mean, rstd, ws, dx, dx_dtype, lddx,
                                                                  maps, hw, C, G);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_sim_im2col(const float* sim, int64_t ld_sim, void* out, int out_dtype, int64_t ldo, int B, int N, int h, int w, int ks, int kpad,
                              void* stream) {
  SVL_CHECK_ARG(sim && out && kpad % 8 == 0 && kpad >= ks * ks, "svl_sim_im2col: bad arguments");
  sim_im2col_kernel<<<ew_grid((int64_t)B * N * h * w * (kpad / 8)), 256, 0, ST>>>(sim, ld_sim, out, out_dtype, ldo, B, N, h, w, ks, kpad);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_sim_col2im(const void* dcol, int dtype, int64_t ld, void* dsim, int ds_dtype, int64_t ld_ds, int ncols, int B, int N, int h,
                              int w, int ks, void* stream) {
  SVL_CHECK_ARG(dcol && dsim && ncols >= N, "svl_sim_col2im: bad arguments");
  sim_col2im_kernel<<<ew_grid((int64_t)B * h * w * ncols), 256, 0, ST>>>(dcol, dtype, ld, dsim, ds_dtype, ld_ds, ncols, B, N, h, w, ks);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_map_sum(const void* x, int dtype, int64_t ld, float* out, int64_t maps, int hw, int C, float scale, void* stream) {
  SVL_CHECK_ARG(x && out && C <= 128, "svl_map_sum: bad arguments");
  if (maps == 0) return SVL_OK;
  if (dtype == SVL_BF16 && C % 8 == 0 && 256 % (C / 8) == 0 && ld % 8 == 0 && al16(x)) {
    map_sum_bf16_kernel<<<(unsigned)maps, 256, 0, ST>>>((const __nv_bfloat16*)x, ld, out, hw, C, scale);
    SVL_LAUNCH_CHECK();
    return SVL_OK;
  }
  int ry = 1024 / C;
  if (ry > 8) ry = 8;
  map_sum_kernel<<<(unsigned)maps, dim3(C, ry), 0, ST>>>(x, dtype, ld, out, hw, C, scale);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_map_bcast_add(float* x, const void* v, int v_dtype, int64_t ldv, int64_t maps, int hw, int C, float scale, void* stream) {
  SVL_CHECK_ARG(x && v, "svl_map_bcast_add: null pointer");
  if (v_dtype == SVL_F32 && C % 4 == 0 && ldv % 4 == 0 && al16(x) && al16(v)) {
    map_bcast_add_f32x4_kernel<<<ew_grid(maps * hw * (C / 4)), 256, 0, ST>>>(x, (const float*)v, ldv, maps, hw, C, scale);
    SVL_LAUNCH_CHECK();
    return SVL_OK;
  }
  map_bcast_add_kernel<<<ew_grid(maps * hw * C), 256, 0, ST>>>(x, v, v_dtype, ldv, maps, hw, C, scale);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_pool_tokens(const void* x, int x_dtype, int64_t ldx, const float* text, float* tok, int B, int N, int h, int w, int C, int Ct,
                               int pool, void* stream) {
  SVL_CHECK_ARG(x && text && tok && pool > 0 && h >= pool && w >= pool, "svl_pool_tokens: bad arguments");
  if (x_dtype == SVL_BF16 && C % 8 == 0 && Ct % 8 == 0 && ldx % 8 == 0 && al16(x) && al16(text) && al16(tok)) {
    pool_tokens_bf16_kernel<<<ew_grid((int64_t)B * (h / pool) * (w / pool) * N * ((C + Ct) / 8)), 256, 0, ST>>>((const __nv_bfloat16*)x, ldx, text, tok, B, N,
                                                                                                            h, w, C, Ct, pool);
    SVL_LAUNCH_CHECK();
    return SVL_OK;
  }
  pool_tokens_kernel<<<ew_grid((int64_t)B * (h / pool) * (w / pool) * N * (C + Ct)), 256, 0, ST>>>(x, x_dtype, ldx, text, tok, B, N, h, w, C, Ct, pool);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_pool_tokens_bwd(const float* dtok, int64_t ldt, float* dx, int B, int N, int h, int w, int C, int pool, void* stream) {
  SVL_CHECK_ARG(dtok && dx, "svl_pool_tokens_bwd: null pointer");
  SVL_CHECK_ARG(C % 4 == 0 && ldt % 4 == 0, "svl_pool_tokens_bwd: C and ldt must be multiples of 4");
  pool_tokens_bwd_kernel<<<ew_grid((int64_t)B * N * h * w * (C / 4)), 256, 0, ST>>>(dtok, ldt, dx, B, N, h, w, C, pool);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_pool_tokens_bwd_from(const void* src, int src_dtype, int64_t lds, const float* dtok, int64_t ldt, float* dx, int B, int N, int h,
                                       int w, int C, int pool, void* stream) {
  SVL_CHECK_ARG(src && dtok && dx && pool > 0, "svl_pool_tokens_bwd_from: bad arguments");
  SVL_CHECK_ARG((src_dtype == SVL_BF16 || src_dtype == SVL_F32) && C % 8 == 0 && ldt % 4 == 0 && lds % 8 == 0 && al16(src) && al16(dtok) && al16(dx),
                "svl_pool_tokens_bwd_from: needs bf16 / f32 rows of 8-channel vectors, 16-byte aligned");
  const int grid = ew_grid((int64_t)B * N * h * w * (C / 8));
  if (src_dtype == SVL_BF16) pool_tokens_bwd_from_kernel<true><<<grid, 256, 0, ST>>>(src, lds, dtok, ldt, dx, B, N, h, w, C, pool);
  else pool_tokens_bwd_from_kernel<false><<<grid, 256, 0, ST>>>(src, lds, dtok, ldt, dx, B, N, h, w, C, pool);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_unpool_add(const void* x, int x_dtype, int64_t ldx, const float* tok, int64_t ldt, void* out, int out_dtype, int64_t ldo, int B,
                              int N, int h, int w, int C, int hp, int wp, void* stream) {
  SVL_CHECK_ARG(x && tok && out && C % 8 == 0, "svl_unpool_add: bad arguments");
  if (x_dtype == SVL_BF16 && out_dtype == SVL_BF16 && ldx % 8 == 0 && ldo % 8 == 0 && ldt % 4 == 0 && al16(x) && al16(out) && al16(tok) &&
      (int64_t)B * N <= 65535 && (int64_t)h * w * (C / 8) < (1ll << 30)) {
    const dim3 grid((unsigned)cdiv((int64_t)h * w * (C / 8), 512), (unsigned)(B * N));
    unpool_add_bf16_kernel<<<grid, 256, 0, ST>>>((const __nv_bfloat16*)x, ldx, tok, ldt, (__nv_bfloat16*)out, ldo, N, h, w, C, hp, wp);
    SVL_LAUNCH_CHECK();
    return SVL_OK;
  }
  unpool_add_kernel<<<ew_grid((int64_t)B * N * h * w * (C / 8)), 256, 0, ST>>>(x, x_dtype, ldx, tok, ldt, out, out_dtype, ldo, B, N, h, w, C, hp, wp);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_unpool_bwd(const void* dout, int dtype, int64_t ld, float* dtok, int64_t ldt, int B, int N, int h, int w, int C, int hp, int wp,
                              void* stream) {
  SVL_CHECK_ARG(dout && dtok, "svl_unpool_bwd: null pointer");
  if ((dtype == SVL_BF16 || dtype == SVL_F32) && ld % 8 == 0 && ldt % 8 == 0 && C % 8 == 0 && al16(dout) && al16(dtok)) {
    if (dtype == SVL_BF16)
      unpool_bwd_vec_kernel<true><<<ew_grid((int64_t)B * hp * wp * N * (ldt / 8)), 256, 0, ST>>>(dout, ld, dtok, ldt, B, N, h, w, C, hp, wp);
    else
      unpool_bwd_vec_kernel<false><<<ew_grid((int64_t)B * hp * wp * N * (ldt / 8)), 256, 0, ST>>>(dout, ld, dtok, ldt, B, N, h, w, C, hp, wp);
    SVL_LAUNCH_CHECK();
    return SVL_OK;
  }
  unpool_bwd_kernel<<<ew_grid((int64_t)B * hp * wp * N * ldt), 256, 0, ST>>>(dout, dtype, ld, dtok, ldt, B, N, h, w, C, hp, wp);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_skip_fill(const void* skip, int s_dtype, int64_t lds, void* cat, int c_dtype, int64_t ldc, int c0, int B, int N, int h, int w,
                             int Cs, int H2, int W2, void* stream) {
  SVL_CHECK_ARG(skip && cat, "svl_skip_fill: null pointer");
  if (s_dtype == SVL_F32 && c_dtype == SVL_BF16 && Cs % 8 == 0 && c0 % 8 == 0 && ldc % 8 == 0 && lds % 4 == 0 && al16(skip) && al16(cat)) {
    skip_fill_bf16_kernel<<<ew_grid((int64_t)B * H2 * W2 * (Cs / 8)), 256, 0, ST>>>((const float*)skip, lds, (__nv_bfloat16*)cat, ldc, c0, B, N, h, w, Cs, H2,
                                                                                  W2);
    SVL_LAUNCH_CHECK();
    return SVL_OK;
  }
  skip_fill_kernel<<<ew_grid((int64_t)B * H2 * W2 * Cs), 256, 0, ST>>>(skip, s_dtype, lds, cat, c_dtype, ldc, c0, B, N, h, w, Cs, H2, W2);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_class_sum(const void* x, int dtype, int64_t ld, int c0, float* out, int B, int N, int64_t P, int Cs, void* stream) {
  SVL_CHECK_ARG(x && out && Cs % 8 == 0, "svl_class_sum: bad arguments");
  if (dtype == SVL_BF16 && ld % 8 == 0 && c0 % 8 == 0 && al16(x) && al16(out)) {
    class_sum_bf16_kernel<<<ew_grid((int64_t)B * P * (Cs / 8)), 256, 0, ST>>>((const __nv_bfloat16*)x, ld, c0, out, B, N, P, Cs);
    SVL_LAUNCH_CHECK();
    return SVL_OK;
  }
  class_sum_kernel<<<ew_grid((int64_t)B * P * (Cs / 8)), 256, 0, ST>>>(x, dtype, ld, c0, out, B, N, P, Cs);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_skip_grad(const void* dcat, int d_dtype, int64_t ldd, int c0, const void* skip, int s_dtype, int64_t lds, void* dskip,
                             int o_dtype, int64_t ldo, int B, int N, int h, int w, int Cs, int H2, int W2, void* stream) {
  SVL_CHECK_ARG(dcat && skip && dskip, "svl_skip_grad: null pointer");
  SVL_CHECK_ARG(Cs % 8 == 0, "svl_skip_grad: Cs must be a multiple of 8");
  skip_grad_kernel<<<ew_grid((int64_t)B * h * w * (Cs / 8), 128), 128, 0, ST>>>(dcat, d_dtype, ldd, c0, skip, s_dtype, lds, dskip, o_dtype, ldo, B, N, h,
                                                                               w, Cs, H2, W2);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_conv_out1_fwd(const void* x, int dtype, int64_t ld, const float* wgt, const float* bias, float* out, int64_t maps, int h, int w,
                                 int C, void* stream) {
  SVL_CHECK_ARG(x && wgt && bias && out && C % 8 == 0, "svl_conv_out1_fwd: bad arguments");
  const int tiles_x = (w + kO1TW - 1) / kO1TW, tiles_y = (h + kO1TH - 1) / kO1TH;
  const size_t smem = (size_t)(9 * C + 9 * kO1HW) * sizeof(float);
  SVL_CHECK_ARG(maps * tiles_x * tiles_y < (1ll << 31), "svl_conv_out1_fwd: grid too large");
  if (dtype == SVL_BF16 && C == kO1C && ld % 8 == 0 && al16(x) && al16(wgt))
    conv_out1_fwd_bf16_kernel<<<(unsigned)(maps * tiles_x * tiles_y), 256, 0, ST>>>((const __nv_bfloat16*)x, ld, wgt, bias, out, h, w, tiles_x, tiles_y);
  else
    conv_out1_fwd_kernel<<<(unsigned)(maps * tiles_x * tiles_y), 256, smem, ST>>>(x, dtype, ld, wgt, bias, out, h, w, C, tiles_x, tiles_y);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_conv_out1_bwd(const float* dout, const void* x, int x_dtype, int64_t ldx, const float* wgt, void* dx, int dx_dtype, int64_t lddx,
                                 float* dw, float* dbias, int64_t maps, int h, int w, int C, void* stream) {
  SVL_CHECK_ARG(dout && x && wgt && dx && dw && dbias && C % 8 == 0 && 9 * C <= 1024, "svl_conv_out1_bwd: bad arguments");
  if (x_dtype == SVL_BF16 && dx_dtype == SVL_BF16 && C == kO1C && ldx % 8 == 0 && lddx % 8 == 0 && al16(x) && al16(dx) && al16(wgt) && maps <= 65535 &&
      (int64_t)h * w * (C / 8) < (1ll << 30)) {
    const int btx = (w + kB1TW - 1) / kB1TW, bty = (h + kB1TH - 1) / kB1TH;
    const dim3 grid((unsigned)(btx * bty), (unsigned)maps);
    conv_out1_dgrad_bf16_kernel<<<grid, 256, 0, ST>>>(dout, wgt, (__nv_bfloat16*)dx, lddx, h, w, btx);
    SVL_LAUNCH_CHECK();
    const int64_t units = maps * btx * bty;
    conv_out1_wgrad_bf16_kernel<<<(unsigned)(units < 148 * 2 ? units : 148 * 2), 256, 0, ST>>>(dout, (const __nv_bfloat16*)x, ldx, dw, dbias, h, w, btx, bty,
                                                                                          units);
    SVL_LAUNCH_CHECK();
    return SVL_OK;
  }
  conv_out1_dgrad_kernel<<<ew_grid(maps * h * w * (C / 8)), 256, 9 * C * sizeof(float), ST>>>(dout, wgt, dx, dx_dtype, lddx, maps, h, w, C);
  SVL_LAUNCH_CHECK();
  const int tiles_x = (w + kO1TW - 1) / kO1TW, tiles_y = (h + kO1TH - 1) / kO1TH;
  const int64_t num_tiles = maps * tiles_x * tiles_y;
  const int grid = (int)(num_tiles < 148 * 4 ? num_tiles : 148 * 4);
  conv_out1_wgrad_kernel<<<grid, 256, 0, ST>>>(dout, x, x_dtype, ldx, dw, dbias, h, w, C, tiles_x, tiles_y, num_tiles);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
