// HBM-bound kernels of the conv encoder of the Cityscapes skr04 model: mmseg ResNetV1c(depth=101, num_stages=1) = deep stem + layer1
// with (Sync)BatchNorm (configs/_base_/models/vlm-vlg-aspp-s2p4-skr04-ftap-mcvitb.py:50-60; model/vlm.py:50-52,120-121).
// The convolutions themselves run on the tcgen05 contraction engine (gemm.cu / wgrad.cu: 1x1 as plain GEMMs, 3x3 as implicit GEMMs; the
// 3 -> 32 stride-2 stem convolution through the im2col below, K = 27 padded to 32).  Activations are NHWC [rows = B*H*W, C].
//
// BatchNorm in training mode needs per-channel statistics over ALL rows of ALL ranks (SyncBN): the statistics kernels reduce in two
// fixed-order stages (per-CTA partials, then one thread per channel) into a small [2][C] buffer that the host side all-reduces over
// NCCL before the finalize / apply kernels run -- the one collective inside the forward pass of this path.
#include "common.cuh"

namespace svl {
namespace {

constexpr int kBnThreads = 256;

inline int ew_grid(int64_t total, int block = 256) {
  int64_t b = cdiv(total, block);
  int64_t cap = 148 * 32;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// ---------------------------------------------------------------------------------------------- stem im2col (3x3, stride 2, pad 1, 3 channels)
// out[(b, oy, ox), c*9 + ky*3 + kx] = img[b, c, 2*oy - 1 + ky, 2*ox - 1 + kx] (zero outside), columns 27..31 zero: the weight [32, 3, 3, 3]
// flattened is the matching [32, 27] operand.  One thread = one output pixel (four 8-column vectors).
__global__ void stem_im2col_kernel(const float* __restrict__ img, void* __restrict__ out, int out_dtype, int64_t ldo, int B, int H, int W, int Ho,
                                   int Wo) {
  const int64_t total = (int64_t)B * Ho * Wo;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % Wo);
    const int oy = (int)((idx / Wo) % Ho);
    const int b = (int)(idx / ((int64_t)Wo * Ho));
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int y = 2 * oy - 1 + ky;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int x = 2 * ox - 1 + kx;
          if (y >= 0 && y < H && x >= 0 && x < W) v[c * 9 + ky * 3 + kx] = __ldg(img + (((int64_t)b * 3 + c) * H + y) * W + x);
        }
      }
#pragma unroll
    for (int g = 0; g < 4; ++g) st8(out, out_dtype, idx * ldo + g * 8, ldo / 2, 8, v + g * 8);
  }
}

// ---------------------------------------------------------------------------------------------- max-pool 3x3, stride 2, pad 1 (NHWC)
// The window position of the maximum (first maximum in (ky, kx) scan order, like ATen) is saved as one byte per element; the backward
// pass is a gather over the <= 4 windows that contain an input pixel, so it needs no atomics.
__global__ void maxpool_fwd_kernel(const void* __restrict__ x, int dtype, int64_t ldx, void* __restrict__ out, int out_dtype, int64_t ldo,
                                   uint8_t* __restrict__ widx, int B, int H, int W, int C, int Ho, int Wo) {
  const int vpp = C / 8;
  const int64_t total = (int64_t)B * Ho * Wo * vpp;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % vpp) * 8;
    const int64_t pix = idx / vpp;
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((int64_t)Wo * Ho));
    float best[8];
    int bi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { best[i] = -INFINITY; bi[i] = 0; }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int y = 2 * oy - 1 + ky;
      if (y < 0 || y >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = 2 * ox - 1 + kx;
        if (xx < 0 || xx >= W) continue;
        float f[8];
        ld8(x, dtype, (((int64_t)b * H + y) * W + xx) * ldx + c8, ldx / 2, 8, f);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (f[i] > best[i]) { best[i] = f[i]; bi[i] = ky * 3 + kx; }
      }
    }
    st8(out, out_dtype, pix * ldo + c8, ldo / 2, 8, best);
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { lo |= (uint32_t)bi[i] << (8 * i); hi |= (uint32_t)bi[4 + i] << (8 * i); }
    *(uint2*)(widx + pix * C + c8) = make_uint2(lo, hi);
  }
}
__global__ void maxpool_bwd_kernel(const void* __restrict__ dy, int dy_dtype, int64_t lddy, const uint8_t* __restrict__ widx, void* __restrict__ dx,
                                   int dx_dtype, int64_t lddx, int B, int H, int W, int C, int Ho, int Wo) {
  const int vpp = C / 8;
  const int64_t total = (int64_t)B * H * W * vpp;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % vpp) * 8;
    const int64_t pix = idx / vpp;
    const int xx = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((int64_t)W * H));
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    // windows oy with 2*oy - 1 <= y <= 2*oy + 1
    const int oy0 = y >> 1, oy1 = (y + 1) >> 1, ox0 = xx >> 1, ox1 = (xx + 1) >> 1;
    for (int oy = oy0; oy <= oy1; ++oy) {
      if (oy >= Ho) continue;
      const int ky = y - (2 * oy - 1);
      for (int ox = ox0; ox <= ox1; ++ox) {
        if (ox >= Wo) continue;
        const int kx = xx - (2 * ox - 1);
        const int64_t op = ((int64_t)b * Ho + oy) * Wo + ox;
        const uint2 w = *(const uint2*)(widx + op * C + c8);
        float g[8];
        ld8(dy, dy_dtype, op * lddy + c8, lddy / 2, 8, g);
        const int me = ky * 3 + kx;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int wi = (int)(((i < 4 ? w.x : w.y) >> (8 * (i & 3))) & 0xffu);
          if (wi == me) acc[i] += g[i];
        }
      }
    }
    st8(dx, dx_dtype, pix * lddx + c8, lddx / 2, 8, acc);
  }
}

// ---------------------------------------------------------------------------------------------- BatchNorm statistics (two fixed-order stages)
// thread = (row slot, 8-channel vector); the 16 per-thread partials are combined through a padded shared-memory table in slot order,
// one CTA writes its [2][C] partial, bn_reduce_kernel sums the CTAs' partials in CTA order: no floating-point atomics anywhere.
//   forward : a = x,  b = x * x
//   backward: a = g,  b = g * xhat     with g = dy * [y > 0] (y = the saved block output, NULL: no ReLU) and xhat = (x - mean) * rstd
template <bool kBackward>
__global__ void __launch_bounds__(kBnThreads)
bn_partial_kernel(const void* __restrict__ x, int x_dtype, int64_t ldx, const void* __restrict__ dy, int dy_dtype, int64_t lddy,
                  const void* __restrict__ y, int y_dtype, int64_t ldy, const float* __restrict__ mean, const float* __restrict__ rstd,
                  int64_t rows, int C, float* __restrict__ part) {
  __shared__ float tab[16][kBnThreads + 1];
  const int vpp = C / 8, ppi = kBnThreads / vpp;
  const int v = threadIdx.x % vpp, c8 = v * 8;
  const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
  float a[8], b[8], mu[8], rs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = b[i] = 0.f; mu[i] = 0.f; rs[i] = 1.f; }
  if (kBackward) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { mu[i] = mean[c8 + i]; rs[i] = rstd[c8 + i]; }
  }
  for (int64_t r = r0 + threadIdx.x / vpp; r < r1; r += ppi) {
    float xv[8];
    ld8(x, x_dtype, r * ldx + c8, ldx / 2, 8, xv);
    if (kBackward) {
      float g[8];
      ld8(dy, dy_dtype, r * lddy + c8, lddy / 2, 8, g);
      if (y) {
        float yv[8];
        ld8(y, y_dtype, r * ldy + c8, ldy / 2, 8, yv);
#pragma unroll
        for (int i = 0; i < 8; ++i) if (!(yv[i] > 0.f)) g[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] += g[i]; b[i] += g[i] * ((xv[i] - mu[i]) * rs[i]); }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] += xv[i]; b[i] += xv[i] * xv[i]; }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { tab[i][threadIdx.x] = a[i]; tab[8 + i][threadIdx.x] = b[i]; }
  __syncthreads();
  if (threadIdx.x < C) {
    const int cv = threadIdx.x / 8, ci = threadIdx.x % 8;
    float sa = 0.f, sb = 0.f;
    for (int ps = 0; ps < ppi; ++ps) { sa += tab[ci][ps * vpp + cv]; sb += tab[8 + ci][ps * vpp + cv]; }
    part[((int64_t)blockIdx.x * 2) * C + threadIdx.x] = sa;
    part[((int64_t)blockIdx.x * 2 + 1) * C + threadIdx.x] = sb;
  }
}
__global__ void bn_reduce_kernel(const float* __restrict__ part, int nblk, int C, float* __restrict__ sums /* [2][C] */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * C) return;
  float s = 0.f;
  for (int k = 0; k < nblk; ++k) s += part[(int64_t)k * 2 * C + i];
  sums[i] = s;
}
// sums = (global) [sum x | sum x^2] over `count` rows -> batch mean / rstd (biased variance) and the running statistics
// (momentum m: running = (1 - m) * running + m * batch, with the UNBIASED batch variance, as nn.BatchNorm2d / SyncBatchNorm do)
__global__ void bn_finalize_kernel(const float* __restrict__ sums, float count, float eps, float momentum, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ running_mean, float* __restrict__ running_var, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mu = sums[c] / count;
  const float var = fmaxf(sums[C + c] / count - mu * mu, 0.f);
  mean[c] = mu;
  rstd[c] = rsqrtf(var + eps);
  if (running_mean) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * (count > 1.f ? count / (count - 1.f) : 1.f);
  }
}
// out = [relu]( (x - mean) * rstd * gamma + beta [+ res] )
__global__ void bn_apply_kernel(const void* __restrict__ x, int x_dtype, int64_t ldx, const float* __restrict__ mean, const float* __restrict__ rstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta, const void* __restrict__ res, int res_dtype,
                                int64_t ldres, void* __restrict__ out, int out_dtype, int64_t ldo, int relu, int64_t rows, int C) {
  const int vpp = C / 8;
  const int64_t total = rows * vpp;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % vpp) * 8;
    const int64_t r = idx / vpp;
    float f[8];
    ld8(x, x_dtype, r * ldx + c8, ldx / 2, 8, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = (f[i] - __ldg(mean + c8 + i)) * __ldg(rstd + c8 + i) * __ldg(gamma + c8 + i) + __ldg(beta + c8 + i);
    if (res) {
      float q[8];
      ld8(res, res_dtype, r * ldres + c8, ldres / 2, 8, q);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += q[i];
    }
    if (relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
    }
    st8(out, out_dtype, r * ldo + c8, ldo / 2, 8, f);
  }
}
// dx = gamma * rstd * (g - S1 / N - xhat * S2 / N) with the GLOBAL sums S = [sum g | sum g * xhat] over N rows; dres = g (optional)
__global__ void bn_bwd_apply_kernel(const void* __restrict__ dy, int dy_dtype, int64_t lddy, const void* __restrict__ x, int x_dtype, int64_t ldx,
                                    const void* __restrict__ y, int y_dtype, int64_t ldy, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ sums, float count,
                                    void* __restrict__ dx, int dx_dtype, int64_t lddx, void* __restrict__ dres, int dres_dtype, int64_t lddres,
                                    int64_t rows, int C) {
  const int vpp = C / 8;
  const int64_t total = rows * vpp;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % vpp) * 8;
    const int64_t r = idx / vpp;
    float xv[8], g[8];
    ld8(x, x_dtype, r * ldx + c8, ldx / 2, 8, xv);
    ld8(dy, dy_dtype, r * lddy + c8, lddy / 2, 8, g);
    if (y) {
      float yv[8];
      ld8(y, y_dtype, r * ldy + c8, ldy / 2, 8, yv);
#pragma unroll
      for (int i = 0; i < 8; ++i) if (!(yv[i] > 0.f)) g[i] = 0.f;
    }
    if (dres) st8(dres, dres_dtype, r * lddres + c8, lddres / 2, 8, g);
    float d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float rs = __ldg(rstd + c8 + i);
      const float xh = (xv[i] - __ldg(mean + c8 + i)) * rs;
      d[i] = __ldg(gamma + c8 + i) * rs * (g[i] - __ldg(sums + c8 + i) / count - xh * __ldg(sums + C + c8 + i) / count);
    }
    st8(dx, dx_dtype, r * lddx + c8, lddx / 2, 8, d);
  }
}

inline int bn_blocks(int64_t rows, int C) {
  const int ppi = kBnThreads / (C / 8);
  int64_t nb = cdiv(rows, (int64_t)ppi * 16);          // >= ~16 row iterations per thread
  if (nb > 148 * 4) nb = 148 * 4;
  return (int)(nb < 1 ? 1 : nb);
}
inline bool bn_shape_ok(int C) { return C % 8 == 0 && C <= 256 && kBnThreads % (C / 8) == 0; }

}  // namespace
}  // namespace svl

using namespace svl;
#define ST (cudaStream_t) stream

extern "C" int svl_stem_im2col(const float* img, void* out, int out_dtype, int64_t ldo, int B, int H, int W, int Ho, int Wo, void* stream) {
  SVL_CHECK_ARG(img && out && out_dtype != SVL_F32 && ldo >= 32, "svl_stem_im2col: bad arguments");
  SVL_CHECK_ARG(Ho == (H - 1) / 2 + 1 && Wo == (W - 1) / 2 + 1, "svl_stem_im2col: output %dx%d is not the 3x3 / stride 2 / pad 1 grid of %dx%d", Ho, Wo, H, W);
  stem_im2col_kernel<<<ew_grid((int64_t)B * Ho * Wo), 256, 0, ST>>>(img, out, out_dtype, ldo, B, H, W, Ho, Wo);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_maxpool3s2_fwd(const void* x, int dtype, int64_t ldx, void* out, int out_dtype, int64_t ldo, uint8_t* widx, int B, int H, int W,
                                  int C, int Ho, int Wo, void* stream) {
  SVL_CHECK_ARG(x && out && widx && C % 8 == 0, "svl_maxpool3s2_fwd: bad arguments");
  SVL_CHECK_ARG(Ho == (H - 1) / 2 + 1 && Wo == (W - 1) / 2 + 1, "svl_maxpool3s2_fwd: bad output size");
  maxpool_fwd_kernel<<<ew_grid((int64_t)B * Ho * Wo * (C / 8)), 256, 0, ST>>>(x, dtype, ldx, out, out_dtype, ldo, widx, B, H, W, C, Ho, Wo);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_maxpool3s2_bwd(const void* dy, int dy_dtype, int64_t lddy, const uint8_t* widx, void* dx, int dx_dtype, int64_t lddx, int B, int H,
                                  int W, int C, int Ho, int Wo, void* stream) {
  SVL_CHECK_ARG(dy && dx && widx && C % 8 == 0, "svl_maxpool3s2_bwd: bad arguments");
  maxpool_bwd_kernel<<<ew_grid((int64_t)B * H * W * (C / 8)), 256, 0, ST>>>(dy, dy_dtype, lddy, widx, dx, dx_dtype, lddx, B, H, W, C, Ho, Wo);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" size_t svl_bn_workspace(int64_t rows, int C) {
  if (rows <= 0 || !bn_shape_ok(C)) return 0;
  return (size_t)bn_blocks(rows, C) * 2 * C;
}

extern "C" int svl_bn_stats(const void* x, int x_dtype, int64_t ldx, int64_t rows, int C, float* ws, float* sums, void* stream) {
  SVL_CHECK_ARG(x && ws && sums && rows > 0 && bn_shape_ok(C), "svl_bn_stats: bad arguments (C=%d)", C);
  const int nb = bn_blocks(rows, C);
  bn_partial_kernel<false><<<nb, kBnThreads, 0, ST>>>(x, x_dtype, ldx, nullptr, 0, 0, nullptr, 0, 0, nullptr, nullptr, rows, C, ws);
  SVL_LAUNCH_CHECK();
  bn_reduce_kernel<<<(2 * C + 255) / 256, 256, 0, ST>>>(ws, nb, C, sums);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_bn_finalize(const float* sums, float count, float eps, float momentum, float* mean, float* rstd, float* running_mean,
                               float* running_var, int C, void* stream) {
  SVL_CHECK_ARG(sums && mean && rstd && count > 0 && (!running_mean == !running_var), "svl_bn_finalize: bad arguments");
  bn_finalize_kernel<<<(C + 255) / 256, 256, 0, ST>>>(sums, count, eps, momentum, mean, rstd, running_mean, running_var, C);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_bn_apply(const void* x, int x_dtype, int64_t ldx, const float* mean, const float* rstd, const float* gamma, const float* beta,
                            const void* res, int res_dtype, int64_t ldres, void* out, int out_dtype, int64_t ldo, int relu, int64_t rows, int C,
                            void* stream) {
  SVL_CHECK_ARG(x && mean && rstd && gamma && beta && out && C % 8 == 0, "svl_bn_apply: bad arguments");
  if (rows == 0) return SVL_OK;
  bn_apply_kernel<<<ew_grid(rows * (C / 8)), 256, 0, ST>>>(x, x_dtype, ldx, mean, rstd, gamma, beta, res, res_dtype, ldres, out, out_dtype, ldo, relu,
                                                          rows, C);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_bn_bwd_stats(const void* dy, int dy_dtype, int64_t lddy, const void* x, int x_dtype, int64_t ldx, const void* y, int y_dtype,
                                int64_t ldy, const float* mean, const float* rstd, int64_t rows, int C, float* ws, float* sums, void* stream) {
  SVL_CHECK_ARG(dy && x && mean && rstd && ws && sums && rows > 0 && bn_shape_ok(C), "svl_bn_bwd_stats: bad arguments (C=%d)", C);
  const int nb = bn_blocks(rows, C);
  bn_partial_kernel<true><<<nb, kBnThreads, 0, ST>>>(x, x_dtype, ldx, dy, dy_dtype, lddy, y, y_dtype, ldy, mean, rstd, rows, C, ws);
  SVL_LAUNCH_CHECK();
  bn_reduce_kernel<<<(2 * C + 255) / 256, 256, 0, ST>>>(ws, nb, C, sums);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_bn_bwd_apply(const void* dy, int dy_dtype, int64_t lddy, const void* x, int x_dtype, int64_t ldx, const void* y, int y_dtype,
                                int64_t ldy, const float* mean, const float* rstd, const float* gamma, const float* sums, float count, void* dx,
                                int dx_dtype, int64_t lddx, void* dres, int dres_dtype, int64_t lddres, int64_t rows, int C, void* stream) {
  SVL_CHECK_ARG(dy && x && mean && rstd && gamma && sums && dx && count > 0 && C % 8 == 0, "svl_bn_bwd_apply: bad arguments");
  if (rows == 0) return SVL_OK;
  bn_bwd_apply_kernel<<<ew_grid(rows * (C / 8)), 256, 0, ST>>>(dy, dy_dtype, lddy, x, x_dtype, ldx, y, y_dtype, ldy, mean, rstd, gamma, sums, count, dx,
                                                              dx_dtype, lddx, dres, dres_dtype, lddres, rows, C);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
