// Strong / scale augmentations of the reference's SemiDataset on uint8 images resident on the device
// (third_party/unimatch/dataset/transform.py:43-64 resize / blur; third_party/unimatch/dataset/semi.py:63-93 ColorJitter, RandomGrayscale,
// GaussianBlur): integer-exact counterparts of the Pillow / torchvision code paths the reference calls -- two-pass fixed-point BILINEAR
// resampling with host-built 22-bit coefficient tables (Resample.c), NEAREST through host-built index tables (ImagingScaleAffine),
// ImageEnhance blends in single precision (Blend.c), rgb2l, rgb2hsv / hsv2rgb (Convert.c) and the 3 x extended box blur of GaussianBlur
// (BoxBlur.c).  All kernels are HBM-bound byte work: one thread per output pixel, interleaved HWC uint8.
#include "common.cuh"

namespace svl {
namespace {

inline int ew_grid(int64_t total, int block = 256) {
  int64_t b = cdiv(total, block);
  int64_t cap = 148 * 32;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// one pass of ImagingResample: out[o, .] = clip8((2^21 + sum_k src[first_o + k, .] * kk[o, k]) >> 22) along x (axis = 1) or y (axis = 0)
__global__ void resample_pass_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int* __restrict__ kk,
                                     const int* __restrict__ bounds, int ksize, int h, int w, int c, int out_size, int axis) {
  const int oh = axis == 0 ? out_size : h, ow = axis == 1 ? out_size : w;
  const int64_t total = (int64_t)oh * ow * c;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(idx % c);
    const int x = (int)((idx / c) % ow), y = (int)(idx / ((int64_t)c * ow));
    const int o = axis == 1 ? x : y;
    const int first = bounds[2 * o], n = bounds[2 * o + 1];
    const int* k = kk + (int64_t)o * ksize;
    int acc = 1 << 21;
    for (int t = 0; t < n; ++t) {
      const int sy = axis == 0 ? first + t : y, sx = axis == 1 ? first + t : x;
      acc += (int)src[((int64_t)sy * w + sx) * c + ch] * k[t];
    }
    acc >>= 22;
    dst[idx] = (uint8_t)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
  }
}

__global__ void gather_nearest_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int* __restrict__ ys,
                                      const int* __restrict__ xs, int w, int oh, int ow) {
  const int64_t total = (int64_t)oh * ow;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x)
    dst[idx] = src[(int64_t)ys[idx / ow] * w + xs[idx % ow]];
}

__device__ __forceinline__ int rgb2l(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// sum of the gray values (uint64) -> the mean ImageEnhance.Contrast blends towards
__global__ void gray_sum_kernel(const uint8_t* __restrict__ img, int64_t npix, unsigned long long* __restrict__ out) {
  unsigned long long s = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x)
    s += (unsigned long long)rgb2l(img[3 * i], img[3 * i + 1], img[3 * i + 2]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);                 // integer sum: the order does not matter
}

// NOTE: every floating-point expression below uses the explicitly rounded intrinsics (__fmul_rn, __fadd_rn, ...): nvcc would otherwise
// contract a * b + c into one FMA, which rounds once where the C code Pillow is compiled from rounds twice.
__device__ __forceinline__ uint8_t blend1(int a, int b, float alpha, bool inside) {
  const float t = __fadd_rn((float)a, __fmul_rn(alpha, (float)(b - a)));
  if (inside) return (uint8_t)(int)t;
  return t <= 0.f ? 0 : (t >= 255.f ? 255 : (uint8_t)(int)t);
}

// mode 0 brightness, 1 contrast (towards the gray mean), 2 saturation (towards the pixel's gray value), 3 grayscale (3 equal channels),
// 4 hue (H channel of the HSV form shifted by `shift` with uint8 wrap-around)
__global__ void color_op_kernel(uint8_t* __restrict__ img, int64_t npix, int mode, float factor, const unsigned long long* __restrict__ gray_sum,
                                int shift) {
  const bool inside = factor >= 0.f && factor <= 1.f;
  int mean = 0;
  if (mode == 1) mean = (int)__dadd_rn(__ddiv_rn((double)(*gray_sum), (double)npix), 0.5);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
    int r = img[3 * i], g = img[3 * i + 1], b = img[3 * i + 2];
    if (mode == 0) {
      r = blend1(0, r, factor, inside); g = blend1(0, g, factor, inside); b = blend1(0, b, factor, inside);
    } else if (mode == 1) {
      r = blend1(mean, r, factor, inside); g = blend1(mean, g, factor, inside); b = blend1(mean, b, factor, inside);
    } else if (mode == 2) {
      const int l = rgb2l(r, g, b);
      r = blend1(l, r, factor, inside); g = blend1(l, g, factor, inside); b = blend1(l, b, factor, inside);
    } else if (mode == 3) {
      r = g = b = rgb2l(r, g, b);
    } else {
      // Convert.c rgb2hsv_row: float variables, double literals (the hue expression and the fmod are double, rounded to float when stored)
      const int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
      int uh = 0, us = 0;
      const int uv = maxc;
      if (minc != maxc) {
        const float cr = (float)(maxc - minc);
        const float s = __fdiv_rn(cr, (float)maxc);
        const float rc = __fdiv_rn((float)(maxc - r), cr), gc = __fdiv_rn((float)(maxc - g), cr), bc = __fdiv_rn((float)(maxc - b), cr);
        float h;
        if (r == maxc) h = (float)__dsub_rn((double)bc, (double)gc);
        else if (g == maxc) h = (float)__dsub_rn(__dadd_rn(2.0, (double)rc), (double)bc);
        else h = (float)__dsub_rn(__dadd_rn(4.0, (double)gc), (double)rc);
        h = (float)fmod(__dadd_rn(__ddiv_rn((double)h, 6.0), 1.0), 1.0);
        uh = (int)__dmul_rn((double)h, 255.0);
        us = (int)__dmul_rn((double)s, 255.0);
        uh = uh < 0 ? 0 : (uh > 255 ? 255 : uh);
        us = us < 0 ? 0 : (us > 255 ? 255 : us);
      }
      uh = (uh + shift) & 255;
      // Convert.c hsv2rgb: single precision, p / q / t rounded half up
      if (us == 0) {
        r = g = b = uv;
      } else {
        const float fs = __fdiv_rn((float)us, 255.0f);
        const float hf = __fdiv_rn(__fmul_rn((float)uh, 6.0f), 255.0f);
        const int i6 = (int)floorf(hf);
        const float f = __fsub_rn(hf, (float)i6);
        const float vf = (float)uv;
        const int p = (int)floorf(__fadd_rn(__fmul_rn(vf, __fsub_rn(1.0f, fs)), 0.5f)),
                  q = (int)floorf(__fadd_rn(__fmul_rn(vf, __fsub_rn(1.0f, __fmul_rn(fs, f))), 0.5f)),
                  t = (int)floorf(__fadd_rn(__fmul_rn(vf, __fsub_rn(1.0f, __fmul_rn(fs, __fsub_rn(1.0f, f)))), 0.5f));
        switch (i6 % 6) {
          case 0: r = uv; g = t; b = p; break;
          case 1: r = q; g = uv; b = p; break;
          case 2: r = p; g = uv; b = t; break;
          case 3: r = p; g = q; b = uv; break;
          case 4: r = t; g = p; b = uv; break;
          default: r = uv; g = p; b = q; break;
        }
        r = min(max(r, 0), 255); g = min(max(g, 0), 255); b = min(max(b, 0), 255);
      }
    }
    img[3 * i] = (uint8_t)r; img[3 * i + 1] = (uint8_t)g; img[3 * i + 2] = (uint8_t)b;
  }
}

// one extended box blur along x (axis = 1) or y (axis = 0) with edge replication, uint32 arithmetic (BoxBlur.c ImagingLineBoxBlur8):
//   out = ((sum_{|d| <= radius} src[i + d]) * ww + (src[i - radius - 1] + src[i + radius + 1]) * fw + 2^23) >> 24
__global__ void box_blur_pass_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int h, int w, int c, int axis, int radius,
                                     uint32_t ww, uint32_t fw) {
  const int64_t total = (int64_t)h * w * c;
  const int n = axis == 1 ? w : h;
  const int64_t stride = axis == 1 ? c : (int64_t)w * c;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)((idx / c) % w), y = (int)(idx / ((int64_t)c * w));
    const int i = axis == 1 ? x : y;
    const uint8_t* line = src + idx - (int64_t)i * stride;
    uint32_t acc = 0;
    for (int d = -radius; d <= radius; ++d) {
      const int j = min(max(i + d, 0), n - 1);
      acc += line[(int64_t)j * stride];
    }
    const int jl = max(i - radius - 1, 0), jr = min(i + radius + 1, n - 1);
    const uint32_t bulk = acc * ww + ((uint32_t)line[(int64_t)jl * stride] + (uint32_t)line[(int64_t)jr * stride]) * fw;
    dst[idx] = (uint8_t)((bulk + (1u << 23)) >> 24);
  }
}

}  // namespace
}  // namespace svl

using namespace svl;
#define ST (cudaStream_t) stream

extern "C" int svl_resample_pass_u8(const uint8_t* src, uint8_t* dst, const int* kk, const int* bounds, int ksize, int h, int w, int c,
                                    int out_size, int axis, void* stream) {
  SVL_CHECK_ARG(src && dst && kk && bounds && ksize > 0 && out_size > 0 && (axis == 0 || axis == 1), "svl_resample_pass_u8: bad arguments");
  const int64_t total = (int64_t)(axis == 0 ? out_size : h) * (axis == 1 ? out_size : w) * c;
  resample_pass_kernel<<<ew_grid(total), 256, 0, ST>>>(src, dst, kk, bounds, ksize, h, w, c, out_size, axis);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_gather_nearest_u8(const uint8_t* src, uint8_t* dst, const int* ys, const int* xs, int w, int oh, int ow, void* stream) {
  SVL_CHECK_ARG(src && dst && ys && xs, "svl_gather_nearest_u8: bad arguments");
  gather_nearest_kernel<<<ew_grid((int64_t)oh * ow), 256, 0, ST>>>(src, dst, ys, xs, w, oh, ow);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_color_op_u8(uint8_t* img, int64_t npix, int mode, float factor, int shift, unsigned long long* scratch, void* stream) {
  SVL_CHECK_ARG(img && mode >= 0 && mode <= 4 && (mode != 1 || scratch), "svl_color_op_u8: bad arguments");
  if (mode == 1) {
    SVL_CUDA(cudaMemsetAsync(scratch, 0, sizeof(unsigned long long), ST));
    gray_sum_kernel<<<ew_grid(npix), 256, 0, ST>>>(img, npix, scratch);
    SVL_LAUNCH_CHECK();
  }
  color_op_kernel<<<ew_grid(npix), 256, 0, ST>>>(img, npix, mode, factor, scratch, shift);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_box_blur_pass_u8(const uint8_t* src, uint8_t* dst, int h, int w, int c, int axis, int radius, unsigned ww, unsigned fw,
                                    void* stream) {
  SVL_CHECK_ARG(src && dst && src != dst && radius >= 0 && (axis == 0 || axis == 1), "svl_box_blur_pass_u8: bad arguments");
  box_blur_pass_kernel<<<ew_grid((int64_t)h * w * c), 256, 0, ST>>>(src, dst, h, w, c, axis, radius, ww, fw);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
