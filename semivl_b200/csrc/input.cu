// Device-side tail of the reference's per-sample input pipeline (third_party/unimatch/dataset/transform.py:9-41,66-84 and
// third_party/unimatch/dataset/semi.py:76-107): zero / ignore padding to the crop size, random crop, horizontal flip, ToTensor +
// Normalize, label / ignore-mask conversion and the CutMix box raster.  The host keeps the random draws (same order as the
// reference), the GPU does the per-pixel work on uint8 sources -- 4x fewer bytes over PCIe than the fp32 tensors the
// reference's DataLoader ships.  Integer outputs are exact; the normalisation repeats torchvision's operation order
// ((u8 / 255 - mean) / std in IEEE fp32, no fused multiply-add), so the floats are bit-identical as well.
#include "common.cuh"

namespace svl {
namespace {

// out[c, y, x] = (pad(src)[y0 + y, x0 + (flip ? size - 1 - x : x), c] / 255 - mean[c]) / std[c];  src uint8 HWC [sh, sw, 3]
__global__ void crop_flip_normalize_kernel(const uint8_t* __restrict__ src, int sh, int sw, float* __restrict__ dst, int size, int x0, int y0, int flip,
                                           float m0, float m1, float m2, float s0, float s1, float s2) {
  const int total = size * size;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / size, x = i - y * size;
    const int sy = y0 + y, sx = x0 + (flip ? size - 1 - x : x);
    uint8_t r = 0, g = 0, b = 0;                                   // ImageOps.expand(..., fill=0)
    if (sy < sh && sx < sw) {
      const uint8_t* p = src + ((int64_t)sy * sw + sx) * 3;
      r = p[0]; g = p[1]; b = p[2];
    }
    dst[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)r, 255.f), m0), s0);
    dst[total + i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)g, 255.f), m1), s1);
    dst[2 * total + i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)b, 255.f), m2), s2);
  }
}
// labels: dst[y, x] = pad(src, pad_value)[y0 + y, x0 + flip(x)] as int64; optional ignore mask: 255 where the label is 254 else 0
// (semi.py:99-103; unlabelled samples pad with 254 so that the padding can be told from real ignore labels)
__global__ void crop_flip_mask_kernel(const uint8_t* __restrict__ src, int sh, int sw, int64_t* __restrict__ dst, int64_t* __restrict__ ignore_mask,
                                      int size, int x0, int y0, int flip, int pad_value) {
  const int total = size * size;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / size, x = i - y * size;
    const int sy = y0 + y, sx = x0 + (flip ? size - 1 - x : x);
    const int v = (sy < sh && sx < sw) ? (int)src[(int64_t)sy * sw + sx] : pad_value;
    if (dst) dst[i] = v;
    if (ignore_mask) ignore_mask[i] = v == 254 ? 255 : 0;
  }
}
// box[y, x] = 1 inside [by, by + bh) x [bx, bx + bw), else 0     (transform.py:66-84: zeros, then mask[y:y+h, x:x+w] = 1)
__global__ void cutmix_box_kernel(float* __restrict__ box, int size, int bx, int by, int bw, int bh) {
  const int total = size * size;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / size, x = i - y * size;
    box[i] = (y >= by && y < by + bh && x >= bx && x < bx + bw) ? 1.f : 0.f;
  }
}

inline int grid_px(int64_t total) {
  int64_t b = cdiv(total, 256);
  return (int)(b < 148 * 8 ? (b > 0 ? b : 1) : 148 * 8);
}

}  // namespace
}  // namespace svl

using namespace svl;
#define ST (cudaStream_t) stream

extern "C" int svl_crop_flip_normalize(const uint8_t* src, int sh, int sw, float* dst, int size, int x0, int y0, int flip, const float* mean3,
                                       const float* std3, void* stream) {
  SVL_CHECK_ARG(src && dst && mean3 && std3 && sh > 0 && sw > 0 && size > 0, "svl_crop_flip_normalize: bad arguments");
  SVL_CHECK_ARG(x0 >= 0 && y0 >= 0 && x0 + size <= (sw > size ? sw : size) && y0 + size <= (sh > size ? sh : size),
                "svl_crop_flip_normalize: crop (%d, %d) + %d leaves the padded %d x %d source", x0, y0, size, sh, sw);
  crop_flip_normalize_kernel<<<grid_px((int64_t)size * size), 256, 0, ST>>>(src, sh, sw, dst, size, x0, y0, flip ? 1 : 0, mean3[0], mean3[1], mean3[2],
                                                                          std3[0], std3[1], std3[2]);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_crop_flip_mask(const uint8_t* src, int sh, int sw, int64_t* dst, int64_t* ignore_mask, int size, int x0, int y0, int flip,
                                  int pad_value, void* stream) {
  SVL_CHECK_ARG(src && (dst || ignore_mask) && sh > 0 && sw > 0 && size > 0, "svl_crop_flip_mask: bad arguments");
  SVL_CHECK_ARG(x0 >= 0 && y0 >= 0 && x0 + size <= (sw > size ? sw : size) && y0 + size <= (sh > size ? sh : size),
                "svl_crop_flip_mask: crop (%d, %d) + %d leaves the padded %d x %d source", x0, y0, size, sh, sw);
  crop_flip_mask_kernel<<<grid_px((int64_t)size * size), 256, 0, ST>>>(src, sh, sw, dst, ignore_mask, size, x0, y0, flip ? 1 : 0, pad_value);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_cutmix_box(float* box, int size, int bx, int by, int bw, int bh, void* stream) {
  SVL_CHECK_ARG(box && size > 0 && bx >= 0 && by >= 0 && bw >= 0 && bh >= 0, "svl_cutmix_box: bad arguments");
  cutmix_box_kernel<<<grid_px((int64_t)size * size), 256, 0, ST>>>(box, size, bx, by, bw, bh);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
