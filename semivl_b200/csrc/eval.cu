// Evaluation-side kernels (third_party/unimatch/supervised.py:40-164 of the reference): sliding-window stitching of the per-window
// logits / probabilities, the final arg-max, and the intersection / union / target histograms of the mIoU (integer, bit-exact).
// All HBM-bound: a thread owns one pixel and walks the class planes (stride = plane size, coalesced across the warp).
#include "common.cuh"

namespace svl {
namespace {

inline int ew_grid_eval(int64_t total, int block = 256) {
  int64_t b = cdiv(total, block);
  const int64_t cap = 148 * 32;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// dst[b, :, y1 + y, x1 + x] += f(src[b, :, sy + y, sx + x]) for y < ch, x < cw;  f = softmax over classes (mode 1) or identity (mode 0);
// count[b, y1 + y, x1 + x] += 1 when count != NULL   (supervised.py:58-59, 88-92, 113-114)
__global__ void window_accumulate_kernel(float* __restrict__ dst, const float* __restrict__ src, float* __restrict__ count, int B, int N, int H,
                                         int W, int h, int w, int y1, int x1, int sy, int sx, int ch, int cw, int mode) {
  const int64_t total = (int64_t)B * ch * cw;
  const int64_t splane = (int64_t)h * w, dplane = (int64_t)H * W;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % cw), y = (int)((idx / cw) % ch), b = (int)(idx / ((int64_t)cw * ch));
    const float* s = src + (int64_t)b * N * splane + (int64_t)(sy + y) * w + sx + x;
    float* d = dst + (int64_t)b * N * dplane + (int64_t)(y1 + y) * W + x1 + x;
    if (mode == 1) {
      float m = -INFINITY;
      for (int n = 0; n < N; ++n) m = fmaxf(m, s[n * splane]);
      float se = 0.f;
      for (int n = 0; n < N; ++n) se += expf(s[n * splane] - m);
      for (int n = 0; n < N; ++n) d[n * dplane] += expf(s[n * splane] - m) / se;          // torch.softmax: exp(x - max) / sum
    } else {
      for (int n = 0; n < N; ++n) d[n * dplane] += s[n * splane];
    }
    if (count) count[(int64_t)b * dplane + (int64_t)(y1 + y) * W + x1 + x] += 1.f;
  }
}
// x[b, n, p] /= count[b, p]        (supervised.py:94)
__global__ void divide_count_kernel(float* __restrict__ x, const float* __restrict__ count, int B, int N, int64_t plane) {
  const int64_t total = (int64_t)B * N * plane;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = idx % plane, b = idx / (plane * N);
    x[idx] = x[idx] / count[b * plane + p];
  }
}
// out[b, p] = index of the FIRST maximal class of x[b, :, p]   (torch.argmax(dim=1))
__global__ void argmax_classes_kernel(const float* __restrict__ x, int64_t* __restrict__ out, int B, int N, int64_t plane) {
  const int64_t total = (int64_t)B * plane;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = idx % plane, b = idx / plane;
    const float* s = x + b * N * plane + p;
    float m = s[0];
    int best = 0;
    for (int n = 1; n < N; ++n) {
      const float v = s[n * plane];
      if (v > m || (v != v && m == m)) { m = v; best = n; }        // NaN counts as the maximum, like torch
    }
    out[idx] = best;
  }
}
// intersectionAndUnion (third_party/unimatch/util/utils.py:91-103): pred is set to `ignore` where target == ignore, then
//   counts[0][k] = #(pred == target == k), counts[1][k] = #(pred == k), counts[2][k] = #(target == k)   for k in [0, K)
// (area_union = counts[1] + counts[2] - counts[0] is formed by the caller).  Block-private histograms in shared memory, 64-bit totals.
__global__ void __launch_bounds__(256)
intersection_union_kernel(const int64_t* __restrict__ pred, const int64_t* __restrict__ target, int64_t n, int K, int ignore,
                          unsigned long long* __restrict__ counts) {
  extern __shared__ unsigned int s_hist[];        // [3][K]
  for (int i = threadIdx.x; i < 3 * K; i += blockDim.x) s_hist[i] = 0u;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = target[i];
    const int64_t p = t == ignore ? (int64_t)ignore : pred[i];
    if (p >= 0 && p < K) {
      atomicAdd(&s_hist[K + (int)p], 1u);
      if (p == t) atomicAdd(&s_hist[(int)p], 1u);
    }
    if (t >= 0 && t < K) atomicAdd(&s_hist[2 * K + (int)t], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * K; i += blockDim.x)
    if (s_hist[i]) atomicAdd(counts + i, (unsigned long long)s_hist[i]);
}

}  // namespace
}  // namespace svl

using namespace svl;
#define ST (cudaStream_t) stream

extern "C" int svl_window_accumulate(float* dst, const float* src, float* count, int B, int N, int H, int W, int h, int w, int y1, int x1, int sy,
                                     int sx, int ch, int cw, int softmax, void* stream) {
  SVL_CHECK_ARG(dst && src && B > 0 && N > 0, "svl_window_accumulate: bad arguments");
  SVL_CHECK_ARG(y1 >= 0 && x1 >= 0 && sy >= 0 && sx >= 0 && y1 + ch <= H && x1 + cw <= W && sy + ch <= h && sx + cw <= w,
                "svl_window_accumulate: window [%d+%d, %d+%d] does not fit %dx%d / source %dx%d", y1, ch, x1, cw, H, W, h, w);
  if (ch == 0 || cw == 0) return SVL_OK;
  window_accumulate_kernel<<<ew_grid_eval((int64_t)B * ch * cw), 256, 0, ST>>>(dst, src, count, B, N, H, W, h, w, y1, x1, sy, sx, ch, cw, softmax ? 1 : 0);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_divide_count(float* x, const float* count, int B, int N, int64_t plane, void* stream) {
  SVL_CHECK_ARG(x && count, "svl_divide_count: null pointer");
  if ((int64_t)B * N * plane == 0) return SVL_OK;
  divide_count_kernel<<<ew_grid_eval((int64_t)B * N * plane), 256, 0, ST>>>(x, count, B, N, plane);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_argmax_classes(const float* x, int64_t* out, int B, int N, int64_t plane, void* stream) {
  SVL_CHECK_ARG(x && out && N >= 1, "svl_argmax_classes: bad arguments");
  if ((int64_t)B * plane == 0) return SVL_OK;
  argmax_classes_kernel<<<ew_grid_eval((int64_t)B * plane), 256, 0, ST>>>(x, out, B, N, plane);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_intersection_union(const int64_t* pred, const int64_t* target, int64_t n, int K, int ignore_index, int64_t* counts, void* stream) {
  SVL_CHECK_ARG(counts && K >= 1 && K <= 4096 && n >= 0, "svl_intersection_union: bad arguments");
  if (n == 0) return SVL_OK;                      // an empty batch (torch gives empty tensors a null data pointer) adds nothing
  SVL_CHECK_ARG(pred && target, "svl_intersection_union: null pointer");
  int64_t blocks = cdiv(n, 256 * 16);
  if (blocks > 148 * 8) blocks = 148 * 8;
  intersection_union_kernel<<<(unsigned)blocks, 256, 3 * K * sizeof(unsigned int), ST>>>(pred, target, n, K, ignore_index,
                                                                                       (unsigned long long*)counts);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
