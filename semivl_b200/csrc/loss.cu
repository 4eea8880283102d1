// Logit-side kernels: bilinear upsampling of the 4x-resolution class maps (align_corners=False), pseudo-labels
// (softmax -> max / argmax), the fused upsample + per-pixel cross-entropy (+ confidence weights) forward/backward that never
// materialises full-resolution logits, label utilities and the fused AdamW update.
// replaces: vlg_head.py:247-248 + builder.py:93-97 (resize), semivl.py:231-232,251-252 (pseudo-labels),
//           semivl.py:52-58,266-323 + utils/train_utils.py:30-49 (losses), semivl.py:326-328 (optimizer step).
#include "common.cuh"

namespace svl {
namespace {

inline int ew_grid(int64_t total, int block = 256) {
  int64_t b = cdiv(total, block);
  int64_t cap = 148 * 32;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// align_corners=False source index (PyTorch area_pixel_compute_source_index)
__device__ __forceinline__ void src_index(int o, float scale, int in_size, int& i0, int& i1, float& w1) {
  float s = (o + 0.5f) * scale - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + 1 < in_size ? i0 + 1 : i0;
  w1 = s - i0;
}

// ---------------------------------------------------------------------------------------------- plain upsample (API output)
__global__ void upsample_kernel(const float* __restrict__ low, float* __restrict__ out, int64_t planes, int hl, int wl, int H, int W) {
  const float sy = (float)hl / H, sx = (float)wl / W;
  const int64_t total = planes * H * W;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(idx % W), Y = (int)((idx / W) % H);
    const int64_t pl = idx / ((int64_t)W * H);
    int y0, y1, x0, x1;
    float wy, wx;
    src_index(Y, sy, hl, y0, y1, wy);
    src_index(X, sx, wl, x0, x1, wx);
    const float* p = low + pl * hl * wl;
    out[idx] = (1.f - wy) * ((1.f - wx) * p[y0 * wl + x0] + wx * p[y0 * wl + x1]) + wy * ((1.f - wx) * p[y1 * wl + x0] + wx * p[y1 * wl + x1]);
  }
}
// align_corners=True resize (mmseg.ops.resize of the stitched logits to the label size, supervised.py:95-100): source = o * (in - 1) / (out - 1)
__global__ void resize_ac_kernel(const float* __restrict__ src, float* __restrict__ out, int64_t planes, int hs, int ws, int H, int W) {
  const float sy = H > 1 ? (float)(hs - 1) / (H - 1) : 0.f, sx = W > 1 ? (float)(ws - 1) / (W - 1) : 0.f;
  const int64_t total = planes * H * W;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(idx % W), Y = (int)((idx / W) % H);
    const int64_t pl = idx / ((int64_t)W * H);
    const float fy = Y * sy, fx = X * sx;
    const int y0 = min((int)fy, hs - 1), x0 = min((int)fx, ws - 1);
    const int y1 = min(y0 + 1, hs - 1), x1 = min(x0 + 1, ws - 1);
    const float wy = fy - y0, wx = fx - x0;
    const float* p = src + pl * hs * ws;
    out[idx] = (1.f - wy) * ((1.f - wx) * p[y0 * ws + x0] + wx * p[y0 * ws + x1]) + wy * ((1.f - wx) * p[y1 * ws + x0] + wx * p[y1 * ws + x1]);
  }
}
// d_low += upsample^T(d_out)   (gather form; used when a caller backpropagates through the materialised logits)
__global__ void upsample_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dlow, int64_t planes, int hl, int wl, int H, int W) {
  const float sy = (float)hl / H, sx = (float)wl / W;
  const int64_t total = planes * hl * wl;
  const float ry = (float)H / hl, rx = (float)W / wl;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % wl), y = (int)((idx / wl) % hl);
    const int64_t pl = idx / ((int64_t)wl * hl);
    const int Ylo = max(0, (int)floorf((y - 1 + 0.5f) * ry - 0.5f) - 1), Yhi = min(H - 1, (int)ceilf((y + 1 + 0.5f) * ry - 0.5f) + 1);
    const int Xlo = max(0, (int)floorf((x - 1 + 0.5f) * rx - 0.5f) - 1), Xhi = min(W - 1, (int)ceilf((x + 1 + 0.5f) * rx - 0.5f) + 1);
    float s = 0.f;
    for (int Y = Ylo; Y <= Yhi; ++Y) {
      int y0, y1; float wy;
      src_index(Y, sy, hl, y0, y1, wy);
      const float ay = (y0 == y ? 1.f - wy : 0.f) + (y1 == y ? wy : 0.f);
      if (ay == 0.f) continue;
      for (int X = Xlo; X <= Xhi; ++X) {
        int x0, x1; float wx;
        src_index(X, sx, wl, x0, x1, wx);
        const float ax = (x0 == x ? 1.f - wx : 0.f) + (x1 == x ? wx : 0.f);
        if (ax != 0.f) s += ay * ax * dout[(pl * H + Y) * W + X];
      }
    }
    dlow[idx] += s;
  }
}

// ---------------------------------------------------------------------------------------------- pseudo labels
// conf[r, Y, X] = max_n softmax(scale * up(low)[r, :, Y, X]), label = argmax (first maximum, like torch.max); optional threshold -> 255
__global__ void softmax_max_kernel(const float* __restrict__ low, float* __restrict__ conf, int64_t* __restrict__ label, int64_t R, int N, int hl,
                                   int wl, int H, int W, float scale, float thresh) {
  const float sy = (float)hl / H, sx = (float)wl / W;
  const int64_t total = R * H * W;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(idx % W), Y = (int)((idx / W) % H);
    const int64_t r = idx / ((int64_t)W * H);
    int y0, y1, x0, x1;
    float wy, wx;
    src_index(Y, sy, hl, y0, y1, wy);
    src_index(X, sx, wl, x0, x1, wx);
    const float w00 = (1.f - wy) * (1.f - wx), w01 = (1.f - wy) * wx, w10 = wy * (1.f - wx), w11 = wy * wx;
    const int o00 = y0 * wl + x0, o01 = y0 * wl + x1, o10 = y1 * wl + x0, o11 = y1 * wl + x1;
    const float* p = low + r * N * hl * wl;
    float m = -INFINITY, sum = 0.f;
    int arg = 0;
    for (int n = 0; n < N; ++n, p += hl * wl) {
      const float v = scale * (w00 * p[o00] + w01 * p[o01] + w10 * p[o10] + w11 * p[o11]);
      if (v > m) {
        sum = sum * __expf(m - v) + 1.f;
        m = v;
        arg = n;
      } else {
        sum += __expf(v - m);
      }
    }
    const float c = 1.f / sum;
    if (conf) conf[idx] = c;
    if (label) label[idx] = (thresh > 0.f && c < thresh) ? 255 : arg;
  }
}

// pixel-major scores [R*hw, ld] (column = concept) -> class-major maps [R, N, hw]: out[r, n, p] = max_{k in [off[n], off[n+1])} in[r*hw + p, k]
// (aggregate_concept_predictions, model/text_embeddings.py:188-193; identity grouping is a plain transpose)
__global__ void group_max_kernel(const float* __restrict__ in, int64_t ld, const int* __restrict__ off, float* __restrict__ out, int64_t R, int N,
                                 int hw) {
  const int64_t total = R * N * hw;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(idx % hw), n = (int)((idx / hw) % N);
    const int64_t r = idx / ((int64_t)hw * N);
    const float* row = in + (r * hw + p) * ld;
    float m = -INFINITY;
    for (int k = off[n]; k < off[n + 1]; ++k) m = fmaxf(m, row[k]);
    out[idx] = m;
  }
}

// ---------------------------------------------------------------------------------------------- fused upsample + CE fwd/bwd
// For every full-resolution pixel: z = up(low)[:, Y, X];  for each target set t:  ce_t = logsumexp(z) - z[label_t] (0 when label_t == ignore);
//   loss[t] += coef[t] * w_t(pix) * ce_t ;   dlow += up^T( sum_t gscale * coef[t] * w_t * (softmax(z) - onehot(label_t)) )
// Tile = 32 x 32 full-res pixels; the low-res footprint of the tile (logits and gradient) lives in shared memory.
constexpr int TILE = 32;
constexpr int kMaxTargets = 3;

struct CeParams {
  const float* low; float* dlow;
  int R, N, hl, wl, H, W, fh, fw;
  int num_targets;
  const int64_t* label[kMaxTargets];
  const float* weight[kMaxTargets];     // per-pixel weight or NULL
  const float* coef[kMaxTargets];       // device scalar
  float* loss;                          // [num_targets], atomically accumulated (already multiplied by coef)
  float gscale;                         // global factor on the gradient only
  int ignore_index;
};

__global__ void __launch_bounds__(256)
upsample_ce_kernel(const __grid_constant__ CeParams p) {
  extern __shared__ float sm[];
  const int fcells = p.fh * p.fw;
  float* s_z = sm;                          // [N][fh][fw]
  float* s_g = sm + (size_t)p.N * fcells;   // [N][fh][fw]
  __shared__ float s_loss[kMaxTargets];
  const int r = blockIdx.z, Y0 = blockIdx.y * TILE, X0 = blockIdx.x * TILE;
  const float sy = (float)p.hl / p.H, sx = (float)p.wl / p.W;
  int ylo, xlo, t0, t1;
  float tw;
  src_index(Y0, sy, p.hl, ylo, t1, tw);
  src_index(X0, sx, p.wl, xlo, t1, tw);
  (void)t0;
  const float* low = p.low + (int64_t)r * p.N * p.hl * p.wl;
  for (int i = threadIdx.x; i < p.N * fcells; i += blockDim.x) {
    const int n = i / fcells, c = i % fcells;
    const int yy = ylo + c / p.fw, xx = xlo + c % p.fw;
    s_z[i] = (yy < p.hl && xx < p.wl) ? low[((int64_t)n * p.hl + yy) * p.wl + xx] : 0.f;
    s_g[i] = 0.f;
  }
  if (threadIdx.x < kMaxTargets) s_loss[threadIdx.x] = 0.f;
  __syncthreads();

  float coef[kMaxTargets];
#pragma unroll
  for (int t = 0; t < kMaxTargets; ++t) coef[t] = t < p.num_targets ? __ldg(p.coef[t]) : 0.f;
  float lsum[kMaxTargets] = {0.f, 0.f, 0.f};

  // Pixel order inside the tile: the 32 lanes of a warp take the SAME sub-position of 32 DIFFERENT 4x4 pixel cells (cell = tid % 64,
  // row of the cell = tid / 64, column = iteration), so at the usual 4x upsampling every lane of a gradient atomic below hits its own
  // low-res entry; with row-major lanes four neighbours shared each entry and the shared-memory atomics serialised 4-way.
  static_assert(TILE == 32, "the cell mapping below covers a 32 x 32 tile with 256 threads x 4 iterations");
  for (int it = 0; it < 4; ++it) {
    const int cell = threadIdx.x & 63, sub = threadIdx.x >> 6;
    const int Y = Y0 + (cell >> 3) * 4 + sub, X = X0 + (cell & 7) * 4 + it;
    if (Y >= p.H || X >= p.W) continue;
    int y0, y1, x0, x1;
    float wy, wx;
    src_index(Y, sy, p.hl, y0, y1, wy);
    src_index(X, sx, p.wl, x0, x1, wx);
    const float w00 = (1.f - wy) * (1.f - wx), w01 = (1.f - wy) * wx, w10 = wy * (1.f - wx), w11 = wy * wx;
    const int o00 = (y0 - ylo) * p.fw + (x0 - xlo), o01 = (y0 - ylo) * p.fw + (x1 - xlo), o10 = (y1 - ylo) * p.fw + (x0 - xlo),
              o11 = (y1 - ylo) * p.fw + (x1 - xlo);
    const int64_t pix = ((int64_t)r * p.H + Y) * p.W + X;
    int lab[kMaxTargets];
    float wt[kMaxTargets];
    float wsum = 0.f;
#pragma unroll
    for (int t = 0; t < kMaxTargets; ++t) {
      lab[t] = -1; wt[t] = 0.f;
      if (t < p.num_targets) {
        const int l = (int)p.label[t][pix];
        float w = coef[t] * (p.weight[t] ? p.weight[t][pix] : 1.f);
        if (l == p.ignore_index || l < 0 || l >= p.N) w = 0.f;
        lab[t] = l; wt[t] = w;
        wsum += w;
      }
    }
    if (wsum == 0.f && wt[0] == 0.f && wt[1] == 0.f && wt[2] == 0.f) continue;
    // pass 1: logsumexp and the label logits
    float m = -INFINITY, se = 0.f, zl[kMaxTargets] = {0.f, 0.f, 0.f};
    const float* z = s_z;
    for (int n = 0; n < p.N; ++n, z += fcells) {
      const float v = w00 * z[o00] + w01 * z[o01] + w10 * z[o10] + w11 * z[o11];
      if (v > m) { se = se * __expf(m - v) + 1.f; m = v; } else se += __expf(v - m);
#pragma unroll
      for (int t = 0; t < kMaxTargets; ++t) if (n == lab[t]) zl[t] = v;
    }
    const float lse = m + __logf(se);
#pragma unroll
    for (int t = 0; t < kMaxTargets; ++t) lsum[t] += wt[t] * (lse - zl[t]);
    // pass 2: gradient scatter into the shared low-res tile
    if (p.dlow) {
      z = s_z;
      float* g = s_g;
      const float gs = p.gscale;
      for (int n = 0; n < p.N; ++n, z += fcells, g += fcells) {
        const float v = w00 * z[o00] + w01 * z[o01] + w10 * z[o10] + w11 * z[o11];
        float d = wsum * __expf(v - lse);
#pragma unroll
        for (int t = 0; t < kMaxTargets; ++t) if (n == lab[t]) d -= wt[t];
        d *= gs;
        if (d != 0.f) {
          atomicAdd(g + o00, d * w00);
          if (w01 != 0.f) atomicAdd(g + o01, d * w01);
          if (w10 != 0.f) atomicAdd(g + o10, d * w10);
          if (w11 != 0.f) atomicAdd(g + o11, d * w11);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < kMaxTargets; ++t) {
    float v = warp_sum(lsum[t]);
    if ((threadIdx.x & 31) == 0 && t < p.num_targets && v != 0.f) atomicAdd(&s_loss[t], v);
  }
  __syncthreads();
  if (threadIdx.x < p.num_targets && s_loss[threadIdx.x] != 0.f) atomicAdd(p.loss + threadIdx.x, s_loss[threadIdx.x]);
  if (p.dlow) {
    float* dl = p.dlow + (int64_t)r * p.N * p.hl * p.wl;
    for (int i = threadIdx.x; i < p.N * fcells; i += blockDim.x) {
      const float v = s_g[i];
      if (v == 0.f) continue;
      const int n = i / fcells, c = i % fcells;
      const int yy = ylo + c / p.fw, xx = xlo + c % p.fw;
      if (yy < p.hl && xx < p.wl) atomicAdd(dl + ((int64_t)n * p.hl + yy) * p.wl + xx, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------- label utilities
// out[0] = 1 / max(count(label != ignore), 1)   (mean reduction of F.cross_entropy with ignore_index);  two-kernel reduce
__global__ void count_valid_kernel(const int64_t* __restrict__ label, int64_t n, int ignore_index, float* __restrict__ count) {
  __shared__ float red[32];
  float c = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) c += label[i] != ignore_index;
  c = block_sum(c, red);
  if (threadIdx.x == 0 && c != 0.f) atomicAdd(count, c);
}
__global__ void reciprocal_kernel(const float* __restrict__ count, float* __restrict__ out, float numer, float floor_) {
  out[0] = numer / fmaxf(count[0], floor_);
}
// CutMix of (label, conf, ignore) + the confidence-weight mask of utils/train_utils.py:30-39 ('pixelwise'):
//   lab = box ? lab_b : lab_a; conf likewise; ign likewise;  w = (conf >= thresh) & (ign != 255);  count += (ign != 255)
__global__ void cutmix_weights_kernel(const int64_t* __restrict__ lab_a, const int64_t* __restrict__ lab_b, const float* __restrict__ conf_a,
                                      const float* __restrict__ conf_b, const int64_t* __restrict__ ign_a, const int64_t* __restrict__ ign_b,
                                      const float* __restrict__ box, int64_t* __restrict__ lab_out, float* __restrict__ w_out,
                                      int64_t* __restrict__ ign_out, float* __restrict__ valid_count, int64_t n, float thresh) {
  __shared__ float red[32];
  float c = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const bool m = box != nullptr && box[i] == 1.f;
    const int64_t l = m ? lab_b[i] : lab_a[i];
    const float cf = conf_a ? (m ? conf_b[i] : conf_a[i]) : 1.f;
    const int64_t ig = ign_a ? (m ? ign_b[i] : ign_a[i]) : 0;
    const bool valid = ig != 255;
    if (lab_out) lab_out[i] = l;
    if (ign_out) ign_out[i] = ig;
    if (w_out) w_out[i] = (cf >= thresh && valid) ? 1.f : 0.f;
    c += valid;
  }
  c = block_sum(c, red);
  if (threadIdx.x == 0 && valid_count && c != 0.f) atomicAdd(valid_count, c);
}
// Per-image statistics of the (CutMixed) confidence map / ignore mask for the 'pixelratio' and 'pixelavg' modes of
// utils/train_utils.py:40-46:  stats[b] = { #valid, #(conf >= thresh & valid), sum(conf * valid) },  valid = (ign != 255).
// grid = (chunks, B); one atomicAdd triple per CTA.
__global__ void conf_stats_kernel(const float* __restrict__ conf_a, const float* __restrict__ conf_b, const int64_t* __restrict__ ign_a,
                                  const int64_t* __restrict__ ign_b, const float* __restrict__ box, float* __restrict__ stats, int64_t hw,
                                  float thresh) {
  __shared__ float red[32];
  const int64_t base = (int64_t)blockIdx.y * hw;
  float nv = 0.f, nh = 0.f, sc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
    const bool m = box != nullptr && box[base + i] == 1.f;
    const float cf = m ? conf_b[base + i] : conf_a[base + i];
    const bool valid = (m ? ign_b[base + i] : ign_a[base + i]) != 255;
    nv += valid;
    nh += (valid && cf >= thresh);
    sc += valid ? cf : 0.f;
  }
  nv = block_sum(nv, red);
  __syncthreads();
  nh = block_sum(nh, red);
  __syncthreads();
  sc = block_sum(sc, red);
  if (threadIdx.x == 0) {
    atomicAdd(stats + 3 * blockIdx.y + 0, nv);
    atomicAdd(stats + 3 * blockIdx.y + 1, nh);
    atomicAdd(stats + 3 * blockIdx.y + 2, sc);
  }
}
// Loss coefficient (device scalar) and per-image weights from the statistics above (one thread; B is the per-GPU batch):
//   mode 0 pixelwise : coef = numer / sum_b nvalid_b                                     (row_w untouched)
//   mode 1 pixelratio: coef = numer / sum_b nvalid_b, row_w[b] = nhigh_b / nvalid_b
//   mode 2 pixelavg  : coef = numer * sum_b (sumconf_b / nvalid_b) / sum_b nvalid_b      (train_utils.py:43-46: loss.sum() * avg_conf)
__global__ void conf_coef_kernel(const float* __restrict__ stats, int B, int mode, float numer, float* __restrict__ coef,
                                 float* __restrict__ row_w) {
  float nv = 0.f, avg = 0.f;
  for (int b = 0; b < B; ++b) {
    nv += stats[3 * b];
    avg += stats[3 * b + 2] / stats[3 * b];
    if (mode == 1 && row_w) row_w[b] = stats[3 * b + 1] / stats[3 * b];
  }
  coef[0] = mode == 2 ? numer * avg / nv : numer / nv;
}
// w[b, :] = row_w[b]  (the per-image 'pixelratio' weight as a per-pixel weight map of the fused CE kernel)
__global__ void fill_rows_kernel(float* __restrict__ w, const float* __restrict__ row_w, int B, int64_t hw) {
  const int64_t total = (int64_t)B * hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) w[i] = row_w[i / hw];
}
// img = box ? img_b : img_a  per pixel, all channels (utils/train_utils.py:19-21), out of place
__global__ void cutmix_img_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ box, float* __restrict__ out,
                                  int B, int C, int64_t hw) {
  const int64_t total = (int64_t)B * C * hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i % hw, bi = i / (C * hw);
    out[i] = box[bi * hw + pix] == 1.f ? b[i] : a[i];
  }
}

// ---------------------------------------------------------------------------------------------- AdamW (torch.optim.AdamW semantics)
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                             float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gr = g[i] * gscale;
    float pv = p[i] * (1.f - lr * wd);
    const float mv = beta1 * m[i] + (1.f - beta1) * gr;
    const float vv = beta2 * v[i] + (1.f - beta2) * gr * gr;
    m[i] = mv;
    v[i] = vv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv -= (lr / bc1) * (mv / denom);
    p[i] = pv;
  }
}

}  // namespace
}  // namespace svl

using namespace svl;
#define ST (cudaStream_t) stream

extern "C" int svl_upsample_bilinear(const float* low, float* out, int64_t planes, int hl, int wl, int H, int W, void* stream) {
  SVL_CHECK_ARG(low && out, "svl_upsample_bilinear: null pointer");
  upsample_kernel<<<ew_grid(planes * H * W), 256, 0, ST>>>(low, out, planes, hl, wl, H, W);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_resize_bilinear_ac(const float* src, float* out, int64_t planes, int hs, int ws, int H, int W, void* stream) {
  SVL_CHECK_ARG(src && out && hs > 0 && ws > 0 && H > 0 && W > 0, "svl_resize_bilinear_ac: bad arguments");
  resize_ac_kernel<<<ew_grid(planes * H * W), 256, 0, ST>>>(src, out, planes, hs, ws, H, W);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_upsample_bilinear_bwd(const float* dout, float* dlow, int64_t planes, int hl, int wl, int H, int W, void* stream) {
  SVL_CHECK_ARG(dout && dlow, "svl_upsample_bilinear_bwd: null pointer");
  upsample_bwd_kernel<<<ew_grid(planes * hl * wl), 256, 0, ST>>>(dout, dlow, planes, hl, wl, H, W);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_softmax_max(const float* low, float* conf, int64_t* label, int64_t R, int N, int hl, int wl, int H, int W, float scale,
                               float thresh, void* stream) {
  SVL_CHECK_ARG(low && (conf || label), "svl_softmax_max: null pointer");
  softmax_max_kernel<<<ew_grid(R * H * W), 256, 0, ST>>>(low, conf, label, R, N, hl, wl, H, W, scale == 0.f ? 1.f : scale, thresh);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_group_max(const float* in, int64_t ld, const int* offsets, float* out, int64_t R, int N, int hw, void* stream) {
  SVL_CHECK_ARG(in && offsets && out, "svl_group_max: null pointer");
  group_max_kernel<<<ew_grid(R * N * hw), 256, 0, ST>>>(in, ld, offsets, out, R, N, hw);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_upsample_ce(const float* low, float* dlow, int R, int N, int hl, int wl, int H, int W, int num_targets,
                               const int64_t* const* labels, const float* const* weights, const float* const* coefs, float* loss, float gscale,
                               int ignore_index, void* stream) {
  SVL_CHECK_ARG(low && labels && coefs && loss && num_targets >= 1 && num_targets <= kMaxTargets, "svl_upsample_ce: bad arguments");
  CeParams p;
  memset(&p, 0, sizeof(p));
  p.low = low; p.dlow = dlow; p.R = R; p.N = N; p.hl = hl; p.wl = wl; p.H = H; p.W = W;
  p.num_targets = num_targets;
  for (int t = 0; t < num_targets; ++t) {
    p.label[t] = labels[t];
    p.weight[t] = weights ? weights[t] : nullptr;
    p.coef[t] = coefs[t];
    SVL_CHECK_ARG(p.label[t] && p.coef[t], "svl_upsample_ce: null target %d", t);
  }
  p.loss = loss; p.gscale = gscale; p.ignore_index = ignore_index;
  // low-res footprint of a 32-pixel tile (+2 for the interpolation neighbour and the fractional start)
  p.fh = (int)ceilf(TILE * (float)hl / H) + 3; if (p.fh > hl) p.fh = hl;
  p.fw = (int)ceilf(TILE * (float)wl / W) + 3; if (p.fw > wl) p.fw = wl;
  const size_t smem = (size_t)2 * N * p.fh * p.fw * sizeof(float);
  SVL_CHECK_ARG(smem <= 200 * 1024, "svl_upsample_ce: %d classes x %dx%d footprint does not fit shared memory", N, p.fh, p.fw);
  static size_t max_set = 0;
  if (smem > 48 * 1024 && smem > max_set) {
    SVL_CUDA(cudaFuncSetAttribute(upsample_ce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    max_set = 200 * 1024;
  }
  dim3 grid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE, R);
  upsample_ce_kernel<<<grid, 256, smem, ST>>>(p);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_count_valid(const int64_t* label, int64_t n, int ignore_index, float* count, void* stream) {
  SVL_CHECK_ARG(label && count, "svl_count_valid: null pointer");
  count_valid_kernel<<<ew_grid(n), 256, 0, ST>>>(label, n, ignore_index, count);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_reciprocal(const float* count, float* out, float numer, float floor_, void* stream) {
  SVL_CHECK_ARG(count && out, "svl_reciprocal: null pointer");
  reciprocal_kernel<<<1, 1, 0, ST>>>(count, out, numer, floor_);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_cutmix_weights(const int64_t* lab_a, const int64_t* lab_b, const float* conf_a, const float* conf_b, const int64_t* ign_a,
                                  const int64_t* ign_b, const float* box, int64_t* lab_out, float* w_out, int64_t* ign_out, float* valid_count,
                                  int64_t n, float thresh, void* stream) {
  SVL_CHECK_ARG(lab_a, "svl_cutmix_weights: null pointer");
  cutmix_weights_kernel<<<ew_grid(n), 256, 0, ST>>>(lab_a, lab_b, conf_a, conf_b, ign_a, ign_b, box, lab_out, w_out, ign_out, valid_count, n, thresh);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_conf_stats(const float* conf_a, const float* conf_b, const int64_t* ign_a, const int64_t* ign_b, const float* box,
                              float* stats, int B, int64_t hw, float thresh, void* stream) {
  SVL_CHECK_ARG(conf_a && ign_a && stats && B > 0 && hw > 0, "svl_conf_stats: bad arguments");
  SVL_CHECK_ARG(box == nullptr || (conf_b && ign_b), "svl_conf_stats: a CutMix box needs the second (conf, ignore) pair");
  const int chunks = (int)((hw + 256 * 8 - 1) / (256 * 8));
  conf_stats_kernel<<<dim3(chunks < 64 ? chunks : 64, B), 256, 0, ST>>>(conf_a, conf_b, ign_a, ign_b, box, stats, hw, thresh);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_conf_coef(const float* stats, int B, int mode, float numer, float* coef, float* row_w, void* stream) {
  SVL_CHECK_ARG(stats && coef && B > 0 && mode >= 0 && mode <= 2, "svl_conf_coef: bad arguments");
  conf_coef_kernel<<<1, 1, 0, ST>>>(stats, B, mode, numer, coef, row_w);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_fill_rows(float* w, const float* row_w, int B, int64_t hw, void* stream) {
  SVL_CHECK_ARG(w && row_w && B > 0 && hw > 0, "svl_fill_rows: bad arguments");
  fill_rows_kernel<<<ew_grid((int64_t)B * hw), 256, 0, ST>>>(w, row_w, B, hw);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_cutmix_img(const float* a, const float* b, const float* box, float* out, int B, int C, int64_t hw, void* stream) {
  SVL_CHECK_ARG(a && b && box && out, "svl_cutmix_img: null pointer");
  cutmix_img_kernel<<<ew_grid((int64_t)B * C * hw), 256, 0, ST>>>(a, b, box, out, B, C, hw);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

// Same update with the per-step scalars read from device memory (hyper = {lr class 0, lr class 1, 1 - beta1^t, sqrt(1 - beta2^t)}), so
// that a captured CUDA graph of the training step can be replayed with a new learning rate / step count.
__global__ void adamw_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                                 const float* __restrict__ hyper, int lr_index, float beta1, float beta2, float eps, float wd, float gscale) {
  const float lr = hyper[lr_index], bc1 = hyper[2], bc2_sqrt = hyper[3];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gr = g[i] * gscale;
    float pv = p[i] * (1.f - lr * wd);
    const float mv = beta1 * m[i] + (1.f - beta1) * gr;
    const float vv = beta2 * v[i] + (1.f - beta2) * gr * gr;
    m[i] = mv;
    v[i] = vv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv -= (lr / bc1) * (mv / denom);
    p[i] = pv;
  }
}
extern "C" int svl_adamw_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, int lr_index, float beta1,
                             float beta2, float eps, float wd, float gscale, void* stream) {
  SVL_CHECK_ARG(p && g && m && v && hyper && (lr_index == 0 || lr_index == 1 || lr_index >= 4), "svl_adamw_dev: bad arguments");
  if (n == 0) return SVL_OK;
  adamw_dev_kernel<<<ew_grid(n), 256, 0, ST>>>(p, g, m, v, n, hyper, lr_index, beta1, beta2, eps, wd, gscale == 0.f ? 1.f : gscale);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
extern "C" int svl_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps, float wd,
                         int step, float gscale, void* stream) {
  SVL_CHECK_ARG(p && g && m && v && step >= 1, "svl_adamw: bad arguments");
  if (n == 0) return SVL_OK;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  adamw_kernel<<<ew_grid(n), 256, 0, ST>>>(p, g, m, v, n, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, gscale == 0.f ? 1.f : gscale);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
