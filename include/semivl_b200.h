/* semivl_b200 -- C ABI of the B200-native SemiVL hot path.
 *
 * Drop-in boundary.  The reference (google-research/semivl) is pure Python/PyTorch and has no FFI of
 * its own; every entry point below replaces the stock ATen/cuDNN/cuBLAS dispatch that
 * the reference makes at the cited file:line.  Conventions (last row):
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless stated;
 *   - the caller owns every buffer; kernels are asynchronous on the passed cudaStream_t (as void*);
 *   - return 0 on success, negative svl_status on failure; svl_last_error() gives the message;
 *   - no internal allocation, no global state beyond the resolved driver entry point.
 * Built for sm_100a only (tcgen05 / TMEM / TMA); there is no fallback path.
 */
#ifndef SEMIVL_B200_H_
#define SEMIVL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  SVL_OK = 0,
  SVL_ERR_INVALID = -1,  /* bad argument / unsupported shape */
  SVL_ERR_CUDA = -2,     /* CUDA runtime / driver error */
  SVL_ERR_ARCH = -3      /* not running on sm_100 */
} svl_status;

/* Storage types of activation tensors.
 * SVL_BF16X2 is the precise-mode operand format: value = hi + lo (two bf16), hi at [row*ld + col] and lo at
 * [row*ld + ld/2 + col]; ld is therefore twice the logical row width.  A contraction over split operands is
 * issued as three tensor-core taps (hi*hi + hi*lo + lo*hi) and reproduces an fp32 contraction to ~2^-17. */
typedef enum { SVL_F32 = 0, SVL_BF16 = 1, SVL_BF16X2 = 2 } svl_dtype;

const char* svl_last_error(void);
int svl_version(void);
/* 0 if the current device is sm_100 and the TMA driver entry point resolved. */
int svl_check_device(void);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core contraction engine (tcgen05.mma, TMEM accumulators, TMA-staged SWIZZLE_128B tiles).
 *
 *   D[m, n] = epilogue( sum_{t < num_taps} sum_{k < k_per_tap} A_t[m, k] * B_t[n, k] )
 *
 * One engine covers: every nn.Linear of the ViT / class-attention blocks (maskclip_vit.py:110-144 via
 * mmcv MultiheadAttention/FFN), the patch-embedding conv as an im2col GEMM (maskclip_vit.py:495), the
 * CLIP projection (maskclip_vit.py:552), the text-similarity einsum (vlg_head.py:217), all 1x1 / 3x3 /
 * dilated convs and the 2x2 transposed convs of the VLG head as implicit GEMMs with one tap per filter
 * position (vlg_head.py:84-137,169,181-190) and their data gradients (same engine, mirrored taps,
 * transposed weights).  A tap selects a pixel shift of A (conv), a column offset into A and a
 * (row, column) offset into B; precise mode (BF16X2 operands) triples the taps.
 * ---------------------------------------------------------------------------------------------- */
#define SVL_MAX_TAPS 32

/* SVL_ACT_GELU_DSAVE (as `act`): out = gelu(z) and preact_out receives gelu'(z) instead of z, so that the data-gradient GEMM of the
 * layer below only multiplies (dact_kind = SVL_ACT_SAVED: dact_src holds the derivative itself). */
typedef enum { SVL_ACT_NONE = 0, SVL_ACT_GELU = 1, SVL_ACT_RELU = 2, SVL_ACT_GELU_DSAVE = 3, SVL_ACT_SAVED = 4 } svl_act;
typedef enum { SVL_OUT_LINEAR = 0, SVL_OUT_CONVT2X2 = 1 } svl_out_mode;

typedef struct {
  /* ---- A operand: activations, bf16, K-major (channels / features contiguous) ---- */
  const void* a;
  int a_conv;           /* 0: 2-D [m, lda] row-major.  1: NHWC image batch [nb, h, w, lda] (m = nb*h*w) */
  int64_t m;            /* output rows */
  int64_t lda;          /* elements per row / pixel */
  int64_t a_cols;       /* valid columns from `a` (TMA inner extent); 0 = lda */
  int nb, h, w;         /* a_conv geometry (of the output pixel grid) */
  int a_map_w;          /* a_conv: row length (pixels) of the A tensor when it differs from w (0 = w).  Used to read a
                           [nb, 2h, 2w, c] tensor as [nb, h, 2w', 2c] with w' = w: the transposed-conv data gradient */
  /* ---- B operand: weights, bf16, [b_rows, ldb], K contiguous ---- */
  const void* b;
  int64_t b_rows;
  int64_t ldb;
  /* ---- contraction ---- */
  int n;                /* output columns */
  int k_per_tap;        /* contraction length per tap, multiple of 16 */
  int num_taps;
  int tap_dy[SVL_MAX_TAPS], tap_dx[SVL_MAX_TAPS];   /* a_conv: input pixel offset per tap (zero padding outside the image) */
  int tap_a_koff[SVL_MAX_TAPS];                     /* column offset into A per tap */
  int tap_b_row[SVL_MAX_TAPS];                      /* row offset into B per tap (added to the output column) */
  int tap_b_col[SVL_MAX_TAPS];                      /* column offset into B per tap (added to k) */
  /* ---- epilogue:  v = alpha*acc + bias[col] + row_bias[(row / row_bias_div) * row_bias_ld + col]
   *                 preact_out <- v ; v = act(v) ; v *= act'(dact_src) ; v += residual ; (v += out) ; out <- v ---- */
  void* out;            /* [m, ldc] */
  int out_dtype;        /* svl_dtype */
  int64_t ldc;
  int out_mode;         /* svl_out_mode; CONVT2X2: row = (img, y, x) of an [*, out_h, out_w] grid, column block q = col / (n/4)
                           scatters to pixel (2y + q/2, 2x + q%2) of an NHWC [*, 2*out_h, 2*out_w, ldc] tensor, channel col % (n/4) */
  int out_h, out_w;
  float alpha;          /* 0 is treated as 1 */
  const float* bias;    /* [n] or NULL */
  const float* row_bias; int64_t row_bias_div; int64_t row_bias_ld;
  int act;              /* svl_act */
  void* preact_out; int preact_dtype; int64_t ld_preact;   /* F32 or BF16, or NULL */
  const void* dact_src; int dact_dtype; int dact_kind; int64_t ld_dact;
                        /* multiply by act'(src): GELU -> src is the saved pre-activation; RELU -> src is the saved OUTPUT, factor (src > 0) */
  const void* residual; int res_dtype; int64_t ldres;      /* F32 or BF16, or NULL */
  int accumulate;       /* 1: out += result (F32 out only) */
  /* ---- tuning ---- */
  int block_n;          /* 0 = auto */
  /* ---- GroupNorm statistics of the output, fused into the epilogue of the 3 x 3 convolutions that svl_conv_gn_splits() accepts ----
   * gn_part != NULL: per (map, slot, group) partial sums (sum, sum of squares) of the STORED (bf16-rounded) output, 16 channels per group, are
   * written to gn_part[((map * splits + slot) * (n / 16) + group) * 2 + {0, 1}] with splits = svl_conv_gn_splits(d): every slot is written
   * exactly once, in an order that depends on the map geometry only; svl_gn_relu_fwd(..., stats_splits = splits) consumes them. */
  float* gn_part;
} svl_gemm_desc;

int svl_gemm(const svl_gemm_desc* d, void* stream);
/* partial sums per map that svl_gemm(d) writes to d->gn_part, or 0 when this problem does not take the fused-statistics path */
int svl_conv_gn_splits(const svl_gemm_desc* d);

/* Weight gradient on the same tensor cores, operands read MN-major straight from their forward layouts:
 *   dw[slot*slot_stride + i*ld_dw + j] += alpha * sum_{taps t of slot} sum_rows DY[row, dy_koff_t + i] * X[shift_t(row), x_koff_t + j]
 * Split-K over row blocks, partial tiles reduced with red.global.add.f32 (caller zero-fills dw or accumulates across calls).
 * replaces: autograd wgrad of nn.Linear / nn.Conv2d / nn.ConvTranspose2d on the hot path (semivl.py:327). */
typedef struct {
  const void* dy; int64_t ld_dy; int64_t dy_cols;   /* bf16 [rows, ld_dy]; dy_cols = valid columns from `dy` (0 = ld_dy) */
  const void* x;  int64_t ld_x;  int64_t x_cols;    /* bf16 [rows, ld_x] or NHWC [nb,h,w,ld_x] */
  int conv;                                         /* 1: both operands are NHWC images and taps shift x */
  int64_t rows; int nb, h, w;
  int m;                                            /* output rows  (dy channels) */
  int n;                                            /* output columns per slot (x channels) */
  int num_taps;
  int tap_dy[SVL_MAX_TAPS], tap_dx[SVL_MAX_TAPS];   /* x is read at pixel + (dy, dx), zero outside the image */
  int tap_dy_koff[SVL_MAX_TAPS], tap_x_koff[SVL_MAX_TAPS];
  int tap_slot[SVL_MAX_TAPS];                       /* output slot per tap; ascending, taps of one slot contiguous */
  int x_map_w;                                      /* conv: row length (pixels) of the x tensor when it differs from w (0 = w); see svl_gemm_desc.a_map_w */
  float* dw; int64_t ld_dw; int64_t slot_stride;
  float alpha;                                      /* 0 is treated as 1 */
  int splits;                                       /* 0 = auto */
} svl_wgrad_desc;

int svl_wgrad(const svl_wgrad_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Token / normalisation kernels (HBM-bound).  `*_dtype` arguments are svl_dtype; `ld*` are row strides in elements.
 * ---------------------------------------------------------------------------------------------- */
/* im2col of non-overlapping p x p patches with bottom/right zero pad (mmseg PatchEmbed 'corner'),
 * img f32 NCHW [b,3,H,W] -> out [b*hp*wp, 3*p*p] (BF16 or BF16X2), column order (c, py, px) = conv weight order.
 * replaces the input side of the PatchEmbed conv, maskclip_vit.py:495. */
int svl_patchify(const float* img, void* out, int out_dtype, int b, int H, int W, int p, int hp, int wp, void* stream);
/* x[b,0,:] = cls + pos[0]; x[b,1+i,:] = patches[b*hw+i,:] + pos[1+i]   (maskclip_vit.py:498-500) */
int svl_assemble_tokens(const float* patches, const float* cls, const float* pos, float* x, int b, int hw, int c, void* stream);
/* Position-table resize for crops whose token grid differs from the table's (maskclip_vit.py:448-459,462-490; 801^2 -> 51^2 from 50^2,
 * 641^2 -> 41^2 from 40^2): out[0] = pos[0] (cls), out[1 + oy*ow + ox] = bicubic(pos[1:] as a [gh, gw, c] grid), torch's kernel
 * (A = -0.75, align_corners=False).  Backward: dpos += resize^T(dout) (red.global.add: dpos is a parameter gradient). */
int svl_pos_resize_fwd(const float* pos, float* out, int gh, int gw, int oh, int ow, int c, void* stream);
int svl_pos_resize_bwd(const float* dout, float* dpos, int gh, int gw, int oh, int ow, int c, void* stream);
/* LayerNorm over the last dim (maskclip_vit.py:111,132,141-142,507,539-541).  mean/rstd [rows] saved when non-NULL. */
int svl_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, void* y, int y_dtype, int64_t ldy,
                      float* mean, float* rstd, int64_t rows, int c, float eps, void* stream);
/* dx = dres1 + dres2 + LN'(dy) (f32, contiguous rows; dres may be NULL); dx_act = optional copy in an operand format;
 * dgamma/dbeta (+=, atomics) when non-NULL. */
int svl_layernorm_bwd(const void* dy, int dy_dtype, int64_t lddy, const float* x, int64_t ldx, const float* gamma,
                      const float* mean, const float* rstd, const float* dres1, const float* dres2, float* dx, void* dx_act,
                      int act_dtype, int64_t ld_act, float* dgamma, float* dbeta, int64_t rows, int c, void* stream);
/* y[r,:] = x[r,:] / max(||x[r,:]||_2, eps)   (maskclip_vit.py:555,589; F.normalize vlg_head.py:215-216) */
int svl_l2norm_fwd(const float* x, int64_t ldx, float* y, void* y_act, int act_dtype, int64_t ld_act, float* inv_norm, int64_t rows,
                   int c, float eps, void* stream);
int svl_l2norm_bwd(const void* dy, int dy_dtype, int64_t lddy, const float* y, const float* inv_norm, float* dx, int64_t lddx,
                   int accumulate, int64_t rows, int c, void* stream);
/* dst = scale * src with a storage-type change (scale 0 is treated as 1); `batch` blocks of `rows` rows, block b starting at
 * element b*batch_stride (dropping / re-inserting the cls token row of [b, L, c] token tensors, maskclip_vit.py:543-546) */
int svl_cast(const void* src, int src_dtype, int64_t ld_src, int64_t src_batch_stride, void* dst, int dst_dtype, int64_t ld_dst,
             int64_t dst_batch_stride, int batch, int64_t rows, int cols, float scale, void* stream);
/* Batched parameter jobs, one launch.  mode 0: refresh the GEMM-operand copies of the trainable weights after the optimizer step
 * (dst[i] = cast(src[idx[i]]), idx < 0 = zero padding; replaces the per-tensor reshape / permute / cast launches the reference's optimizer.step
 * + autocast would do, semivl.py:326-328).  mode 1: scatter staged weight gradients into the parameter layout (dst[idx[i]] += src[i]; src[i] = 0).
 * jobs: device int64 [njobs][5] = {src, dst, idx (int32*), n, flags}; flags bit 0: f32 dst; bit 1 (mode 0): the layout is the plain transpose of
 * the parameter viewed as [R, n / R] with R = flags >> 8 (32 x 32 tiles, idx unused); block_start: device int32 [njobs + 1], first block of
 * each job (a block = 1024 elements, or one 32 x 32 tile of a transpose job). */
int svl_param_jobs(const void* jobs, const void* block_start, int njobs, int total_blocks, int mode, void* stream);
/* out[col] += sum_rows x[row, col]  (bias gradients) */
int svl_colsum(const void* x, int x_dtype, int64_t ld, int64_t rows, int cols, float* out, void* stream);
/* out[i] (+)= sum_b x[b, i]  (pos_embed gradient, maskclip_vit.py:500) */
int svl_batch_sum(const float* x, float* out, int b, int64_t inner, int accumulate, void* stream);
int svl_axpy(float* dst, const float* src, float alpha, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-head self-attention, head_dim 64 (nn.MultiheadAttention core, maskclip_vit.py:77-84,141; also the
 * class-attention of SemanticTransformer, vlg_head.py:39-67).  qkv packed [b, L, 3E] bf16 (q | k | v, heads contiguous
 * inside each third; split != 0: BF16X2 rows of 6E), out [b, L, E] (2E when split), lse [b, heads, L]; never materialises LxL.
 * Backward: delta_ws is a caller-provided f32 scratch of svl_attention_bwd_workspace(split, b, L, heads) floats; dv_add (optional, [b*L, E]) is added to dV before it is
 * stored (the MaskCLIP v-path gradient, maskclip_vit.py:110-118); dqkv has the layout of qkv.
 * ---------------------------------------------------------------------------------------------- */
int svl_attention_fwd(const void* qkv, int split, void* out, float* lse, int b, int L, int heads, float scale, void* stream);
size_t svl_attention_bwd_workspace(int split, int b, int L, int heads);
int svl_attention_bwd(const void* qkv, const void* out, const void* dout, int split, const float* lse, float* delta_ws,
                      const void* dv_add, int dv_add_dtype, int64_t ld_dv_add, void* dqkv, int b, int L, int heads, float scale,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * VLG decode head kernels (model/decode_heads/vlg_head.py).  Activations are NHWC, a "map" is one (image, class) pair,
 * maps ordered (image, class).  HBM-bound: vectorised 16-byte accesses, one pass (two for GroupNorm, second L2-resident).
 * ---------------------------------------------------------------------------------------------- */
/* out = relu(GroupNorm(x)) (+ res); statistics per (map, group) saved to mean/rstd [maps, G]   (vlg_head.py:99-111,132-135).
 * The statistics are reduced in two fixed-order stages (per-CTA partials in `ws`, then one thread per (map, group)): no
 * floating-point atomics, so the result is bit-reproducible from run to run.  `ws`: svl_gn_workspace(maps, hw, C, G) floats.
 * stats_splits > 0: the first stage already happened -- `ws` holds [maps, stats_splits, G, 2] partial sums written by the producing convolution
 * (svl_gemm_desc.gn_part) -- and only the second stage and the apply pass run. */
size_t svl_gn_workspace(int64_t maps, int hw, int C, int G);
int svl_gn_relu_fwd(const void* x, int x_dtype, int64_t ldx, const float* gamma, const float* beta, void* out, int out_dtype,
                    int64_t ldo, const void* res, int res_dtype, int64_t ldres, float* mean, float* rstd, float* ws, int64_t maps,
                    int hw, int C, int G, float eps, int stats_splits, void* stream);
/* dx = GN'(dy * [y > 0]) (data gradient reduced in fixed order like the forward); dgamma/dbeta += (atomics, may be NULL) */
int svl_gn_relu_bwd(const void* dy, int dy_dtype, int64_t lddy, const void* x, int x_dtype, int64_t ldx, const float* gamma,
                    const float* beta, const float* mean, const float* rstd, void* dx, int dx_dtype, int64_t lddx, float* dgamma,
                    float* dbeta, float* ws /* svl_gn_workspace floats */, int64_t maps, int hw, int C, int G, void* stream);
/* im2col of the per-class similarity maps for conv1 (ks x ks, 1 -> C; vlg_head.py:169,220-221) and its transpose */
int svl_sim_im2col(const float* sim, int64_t ld_sim, void* out, int out_dtype, int64_t ldo, int B, int N, int h, int w, int ks, int kpad,
                   void* stream);
int svl_sim_col2im(const void* dcol, int dtype, int64_t ld, void* dsim, int ds_dtype, int64_t ld_ds, int ncols, int B, int N, int h,
                   int w, int ks, void* stream);
/* out[map, c] = scale * sum_pix x[map, pix, c] (ASPP image pooling, vlg_head.py:70-81);  x[map, pix, c] += scale * v[map, c] */
int svl_map_sum(const void* x, int dtype, int64_t ld, float* out, int64_t maps, int hw, int C, float scale, void* stream);
int svl_map_bcast_add(float* x, const void* v, int v_dtype, int64_t ldv, int64_t maps, int hw, int C, float scale, void* stream);
/* SemanticTransformer (vlg_head.py:39-67): tokens [(b, py, px), n, C + Ct] = (avg-pooled feature | projected text);
 * un-pooling = bilinear (align_corners=True) of the first C token channels added to the feature map. */
int svl_pool_tokens(const void* x, int x_dtype, int64_t ldx, const float* text, float* tok, int B, int N, int h, int w, int C, int Ct,
                    int pool, void* stream);
int svl_pool_tokens_bwd(const float* dtok, int64_t ldt, float* dx, int B, int N, int h, int w, int C, int pool, void* stream);
/* dx = src + pool-gradient of dtok in ONE pass (src bf16 or f32 [B*N*h*w, lds]; dx f32 [.., C]): replaces svl_cast + svl_pool_tokens_bwd */
int svl_pool_tokens_bwd_from(const void* src, int src_dtype, int64_t lds, const float* dtok, int64_t ldt, float* dx, int B, int N, int h, int w,
                             int C, int pool, void* stream);
int svl_unpool_add(const void* x, int x_dtype, int64_t ldx, const float* tok, int64_t ldt, void* out, int out_dtype, int64_t ldo, int B,
                   int N, int h, int w, int C, int hp, int wp, void* stream);
int svl_unpool_bwd(const void* dout, int dtype, int64_t ld, float* dtok, int64_t ldt, int B, int N, int h, int w, int C, int hp, int wp,
                   void* stream);
/* Up (vlg_head.py:116-137): channels [c0, c0+Cs) of the concat buffer = bilinear (align_corners=True) skip, repeated over classes;
 * gradient w.r.t. the pre-ReLU skip projection. */
int svl_skip_fill(const void* skip, int s_dtype, int64_t lds, void* cat, int c_dtype, int64_t ldc, int c0, int B, int N, int h, int w,
                  int Cs, int H2, int W2, void* stream);
/* out[b, p, c] = sum_n x[(b, n), p, c0 + c] (f32): collapses the per-class copies (transpose of the repeat in vlg_head.py:129) */
int svl_class_sum(const void* x, int dtype, int64_t ld, int c0, float* out, int B, int N, int64_t P, int Cs, void* stream);
int svl_skip_grad(const void* dcat, int d_dtype, int64_t ldd, int c0, const void* skip, int s_dtype, int64_t lds, void* dskip,
                  int o_dtype, int64_t ldo, int B, int N, int h, int w, int Cs, int H2, int W2, void* stream);
/* output conv 3x3, C -> 1 (vlg_head.py:190,239-240); wgt/dw layout [9][C] */
int svl_conv_out1_fwd(const void* x, int dtype, int64_t ld, const float* wgt, const float* bias, float* out, int64_t maps, int h, int w,
                      int C, void* stream);
int svl_conv_out1_bwd(const float* dout, const void* x, int x_dtype, int64_t ldx, const float* wgt, void* dx, int dx_dtype, int64_t lddx,
                      float* dw, float* dbias, int64_t maps, int h, int w, int C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Conv encoder of the Cityscapes skr04 model: mmseg ResNetV1c(depth=101, num_stages=1) = deep stem + layer1 with SyncBN
 * (configs/_base_/models/vlm-vlg-aspp-s2p4-skr04-ftap-mcvitb.py:50-60, model/vlm.py:50-52,120-121).  The convolutions run on
 * svl_gemm / svl_wgrad; these are the HBM-bound pieces around them.  NHWC activations, rows = B*H*W.
 * ---------------------------------------------------------------------------------------------- */
/* im2col of the 3x3 / stride 2 / pad 1 stem convolution on the f32 NCHW image: out[(b,oy,ox), c*9 + ky*3 + kx], columns 27..31 zero
 * (out_dtype BF16 or BF16X2; the [32, 27] flattened weight padded to K = 32 is the matching operand) */
int svl_stem_im2col(const float* img, void* out, int out_dtype, int64_t ldo, int B, int H, int W, int Ho, int Wo, void* stream);
/* max-pool 3x3 / stride 2 / pad 1; widx [B*Ho*Wo, C] bytes = window position of the (first) maximum, consumed by the backward gather */
int svl_maxpool3s2_fwd(const void* x, int dtype, int64_t ldx, void* out, int out_dtype, int64_t ldo, uint8_t* widx, int B, int H, int W, int C,
                       int Ho, int Wo, void* stream);
int svl_maxpool3s2_bwd(const void* dy, int dy_dtype, int64_t lddy, const uint8_t* widx, void* dx, int dx_dtype, int64_t lddx, int B, int H, int W,
                       int C, int Ho, int Wo, void* stream);
/* BatchNorm in training mode, split so that the host can all-reduce the statistics over the ranks (SyncBN):
 *   svl_bn_stats     sums[0..C) = sum_rows x, sums[C..2C) = sum_rows x^2 (two fixed-order stages, no atomics; ws: svl_bn_workspace floats)
 *   [all-reduce sums and the row count over NCCL]
 *   svl_bn_finalize  mean / rstd of the (global) batch (biased variance), running statistics updated with `momentum` (unbiased variance;
 *                    running_* may be NULL)
 *   svl_bn_apply     out = [relu]((x - mean) * rstd * gamma + beta [+ res])      (eval mode: mean = running_mean, rstd = rsqrt(running_var + eps))
 *   svl_bn_bwd_stats sums = [sum g | sum g * xhat], g = dy * [y > 0] (y = saved block output, NULL: no ReLU); dgamma = sums[C..), dbeta = sums[0..C)
 *   [all-reduce sums]
 *   svl_bn_bwd_apply dx = gamma * rstd * (g - S1/N - xhat * S2/N) with the global sums / row count; dres = g when non-NULL (residual branch) */
size_t svl_bn_workspace(int64_t rows, int C);
int svl_bn_stats(const void* x, int x_dtype, int64_t ldx, int64_t rows, int C, float* ws, float* sums, void* stream);
int svl_bn_finalize(const float* sums, float count, float eps, float momentum, float* mean, float* rstd, float* running_mean,
                    float* running_var, int C, void* stream);
int svl_bn_apply(const void* x, int x_dtype, int64_t ldx, const float* mean, const float* rstd, const float* gamma, const float* beta,
                 const void* res, int res_dtype, int64_t ldres, void* out, int out_dtype, int64_t ldo, int relu, int64_t rows, int C,
                 void* stream);
int svl_bn_bwd_stats(const void* dy, int dy_dtype, int64_t lddy, const void* x, int x_dtype, int64_t ldx, const void* y, int y_dtype,
                     int64_t ldy, const float* mean, const float* rstd, int64_t rows, int C, float* ws, float* sums, void* stream);
int svl_bn_bwd_apply(const void* dy, int dy_dtype, int64_t lddy, const void* x, int x_dtype, int64_t ldx, const void* y, int y_dtype,
                     int64_t ldy, const float* mean, const float* rstd, const float* gamma, const float* sums, float count, void* dx,
                     int dx_dtype, int64_t lddx, void* dres, int dres_dtype, int64_t lddres, int64_t rows, int C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Logit-side kernels.  `low` are the head's 4x-resolution class maps [R, N, hl, wl] f32; full-resolution logits
 * [R, N, H, W] = bilinear(low, align_corners=False) (vlg_head.py:247-248; the second resize of builder.py:93-97 is the identity).
 * ---------------------------------------------------------------------------------------------- */
int svl_upsample_bilinear(const float* low, float* out, int64_t planes, int hl, int wl, int H, int W, void* stream);
/* out = bilinear resize with align_corners=True of `planes` maps [hs, ws] -> [H, W]  (mmseg.ops.resize of the stitched evaluation logits to the
 * label size, third_party/unimatch/supervised.py:95-100) */
int svl_resize_bilinear_ac(const float* src, float* out, int64_t planes, int hs, int ws, int H, int W, void* stream);
int svl_upsample_bilinear_bwd(const float* dout, float* dlow, int64_t planes, int hl, int wl, int H, int W, void* stream);   /* dlow += */
/* conf = max_n softmax(scale * logits), label = argmax; label = 255 where conf < thresh (thresh > 0)
 * (semivl.py:231-232,251-252; model/vlm.py:103-109 with scale 100) -- the full-resolution logits are never materialised */
int svl_softmax_max(const float* low, float* conf, int64_t* label, int64_t R, int N, int hl, int wl, int H, int W, float scale,
                    float thresh, void* stream);
/* pixel-major scores [R*hw, ld] -> class-major maps [R, N, hw] with a max over concept columns [offsets[n], offsets[n+1])
 * (device int offsets[N+1]; aggregate_concept_predictions, model/text_embeddings.py:188-193; identity offsets = transpose) */
int svl_group_max(const float* in, int64_t ld, const int* offsets, float* out, int64_t R, int N, int hw, void* stream);
/* Fused upsample + per-pixel cross-entropy forward AND backward for up to 3 target sets on the same logits:
 *   loss[t] += coef[t] * sum_pix weight_t[pix] * CE(logits[pix], label_t[pix])      (label == ignore_index contributes 0)
 *   dlow    += gscale * d(sum_t loss[t]) / d low                                     (skipped when dlow == NULL)
 * labels[t]: int64 [R,H,W]; weights[t]: f32 [R,H,W] or NULL; coefs[t]: DEVICE scalar (e.g. 1/valid-count, no host sync).
 * replaces semivl.py:52-58,266-323 and utils/train_utils.py:30-49 (F.cross_entropy x7 + masks + reductions). */
int svl_upsample_ce(const float* low, float* dlow, int R, int N, int hl, int wl, int H, int W, int num_targets,
                    const int64_t* const* labels, const float* const* weights, const float* const* coefs, float* loss, float gscale,
                    int ignore_index, void* stream);
int svl_count_valid(const int64_t* label, int64_t n, int ignore_index, float* count, void* stream);        /* count += #(label != ignore) */
int svl_reciprocal(const float* count, float* out, float numer, float floor_, void* stream);               /* out = numer / max(count, floor) */
/* CutMix of pseudo-labels / confidences / ignore masks (utils/train_utils.py:24-27) fused with the 'pixelwise' confidence
 * weight of utils/train_utils.py:36-39: w = (conf >= thresh) & (ign != 255); valid_count += #(ign != 255).  NULL inputs are skipped. */
int svl_cutmix_weights(const int64_t* lab_a, const int64_t* lab_b, const float* conf_a, const float* conf_b, const int64_t* ign_a,
                       const int64_t* ign_b, const float* box, int64_t* lab_out, float* w_out, int64_t* ign_out, float* valid_count,
                       int64_t n, float thresh, void* stream);
/* Confidence modes 'pixelratio' / 'pixelavg' of confidence_weighted_loss (utils/train_utils.py:40-46), no host sync:
 * svl_conf_stats: per image b of the CutMixed (conf, ignore) maps, stats[3b..3b+2] += {#valid, #(conf >= thresh & valid), sum(conf * valid)}
 *                 (box == NULL: the first pair unmixed);
 * svl_conf_coef : mode 0 pixelwise  coef = numer / sum_b #valid_b
 *                 mode 1 pixelratio coef = numer / sum_b #valid_b and row_w[b] = #high_b / #valid_b
 *                 mode 2 pixelavg   coef = numer * sum_b(sumconf_b / #valid_b) / sum_b #valid_b;
 * svl_fill_rows : w[b, :] = row_w[b] (per-image weight as the per-pixel weight map svl_upsample_ce takes). */
int svl_conf_stats(const float* conf_a, const float* conf_b, const int64_t* ign_a, const int64_t* ign_b, const float* box, float* stats,
                   int B, int64_t hw, float thresh, void* stream);
int svl_conf_coef(const float* stats, int B, int mode, float numer, float* coef, float* row_w, void* stream);
int svl_fill_rows(float* w, const float* row_w, int B, int64_t hw, void* stream);
int svl_cutmix_img(const float* a, const float* b, const float* box, float* out, int B, int C, int64_t hw, void* stream);
/* torch.optim.AdamW single-tensor update on a flat buffer (semivl.py:326-328; experiments.py:246-255); g is scaled by gscale first */
int svl_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps, float wd,
              int step, float gscale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Evaluation (third_party/unimatch/supervised.py:40-164): sliding-window stitching, arg-max, mIoU histograms.
 * ---------------------------------------------------------------------------------------------- */
/* dst[b, :, y1+y, x1+x] += f(src[b, :, sy+y, sx+x]) for y < ch, x < cw; f = softmax over the N classes (softmax != 0) or identity;
 * count[b, y1+y, x1+x] += 1 when count != NULL.  dst [B,N,H,W], src [B,N,h,w], count [B,H,W], all f32
 * (padded / plain sliding window: supervised.py:58-59,113-114; ZegCLIP window: supervised.py:88-92). */
int svl_window_accumulate(float* dst, const float* src, float* count, int B, int N, int H, int W, int h, int w, int y1, int x1, int sy, int sx,
                          int ch, int cw, int softmax, void* stream);
int svl_divide_count(float* x, const float* count, int B, int N, int64_t plane, void* stream);                  /* x[b,n,p] /= count[b,p] */
int svl_argmax_classes(const float* x, int64_t* out, int B, int N, int64_t plane, void* stream);              /* first maximal class */
/* intersectionAndUnion (third_party/unimatch/util/utils.py:91-103): counts int64 [3][K] += {#(pred==target==k), #(pred==k), #(target==k)}
 * with pred forced to ignore_index where target == ignore_index; bit-exact integer histograms */
int svl_intersection_union(const int64_t* pred, const int64_t* target, int64_t n, int K, int ignore_index, int64_t* counts, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Input stage (third_party/unimatch/dataset/transform.py:9-41,66-84; semi.py:76-107): the per-pixel tail of the sample pipeline on
 * uint8 sources.  mean3 / std3 are HOST pointers to 3 floats.  Outputs are bit-identical to PIL crop / flip + torchvision
 * ToTensor + Normalize for the same crop offset and flip decision.
 * ---------------------------------------------------------------------------------------------- */
int svl_crop_flip_normalize(const uint8_t* src_hwc, int sh, int sw, float* dst_chw, int size, int x0, int y0, int flip, const float* mean3,
                            const float* std3, void* stream);
int svl_crop_flip_mask(const uint8_t* src, int sh, int sw, int64_t* dst, int64_t* ignore_mask, int size, int x0, int y0, int flip, int pad_value,
                       void* stream);
int svl_cutmix_box(float* box, int size, int bx, int by, int bw, int bh, void* stream);
/* Scale / strong augmentations of the unlabelled stream on uint8 HWC images (transform.py:43-64 resize, blur; semi.py:84-93 ColorJitter,
 * RandomGrayscale), integer-exact counterparts of the Pillow / torchvision code paths the reference calls:
 *   svl_resample_pass_u8   one pass of Image.resize(BILINEAR) along x (axis 1) or y (axis 0): kk [out, ksize] 22-bit fixed-point coefficients
 *                          and bounds [out, 2] = (first source index, taps) are DEVICE tables built on the host (Resample.c precompute_coeffs)
 *   svl_gather_nearest_u8  Image.resize(NEAREST) of a label map through host-built index tables ys [oh], xs [ow] (ImagingScaleAffine)
 *   svl_color_op_u8        in place: mode 0 brightness, 1 contrast, 2 saturation (ImageEnhance blends), 3 grayscale, 4 hue shift (`shift` =
 *                          uint8(hue_factor * 255)); scratch = one device uint64 (the gray sum of mode 1)
 *   svl_box_blur_pass_u8   one extended box blur pass of ImageFilter.GaussianBlur (3 along x, then 3 along y; BoxBlur.c): integer radius and
 *                          the 8.24 fixed-point weights ww / fw come from the host */
int svl_resample_pass_u8(const uint8_t* src, uint8_t* dst, const int* kk, const int* bounds, int ksize, int h, int w, int c, int out_size,
                         int axis, void* stream);
int svl_gather_nearest_u8(const uint8_t* src, uint8_t* dst, const int* ys, const int* xs, int w, int oh, int ow, void* stream);
int svl_color_op_u8(uint8_t* img, int64_t npix, int mode, float factor, int shift, unsigned long long* scratch, void* stream);
int svl_box_blur_pass_u8(const uint8_t* src, uint8_t* dst, int h, int w, int c, int axis, int radius, unsigned ww, unsigned fw, void* stream);

/* svl_adamw with the per-step scalars in DEVICE memory: hyper = {lr of class 0, lr of class 1, 1 - beta1^t, sqrt(1 - beta2^t), lr of class 4, ...}
 * (lr_index 0, 1 or >= 4: further learning-rate classes follow the two bias corrections, e.g. the conv_encoder class of the skr04 model);
 * lets a captured CUDA graph of the whole training step be replayed under the poly LR schedule (semivl.py:338-345). */
int svl_adamw_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, int lr_index, float beta1, float beta2,
                  float eps, float wd, float gscale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEMIVL_B200_H_ */
