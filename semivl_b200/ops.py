"""Host-side launch helpers over the C ABI: they fill the descriptors of include/semivl_b200.h from torch tensors.

Storage conventions (see svl_dtype):
  * fast mode    : GEMM operands are bf16 [rows, C]
  * precise mode : GEMM operands are split bf16 pairs [rows, 2C] (hi | lo) -> three tensor-core taps per contraction
A `Mat` is just (tensor, split flag); weights are prepared once per step by `prep_weight`.
"""
import torch

from . import lib as L


def split_bf16(x):
    """fp32 [..., C] -> bf16 [..., 2C] = (hi | lo) with hi + lo ~= x to 2^-17."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.cat((hi, lo), dim=-1).contiguous()


def prep_weight(w2d, precise):
    """fp32 [rows, K] -> bf16 operand ([rows, K] or split [rows, 2K])."""
    w2d = w2d.detach().float()
    return split_bf16(w2d) if precise else w2d.to(torch.bfloat16).contiguous()


def _set_taps(d, taps):
    d.num_taps = len(taps)
    assert len(taps) <= L.MAX_TAPS, len(taps)
    for i, (dy, dx, ak, br, bc) in enumerate(taps):
        d.tap_dy[i], d.tap_dx[i], d.tap_a_koff[i], d.tap_b_row[i], d.tap_b_col[i] = dy, dx, ak, br, bc


def gemm(a, b, out, *, n, k, precise=False, conv=None, filt=None, b_row_stride=0, a_koff=0, out_dtype=None, ldc=None,
         bias=None, act=L.ACT_NONE, alpha=1.0, residual=None, preact_out=None, dact_src=None, dact_kind=L.ACT_NONE,
         dact_split=False, row_bias=None, row_bias_div=1, accumulate=False, out_mode=L.OUT_LINEAR, out_hw=None,
         block_n=0, m=None, lda=None, a_cols=None):
    """out = epilogue(sum_taps A_t @ B_t^T).

    a     : bf16 tensor; 2-D [M, lda] or (conv=(nb,h,w)) NHWC [nb,h,w,lda]; split (hi|lo) when precise
    b     : bf16 weights [b_rows, K] (or [b_rows, 2K] when precise)
    filt  : list of (dy, dx) pixel offsets for conv taps; tap t uses weight rows [t*b_row_stride, ...)
    k     : logical contraction length per tap;   a_koff: logical column offset into A
    """
    d = L.GemmDesc()
    d.a = a.data_ptr()
    d.lda = lda if lda is not None else a.shape[-1]
    d.a_cols = a_cols if a_cols is not None else 0
    if conv is not None:
        d.a_conv = 1
        d.nb, d.h, d.w = conv
        d.m = conv[0] * conv[1] * conv[2]
    else:
        d.a_conv = 0
        d.m = m if m is not None else a.numel() // a.shape[-1]
    d.b = b.data_ptr()
    d.b_rows = b.shape[0]
    d.ldb = b.shape[1]
    d.n, d.k_per_tap = n, k
    kl_a = d.lda // 2 if precise else 0      # lo offset of A
    kl_b = d.ldb // 2 if precise else 0
    taps = []
    for t, (dy, dx) in enumerate(filt if filt is not None else [(0, 0)]):
        br = t * b_row_stride
        taps.append((dy, dx, a_koff, br, 0))
        if precise:
            taps.append((dy, dx, a_koff, br, kl_b))
            taps.append((dy, dx, a_koff + kl_a, br, 0))
    _set_taps(d, taps)
    d.out = out.data_ptr()
    d.out_dtype = out_dtype if out_dtype is not None else L.dtype_of(out)
    d.ldc = ldc if ldc is not None else out.shape[-1]
    d.out_mode = out_mode
    if out_hw is not None:
        d.out_h, d.out_w = out_hw
    d.alpha = alpha
    d.bias = bias.data_ptr() if bias is not None else None
    if row_bias is not None:
        d.row_bias, d.row_bias_div, d.row_bias_ld = row_bias.data_ptr(), row_bias_div, row_bias.shape[-1]
    d.act = act
    if preact_out is not None:
        d.preact_out, d.preact_dtype, d.ld_preact = preact_out.data_ptr(), L.dtype_of(preact_out), preact_out.shape[-1]
    if dact_src is not None:
        d.dact_src, d.dact_dtype, d.dact_kind, d.ld_dact = dact_src.data_ptr(), L.dtype_of(dact_src, dact_split), dact_kind, dact_src.shape[-1]
    if residual is not None:
        d.residual, d.res_dtype, d.ldres = residual.data_ptr(), L.dtype_of(residual), residual.shape[-1]
    d.accumulate = 1 if accumulate else 0
    d.block_n = block_n
    L.raw_gemm(d)
    return out


def wgrad(dy, x, dw, *, m, n, precise=False, conv=None, filt=None, dy_koff=0, x_koff=0, ld_dw=None, slot_stride=None,
          alpha=1.0, splits=0, rows=None, ld_dy=None, ld_x=None):
    """dw[slot, i, j] += alpha * sum_rows dy[row, dy_koff + i] * x[row shifted by filt[slot], x_koff + j]   (fp32 dw, atomically reduced).

    dy, x: bf16 [rows, ld] (or NHWC with conv=(nb,h,w)); split (hi|lo) when precise.  One output slot per filter position.
    """
    d = L.WgradDesc()
    d.dy, d.x = dy.data_ptr(), x.data_ptr()
    d.ld_dy = ld_dy if ld_dy is not None else dy.shape[-1]
    d.ld_x = ld_x if ld_x is not None else x.shape[-1]
    if conv is not None:
        d.conv = 1
        d.nb, d.h, d.w = conv
        d.rows = conv[0] * conv[1] * conv[2]
    else:
        d.rows = rows if rows is not None else dy.numel() // dy.shape[-1]
    d.m, d.n = m, n
    lo_dy = d.ld_dy // 2 if precise else 0
    lo_x = d.ld_x // 2 if precise else 0
    taps = []
    for s, (fy, fx) in enumerate(filt if filt is not None else [(0, 0)]):
        taps.append((fy, fx, dy_koff, x_koff, s))
        if precise:
            taps.append((fy, fx, dy_koff, x_koff + lo_x, s))
            taps.append((fy, fx, dy_koff + lo_dy, x_koff, s))
    assert len(taps) <= L.MAX_TAPS
    d.num_taps = len(taps)
    for i, (fy, fx, ka, kx, s) in enumerate(taps):
        d.tap_dy[i], d.tap_dx[i], d.tap_dy_koff[i], d.tap_x_koff[i], d.tap_slot[i] = fy, fx, ka, kx, s
    d.dw = dw.data_ptr()
    d.ld_dw = ld_dw if ld_dw is not None else dw.shape[-1]
    d.slot_stride = slot_stride if slot_stride is not None else (dw.shape[-2] * dw.shape[-1] if dw.dim() >= 2 else 0)
    d.alpha = alpha
    d.splits = splits
    L.raw_wgrad(d)
    return dw
