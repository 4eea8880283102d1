"""Host-side launch helpers over the C ABI: they fill the descriptors of include/semivl_b200.h from torch tensors.

Storage conventions (see svl_dtype):
  * fast mode    : GEMM operands are bf16 [rows, C]
  * precise mode : GEMM operands are split bf16 pairs [rows, 2C] (hi | lo) -> three tensor-core taps per contraction
A `Mat` is just (tensor, split flag); weights are prepared once per step by `prep_weight`.
"""
import ctypes as C

import torch

from . import lib as L


PROFILE = None      # set to a list to record (start_event, end_event, flops) around every contraction launch (bench.py roofline)


def _profiled(fn, flops, label=None):
    if PROFILE is None:
        fn()
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    PROFILE.append((e0, e1, flops, label))


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
         block_n=0, m=None, lda=None, a_cols=None, b_col0=0, a_map_w=0, taps=None, a_lo=None, gn_stats=None):
    """out = epilogue(sum_taps A_t @ B_t^T).

    a     : bf16 tensor; 2-D [M, lda] or (conv=(nb,h,w)) NHWC [nb,h,w,lda]; split (hi|lo) when precise
    b     : bf16 weights [b_rows, K] (or [b_rows, 2K] when precise)
    filt  : list of (dy, dx) pixel offsets for conv taps; tap t uses weight rows [t*b_row_stride, ...)
    k     : logical contraction length per tap;   a_koff: logical column offset into A
    gn_stats: optional dict {'maps': .., 'G': ..}: when the library can produce the GroupNorm statistics of the output in this launch's epilogue
            (svl_conv_gn_splits > 0) the dict receives 'ws' (partial sums) and 'splits' for gn_relu_fwd(stats=...)
    """
    d = L.GemmDesc()
    d.a = a.data_ptr()
    d.lda = lda if lda is not None else a.shape[-1]
    d.a_cols = a_cols if a_cols is not None else 0
    if conv is not None:
        d.a_conv = 1
        d.nb, d.h, d.w = conv
        d.a_map_w = a_map_w
        d.m = conv[0] * conv[1] * conv[2]
    else:
        d.a_conv = 0
        d.m = m if m is not None else a.numel() // a.shape[-1]
    d.b = b.data_ptr()
    d.b_rows = b.shape[0]
    d.ldb = b.shape[1]
    d.n, d.k_per_tap = n, k
    kl_a = (a_lo if a_lo is not None else d.lda // 2) if precise else 0      # lo offset of A
    kl_b = d.ldb // 2 if precise else 0
    if taps is None:
        taps = [(dy, dx, a_koff, t * b_row_stride, b_col0) for t, (dy, dx) in enumerate(filt if filt is not None else [(0, 0)])]
    full = []
    for (dy, dx, ak, br, bc) in taps:          # logical taps: (pixel dy, dx, A column offset, B row offset, B column offset)
        full.append((dy, dx, ak, br, bc))
        if precise:
            full.append((dy, dx, ak, br, bc + kl_b))
            full.append((dy, dx, ak + kl_a, br, bc))
    _set_taps(d, full)
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
    if gn_stats is not None and not precise and n == 16 * gn_stats["G"]:
        splits = L.lib().svl_conv_gn_splits(C.byref(d))
        if splits > 0:
            gn_stats["ws"] = torch.empty(gn_stats["maps"] * splits * gn_stats["G"] * 2, device=out.device, dtype=torch.float32)
            gn_stats["splits"] = splits
            d.gn_part = gn_stats["ws"].data_ptr()
    _profiled(lambda: L.raw_gemm(d), 2.0 * d.m * n * k * len(taps), f"gemm m{d.m} n{n} k{k} taps{len(taps)} conv{d.a_conv} out{d.out_dtype} mode{d.out_mode}")
    return out


def wgrad(dy, x, dw, *, m, n, precise=False, conv=None, filt=None, dy_koff=0, x_koff=0, ld_dw=None, slot_stride=None,
          alpha=1.0, splits=0, rows=None, ld_dy=None, ld_x=None, x_map_w=0, x_lo=None, dy_lo=None, taps=None):
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
    d.x_map_w = x_map_w
    lo_dy = (dy_lo if dy_lo is not None else d.ld_dy // 2) if precise else 0
    lo_x = (x_lo if x_lo is not None else d.ld_x // 2) if precise else 0
    if taps is None:        # logical taps: (pixel dy, dx, dy column offset, x column offset, slot)
        taps = [(fy, fx, dy_koff, x_koff, s) for s, (fy, fx) in enumerate(filt if filt is not None else [(0, 0)])]
    full = []
    for (fy, fx, ka, kx, s) in taps:
        full.append((fy, fx, ka, kx, s))
        if precise:
            full.append((fy, fx, ka, kx + lo_x, s))
            full.append((fy, fx, ka + lo_dy, kx, s))
    taps = full
    assert len(taps) <= L.MAX_TAPS
    d.num_taps = len(taps)
    for i, (fy, fx, ka, kx, s) in enumerate(taps):
        d.tap_dy[i], d.tap_dx[i], d.tap_dy_koff[i], d.tap_x_koff[i], d.tap_slot[i] = fy, fx, ka, kx, s
    d.dw = dw.data_ptr()
    d.ld_dw = ld_dw if ld_dw is not None else dw.shape[-1]
    d.slot_stride = slot_stride if slot_stride is not None else (dw.shape[-2] * dw.shape[-1] if dw.dim() >= 2 else 0)
    d.alpha = alpha
    d.splits = splits
    nlog = len(taps) // (3 if precise else 1)
    _profiled(lambda: L.raw_wgrad(d), 2.0 * d.rows * m * n * nlog, f"wgrad rows{d.rows} m{m} n{n} taps{nlog} conv{d.conv}")
    return dw


# ---------------------------------------------------------------------------------------------- token kernels
def act_dtype(precise):
    return L.BF16X2 if precise else L.BF16


def new_act(rows, cols, precise, device="cuda"):
    """Uninitialised GEMM-operand tensor [rows, cols] (bf16) or its split form [rows, 2*cols]."""
    return torch.empty(rows, cols * (2 if precise else 1), device=device, dtype=torch.bfloat16)


def patchify(img, precise, patch=16):
    b, _, H, W = img.shape
    hp, wp = -(-H // patch), -(-W // patch)
    out = new_act(b * hp * wp, 3 * patch * patch, precise)
    L.call("svl_patchify", img, out, act_dtype(precise), b, H, W, patch, hp, wp)
    return out, hp, wp


def assemble_tokens(patches, cls, pos, b, hw):
    c = patches.shape[-1]
    x = torch.empty(b, hw + 1, c, device=patches.device, dtype=torch.float32)
    L.call("svl_assemble_tokens", patches, cls, pos, x, b, hw, c)
    return x


def layernorm_fwd(x2d, gamma, beta, eps, *, out=None, out_dtype=None, precise=False, save_stats=True, ldx=None):
    """x2d: f32 [rows, c] (row stride ldx).  Returns (y, mean, rstd); y is an operand tensor unless out_dtype == F32."""
    rows, c = x2d.shape
    odt = out_dtype if out_dtype is not None else act_dtype(precise)
    if out is None:
        out = torch.empty(rows, c, device=x2d.device, dtype=torch.float32) if odt == L.F32 else new_act(rows, c, odt == L.BF16X2)
    mean = torch.empty(rows, device=x2d.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty(rows, device=x2d.device, dtype=torch.float32) if save_stats else None
    L.call("svl_layernorm_fwd", x2d, ldx if ldx is not None else x2d.stride(0), gamma, beta, out, odt, out.shape[-1], mean, rstd, rows, c, eps)
    return out, mean, rstd


def layernorm_bwd(dy, dy_dtype, x2d, gamma, mean, rstd, *, dres1=None, dres2=None, want_dx=True, act_precise=None,
                  dgamma=None, dbeta=None):
    """Returns (dx f32 or None, dx_act or None)."""
    rows, c = x2d.shape
    dx = torch.empty(rows, c, device=x2d.device, dtype=torch.float32) if want_dx else None
    dx_act = new_act(rows, c, act_precise) if act_precise is not None else None
    L.call("svl_layernorm_bwd", dy, dy_dtype, dy.shape[-1], x2d, x2d.stride(0), gamma, mean, rstd, dres1, dres2, dx, dx_act,
           act_dtype(bool(act_precise)), dx_act.shape[-1] if dx_act is not None else 0, dgamma, dbeta, rows, c)
    return dx, dx_act


def l2norm_fwd(x2d, *, want_f32=True, act_precise=None, eps=1e-12, ldx=None):
    rows, c = x2d.shape
    y = torch.empty(rows, c, device=x2d.device, dtype=torch.float32) if want_f32 else None
    y_act = new_act(rows, c, act_precise) if act_precise is not None else None
    inv = torch.empty(rows, device=x2d.device, dtype=torch.float32)
    L.call("svl_l2norm_fwd", x2d, ldx if ldx is not None else x2d.stride(0), y, y_act, act_dtype(bool(act_precise)),
           y_act.shape[-1] if y_act is not None else 0, inv, rows, c, eps)
    return y, y_act, inv


def l2norm_bwd(dy, dy_dtype, y, inv, dx, *, accumulate=False, lddy=None):
    rows, c = y.shape
    L.call("svl_l2norm_bwd", dy, dy_dtype, lddy if lddy is not None else dy.shape[-1], y, inv, dx, dx.stride(0), 1 if accumulate else 0, rows, c)
    return dx


def cast(src, src_dtype, dst, dst_dtype, rows, cols, *, ld_src=None, ld_dst=None, scale=1.0, batch=1, src_batch_stride=0,
         dst_batch_stride=0, src_offset=0, dst_offset=0):
    """dst = scale * src, changing the storage type; `batch` blocks of `rows` rows (block strides / element offsets optional)."""
    esz_s, esz_d = src.element_size(), dst.element_size()
    L.call("svl_cast", src.data_ptr() + src_offset * esz_s, src_dtype, ld_src if ld_src is not None else src.shape[-1], src_batch_stride,
           dst.data_ptr() + dst_offset * esz_d, dst_dtype, ld_dst if ld_dst is not None else dst.shape[-1], dst_batch_stride, batch, rows,
           cols, scale)
    return dst


def to_act(x2d, precise):
    """f32 [rows, c] -> GEMM operand."""
    rows, c = x2d.shape
    return cast(x2d, L.F32, new_act(rows, c, precise), act_dtype(precise), rows, c, ld_src=x2d.stride(0))


def colsum(x, x_dtype, rows, cols, out, ld=None):
    L.call("svl_colsum", x, x_dtype, ld if ld is not None else x.shape[-1], rows, cols, out)
    return out


def batch_sum(x, out, accumulate=False):
    b = x.shape[0]
    L.call("svl_batch_sum", x, out, b, x.numel() // b, 1 if accumulate else 0)
    return out


def axpy(dst, src, alpha=1.0):
    L.call("svl_axpy", dst, src, alpha, dst.numel())
    return dst


def attention_fwd(qkv, b, seq, heads, precise, want_lse=True):
    """qkv operand [b*seq, 3E] (6E when precise) -> (out operand [b*seq, E], lse [b, heads, seq])."""
    E = heads * 64
    out = new_act(b * seq, E, precise, qkv.device)
    lse = torch.empty(b, heads, seq, device=qkv.device, dtype=torch.float32) if want_lse else None
    L.call("svl_attention_fwd", qkv, 1 if precise else 0, out, lse, b, seq, heads, 0.125)
    return out, lse


def attention_bwd(qkv, out, dout, lse, b, seq, heads, precise, dv_add=None, dv_add_dtype=L.F32):
    E = heads * 64
    dqkv = new_act(b * seq, 3 * E, precise, qkv.device)
    delta = torch.empty(L.lib().svl_attention_bwd_workspace(1 if precise else 0, b, seq, heads), device=qkv.device, dtype=torch.float32)
    L.call("svl_attention_bwd", qkv, out, dout, 1 if precise else 0, lse, delta, dv_add, dv_add_dtype,
           dv_add.shape[-1] if dv_add is not None else 0, dqkv, b, seq, heads, 0.125, n_launch=3)
    return dqkv


# ---------------------------------------------------------------------------------------------- head kernels
def gn_relu_fwd(x, x_dtype, gamma, beta, out, out_dtype, maps, hw, C, G, *, ldx=None, ldo=None, out_col0=0, res=None, res_dtype=L.BF16,
                ldres=None, save_stats=True, eps=1e-5, stats=None):
    """stats: the dict a producing ops.gemm(gn_stats=...) filled ('ws', 'splits'): the statistics pass is skipped."""
    dev = x.device
    mean = torch.empty(maps, G, device=dev, dtype=torch.float32)
    rstd = torch.empty(maps, G, device=dev, dtype=torch.float32)
    fused = stats is not None and "ws" in stats
    ws = stats["ws"] if fused else torch.empty(L.lib().svl_gn_workspace(maps, hw, C, G), device=dev, dtype=torch.float32)     # fixed-order partial sums
    L.call("svl_gn_relu_fwd", x, x_dtype, ldx if ldx is not None else x.shape[-1], gamma, beta,
           out.data_ptr() + out_col0 * out.element_size(), out_dtype, ldo if ldo is not None else out.shape[-1],
           res, res_dtype, (ldres if ldres is not None else (res.shape[-1] if res is not None else 0)), mean, rstd, ws, maps, hw, C, G, eps,
           stats["splits"] if fused else 0, n_launch=2 if fused else 3)
    return mean, rstd


def gn_relu_bwd(dy, dy_dtype, x, x_dtype, gamma, beta, mean, rstd, dx, dx_dtype, dgamma, dbeta, maps, hw, C, G, *, lddy=None, dy_col0=0,
                ldx=None, lddx=None):
    L.call("svl_gn_relu_bwd", dy.data_ptr() + dy_col0 * dy.element_size(), dy_dtype, lddy if lddy is not None else dy.shape[-1],
           x, x_dtype, ldx if ldx is not None else x.shape[-1], gamma, beta, mean, rstd, dx, dx_dtype,
           lddx if lddx is not None else dx.shape[-1], dgamma, dbeta,
           torch.empty(L.lib().svl_gn_workspace(maps, hw, C, G), device=dx.device, dtype=torch.float32), maps, hw, C, G, n_launch=3)
    return dx
