"""Conv encoder of the Cityscapes skr04 model: mmseg `ResNetV1c(depth=101, num_stages=1, out_indices=[0], strides=[1], dilations=[1],
style='pytorch', norm_cfg=SyncBN)` -- deep stem + layer1 -- forward and hand-scheduled backward over the C-ABI kernels
(configs/_base_/models/vlm-vlg-aspp-s2p4-skr04-ftap-mcvitb.py:50-60; model/vlm.py:50-52,120-121).

Layout: NHWC activations [B*H*W, C]; convolutions on the tcgen05 engine (3x3 as implicit GEMMs, 1x1 as plain GEMMs, the 3 -> 32 stride-2
stem convolution through svl_stem_im2col with K = 27 padded to 32); GEMM operands bf16 (fast) or split bf16 pairs (precise); raw
convolution outputs that feed a BatchNorm are kept (bf16 / fp32 in precise mode) next to the normalised operand: the backward needs both.

SyncBN: in training mode the per-channel sums of a BatchNorm are all-reduced over the ranks before mean / rstd are formed, in the backward
pass the sums of dy and dy * xhat likewise -- `sync(t)` is that hook (an in-place SUM all-reduce of a [2, C] fp32 tensor plus the row
count); it is the identity for one process.  Parameter names follow mmseg's state dict (`stem.0.weight`, `stem.1.{weight,bias,
running_mean,running_var}`, `layer1.0.conv1.weight`, `layer1.0.bn1.*`, `layer1.0.downsample.{0,1}.*`), i.e. what
`pretrained/resnet101_v1c-e67eebb6.pth` holds.
"""
import torch

from .. import lib as L
from .. import ops
from .vit import WeightCache

STEM = ((0, 1, 3, 32, 2), (3, 4, 32, 32, 1), (6, 7, 32, 64, 1))        # (conv index, bn index, cin, cout, stride)
_F3 = [(i - 1, j - 1) for i in range(3) for j in range(3)]


def no_sync(t, count):
    return count


class ConvEncEngine:
    def __init__(self, precise=False, blocks=3, eps=1e-5, momentum=0.1):
        self.precise, self.blocks, self.eps, self.momentum = precise, blocks, eps, momentum
        self.cache = WeightCache()

    # ------------------------------------------------------------------ weights
    def _w(self, p, name, kind):
        pr = self.precise

        def make(t):
            t = t.float()
            if kind == "stem0":            # [32,3,3,3] -> [32, 27 -> 32]
                w = torch.zeros(t.shape[0], 32, device=t.device)
                w[:, :27] = t.reshape(t.shape[0], -1)
            elif kind == "lin":            # 1x1 conv [Co,Ci,1,1] -> [Co, Ci]
                w = t.reshape(t.shape[0], -1)
            elif kind == "lin_t":
                w = t.reshape(t.shape[0], -1).t()
            elif kind == "conv":           # [Co,Ci,3,3] -> [(tap, co), ci]
                w = t.permute(2, 3, 0, 1).reshape(-1, t.shape[1])
            elif kind == "conv_t":         # -> [(tap, ci), co]
                w = t.permute(2, 3, 1, 0).reshape(-1, t.shape[0])
            else:
                raise ValueError(kind)
            return ops.prep_weight(w.contiguous(), pr)
        return self.cache.get((name, kind, pr), p[name], make)

    # ------------------------------------------------------------------ BatchNorm pieces
    def _bn_fwd(self, raw, rows, C, p, bn, training, sync, out, relu=True, res=None, update_running=True):
        """out <- [relu](BN(raw) [+ res]); returns (mean, rstd, global row count) for the backward (None in eval mode)."""
        dev = raw.device
        adt = ops.act_dtype(self.precise)
        gamma, beta = p[bn + "weight"], p[bn + "bias"]
        if training:
            sums = torch.empty(2, C, device=dev, dtype=torch.float32)
            ws = torch.empty(L.lib().svl_bn_workspace(rows, C), device=dev, dtype=torch.float32)
            L.call("svl_bn_stats", raw, L.dtype_of(raw), raw.shape[-1], rows, C, ws, sums, n_launch=2)
            count = sync(sums, rows)
            mean = torch.empty(C, device=dev, dtype=torch.float32)
            rstd = torch.empty(C, device=dev, dtype=torch.float32)
            rm = p.get(bn + "running_mean") if update_running else None
            rv = p.get(bn + "running_var") if update_running else None
            L.call("svl_bn_finalize", sums, float(count), self.eps, self.momentum, mean, rstd, rm, rv, C)
        else:
            mean, count = p[bn + "running_mean"], rows
            rstd = torch.rsqrt(p[bn + "running_var"].float() + self.eps)
        L.call("svl_bn_apply", raw, L.dtype_of(raw), raw.shape[-1], mean, rstd, gamma, beta, res, adt, res.shape[-1] if res is not None else 0,
               out, adt, out.shape[-1], 1 if relu else 0, rows, C)
        return (mean, rstd, count) if training else None

    def _bn_bwd(self, dy, dy_dtype, raw, y, stats, rows, C, p, bn, grads, sync, dx, dx_dtype, dres=None, dres_dtype=L.F32):
        """Backward of [relu](BN(raw) [+ res]): dx <- d raw; dres <- masked dy (gradient of the residual branch); dgamma / dbeta accumulated."""
        dev = raw.device
        mean, rstd, count = stats
        adt = ops.act_dtype(self.precise)
        sums = torch.empty(2, C, device=dev, dtype=torch.float32)
        ws = torch.empty(L.lib().svl_bn_workspace(rows, C), device=dev, dtype=torch.float32)
        L.call("svl_bn_bwd_stats", dy, dy_dtype, dy.shape[-1], raw, L.dtype_of(raw), raw.shape[-1], y, adt, y.shape[-1] if y is not None else 0,
               mean, rstd, rows, C, ws, sums, n_launch=2)
        # parameter gradients use the LOCAL sums (the data-parallel exchange of the flat gradient buffer sums them over the ranks)
        ops.axpy(grads[bn + "bias"], sums[0])
        ops.axpy(grads[bn + "weight"], sums[1])
        sync(sums, rows)
        L.call("svl_bn_bwd_apply", dy, dy_dtype, dy.shape[-1], raw, L.dtype_of(raw), raw.shape[-1], y, adt, y.shape[-1] if y is not None else 0,
               mean, rstd, p[bn + "weight"], sums, float(count), dx, dx_dtype, dx.shape[-1], dres, dres_dtype,
               dres.shape[-1] if dres is not None else 0, rows, C)

    def _raw(self, rows, C, dev):
        return torch.empty(rows, C, device=dev, dtype=torch.float32 if self.precise else torch.bfloat16)

    # ------------------------------------------------------------------ forward
    def forward(self, img, p, training=True, need_grad=True, sync=no_sync):
        """img f32 [B,3,H,W] -> (feature f32 NHWC [B, H/4, W/4, 256], ctx).  `training` selects batch statistics (and updates the running
        ones) as nn.BatchNorm2d / SyncBatchNorm do; `need_grad` keeps what the backward needs."""
        pr = self.precise
        adt = ops.act_dtype(pr)
        dev = img.device
        img = img.contiguous().float()
        B, _, H, W = img.shape
        H1, W1 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        rows1 = B * H1 * W1
        keep = training and need_grad
        ctx = dict(B=B, H1=H1, W1=W1, stem=[]) if keep else None
        # ---- deep stem: conv3x3 s2 (3->32), conv3x3 (32->32), conv3x3 (32->64), each + BN + ReLU
        col = ops.new_act(rows1, 32, pr, dev)
        L.call("svl_stem_im2col", img, col, adt, col.shape[-1], B, H, W, H1, W1)
        x = col
        for li, (ci, bi, cin, cout, _) in enumerate(STEM):
            raw = self._raw(rows1, cout, dev)
            if li == 0:
                ops.gemm(x, self._w(p, "stem.0.weight", "stem0"), raw, n=cout, k=32, precise=pr)
            else:
                ops.gemm(x, self._w(p, f"stem.{ci}.weight", "conv"), raw, n=cout, k=cin, precise=pr, conv=(B, H1, W1), filt=_F3, b_row_stride=cout)
            y = ops.new_act(rows1, cout, pr, dev)
            st = self._bn_fwd(raw, rows1, cout, p, f"stem.{bi}.", training, sync, y)
            if keep:
                ctx["stem"].append(dict(x=x, raw=raw, y=y, stats=st))
            x = y
        # ---- max-pool 3x3 s2
        H2, W2 = (H1 - 1) // 2 + 1, (W1 - 1) // 2 + 1
        rows2 = B * H2 * W2
        pooled = ops.new_act(rows2, 64, pr, dev)
        widx = torch.empty(rows2, 64, device=dev, dtype=torch.uint8)
        L.call("svl_maxpool3s2_fwd", x, adt, x.shape[-1], pooled, adt, pooled.shape[-1], widx, B, H1, W1, 64, H2, W2)
        if keep:
            ctx.update(H2=H2, W2=W2, widx=widx, blocks=[])
        # ---- layer1: Bottlenecks 1x1 (->64) / 3x3 (64) / 1x1 (->256), BN after each, ReLU after the first two and after the sum
        x = pooled
        cin = 64
        for i in range(self.blocks):
            b = f"layer1.{i}."
            r1 = self._raw(rows2, 64, dev)
            ops.gemm(x, self._w(p, b + "conv1.weight", "lin"), r1, n=64, k=cin, precise=pr)
            y1 = ops.new_act(rows2, 64, pr, dev)
            s1 = self._bn_fwd(r1, rows2, 64, p, b + "bn1.", training, sync, y1)
            r2 = self._raw(rows2, 64, dev)
            ops.gemm(y1, self._w(p, b + "conv2.weight", "conv"), r2, n=64, k=64, precise=pr, conv=(B, H2, W2), filt=_F3, b_row_stride=64)
            y2 = ops.new_act(rows2, 64, pr, dev)
            s2 = self._bn_fwd(r2, rows2, 64, p, b + "bn2.", training, sync, y2)
            r3 = self._raw(rows2, 256, dev)
            ops.gemm(y2, self._w(p, b + "conv3.weight", "lin"), r3, n=256, k=64, precise=pr)
            has_ds = (b + "downsample.0.weight") in p
            rd = sd = None
            if has_ds:
                rd = self._raw(rows2, 256, dev)
                ops.gemm(x, self._w(p, b + "downsample.0.weight", "lin"), rd, n=256, k=cin, precise=pr)
                ident = ops.new_act(rows2, 256, pr, dev)
                sd = self._bn_fwd(rd, rows2, 256, p, b + "downsample.1.", training, sync, ident, relu=False)
            else:
                ident = x
            out = ops.new_act(rows2, 256, pr, dev)
            s3 = self._bn_fwd(r3, rows2, 256, p, b + "bn3.", training, sync, out, relu=True, res=ident)
            if keep:
                ctx["blocks"].append(dict(x=x, cin=cin, r1=r1, y1=y1, s1=s1, r2=r2, y2=y2, s2=s2, r3=r3, s3=s3, rd=rd, sd=sd, out=out, has_ds=has_ds))
            x, cin = out, 256
        feat = torch.empty(B, H2, W2, 256, device=dev, dtype=torch.float32)
        ops.cast(x, adt, feat.view(rows2, 256), L.F32, rows2, 256)
        return feat, ctx

    # ------------------------------------------------------------------ backward
    def backward(self, ctx, dfeat, p, grads, sync=no_sync):
        """dfeat: f32 NHWC [B, H/4, W/4, 256].  Accumulates parameter gradients into `grads` (name -> param-shaped f32 tensor)."""
        pr = self.precise
        adt = ops.act_dtype(pr)
        gdt = L.F32 if pr else L.BF16
        gtorch = torch.float32 if pr else torch.bfloat16
        dev = dfeat.device
        B, H1, W1, H2, W2 = ctx["B"], ctx["H1"], ctx["W1"], ctx["H2"], ctx["W2"]
        rows1, rows2 = B * H1 * W1, B * H2 * W2
        f32 = dict(device=dev, dtype=torch.float32)

        def wgrad_lin(dy_act, x_act, name, cout, cin):
            ops.wgrad(dy_act, x_act, grads[name].view(cout, cin), m=cout, n=cin, precise=pr)

        def wgrad_conv(dy_act, x_act, name, cout, cin, geo):
            dw = torch.zeros(9, cout, cin, **f32)
            ops.wgrad(dy_act, x_act, dw, m=cout, n=cin, precise=pr, conv=geo, filt=_F3)
            grads[name].add_(dw.view(3, 3, cout, cin).permute(2, 3, 0, 1))

        d_out, d_out_dtype = dfeat.contiguous().view(rows2, 256), L.F32
        for i in reversed(range(self.blocks)):
            S = ctx["blocks"][i]
            b = f"layer1.{i}."
            cin = S["cin"]
            # out = relu(bn3(r3) + ident): d r3 (operand) and the masked gradient of the identity branch
            d_r3 = ops.new_act(rows2, 256, pr, dev)
            d_ident = torch.empty(rows2, 256, device=dev, dtype=gtorch)
            self._bn_bwd(d_out, d_out_dtype, S["r3"], S["out"], S["s3"], rows2, 256, p, b + "bn3.", grads, sync, d_r3, adt, dres=d_ident, dres_dtype=gdt)
            wgrad_lin(d_r3, S["y2"], b + "conv3.weight", 256, 64)
            d_y2 = torch.empty(rows2, 64, device=dev, dtype=gtorch)
            ops.gemm(d_r3, self._w(p, b + "conv3.weight", "lin_t"), d_y2, n=64, k=256, precise=pr)
            d_r2 = ops.new_act(rows2, 64, pr, dev)
            self._bn_bwd(d_y2, gdt, S["r2"], S["y2"], S["s2"], rows2, 64, p, b + "bn2.", grads, sync, d_r2, adt)
            wgrad_conv(d_r2, S["y1"], b + "conv2.weight", 64, 64, (B, H2, W2))
            d_y1 = torch.empty(rows2, 64, device=dev, dtype=gtorch)
            ops.gemm(d_r2, self._w(p, b + "conv2.weight", "conv_t"), d_y1, n=64, k=64, precise=pr, conv=(B, H2, W2), filt=[(-a, -c) for a, c in _F3],
                     b_row_stride=64)
            d_r1 = ops.new_act(rows2, 64, pr, dev)
            self._bn_bwd(d_y1, gdt, S["r1"], S["y1"], S["s1"], rows2, 64, p, b + "bn1.", grads, sync, d_r1, adt)
            wgrad_lin(d_r1, S["x"], b + "conv1.weight", 64, cin)
            d_x = torch.empty(rows2, cin, **f32)                      # fp32 accumulator of the two branches
            ops.gemm(d_r1, self._w(p, b + "conv1.weight", "lin_t"), d_x, n=cin, k=64, precise=pr)
            if S["has_ds"]:
                d_rd = ops.new_act(rows2, 256, pr, dev)
                self._bn_bwd(d_ident, gdt, S["rd"], None, S["sd"], rows2, 256, p, b + "downsample.1.", grads, sync, d_rd, adt)
                wgrad_lin(d_rd, S["x"], b + "downsample.0.weight", 256, cin)
                ops.gemm(d_rd, self._w(p, b + "downsample.0.weight", "lin_t"), d_x, n=cin, k=256, precise=pr, accumulate=True)
            else:
                d_x.add_(d_ident.float())
            d_out, d_out_dtype = d_x, L.F32
        # max-pool: gather over the windows that selected each stem pixel
        d_stem = torch.empty(rows1, 64, device=dev, dtype=gtorch)
        L.call("svl_maxpool3s2_bwd", d_out, L.F32, 64, ctx["widx"], d_stem, gdt, 64, B, H1, W1, 64, H2, W2)
        dy, dy_dtype = d_stem, gdt
        for li in reversed(range(3)):
            ci, bi, cin, cout, _ = STEM[li]
            S = ctx["stem"][li]
            d_raw = ops.new_act(rows1, cout, pr, dev)
            self._bn_bwd(dy, dy_dtype, S["raw"], S["y"], S["stats"], rows1, cout, p, f"stem.{bi}.", grads, sync, d_raw, adt)
            if li == 0:
                dw = torch.zeros(cout, 32, **f32)
                ops.wgrad(d_raw, S["x"], dw, m=cout, n=32, precise=pr)
                grads["stem.0.weight"].add_(dw[:, :27].reshape(cout, 3, 3, 3))
            else:
                wgrad_conv(d_raw, S["x"], f"stem.{ci}.weight", cout, cin, (B, H1, W1))
                dy = torch.empty(rows1, cin, device=dev, dtype=gtorch)
                ops.gemm(d_raw, self._w(p, f"stem.{ci}.weight", "conv_t"), dy, n=cin, k=cout, precise=pr, conv=(B, H1, W1),
                         filt=[(-a, -c) for a, c in _F3], b_row_stride=cin)
                dy_dtype = gdt
        return grads
