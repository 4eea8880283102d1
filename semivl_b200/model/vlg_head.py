"""`VLGHead` decode head: the reference's registry type name, constructor keywords, parameter names and forward contract
(model/decode_heads/vlg_head.py:140-251) over the B200 head engine.  The torch modules are parameter containers; the
arithmetic runs in semivl_b200.engine.head through the C ABI."""
import torch
import torch.nn as nn

from .. import lib as L
from ..engine.head import HeadCfg, HeadEngine
from ..registry import HEADS
from .maskclip_vit import EncoderLayerParams


def _gn_conv(cin, cout, k, **kw):
    return nn.Sequential(nn.Conv2d(cin, cout, k, bias=False, **kw), nn.GroupNorm(cout // 16, cout), nn.ReLU(True))


class _ASPPPooling(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.gap = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(cin, cout, 1, bias=False), nn.GroupNorm(cout // 16, cout), nn.ReLU(True))


class _ASPP(nn.Module):
    def __init__(self, c, rates=(1, 6, 12, 18)):
        super().__init__()
        self.aspp_convs = nn.ModuleList([_gn_conv(c, c, 1 if d == 1 else 3, padding=0 if d == 1 else d, dilation=d) for d in rates])
        self.aspp_convs.append(_ASPPPooling(c, c))
        self.project = _gn_conv(c * (len(rates) + 1), c, 1)


class _SemanticTransformer(nn.Module):
    def __init__(self, channels, text_channels, eps=1e-5):
        super().__init__()
        self.transformer = EncoderLayerParams(channels + text_channels, 4 * channels, eps)


class _Up(nn.Module):
    def __init__(self, cin, cout, cskip):
        super().__init__()
        self.up = nn.ConvTranspose2d(cin, cin - cskip, kernel_size=2, stride=2)
        self.conv = nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.GroupNorm(cout // 16, cout), nn.ReLU(True),
                                  nn.Conv2d(cout, cout, 3, padding=1, bias=False), nn.GroupNorm(cout // 16, cout), nn.ReLU(True))


class _HeadFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, text, nfeat, grad_mode, names, *tensors):
        """tensors = (*pyramid, *conv_feats, *params); nfeat = (len(pyramid), len(conv_feats))"""
        npyr, nconv = nfeat
        nfeat = npyr + nconv
        feats, params = tensors[:nfeat], tensors[nfeat:]
        p = {k: v.detach() for k, v in zip(names, params)}
        need_grad = grad_mode and any(t.requires_grad for t in tensors)
        fin = [f.detach().permute(0, 2, 3, 1).contiguous().float() for f in feats]           # NHWC; no copy for the backbone's own outputs
        low, hctx = module.engine.forward(fin[:npyr], text, p, need_grad=need_grad, conv_feats=fin[npyr:] if nconv else None)
        ctx.module, ctx.names, ctx.hctx, ctx.nfeat = module, names, hctx, nfeat
        ctx.req = [t.requires_grad for t in tensors]
        ctx.save_for_backward(*params)
        return low

    @staticmethod
    def backward(ctx, d_low):
        params = ctx.saved_tensors
        p = {k: v.detach() for k, v in zip(ctx.names, params)}
        grads = {k: torch.zeros_like(v) for k, v in p.items()}
        dfe = ctx.module.engine.backward(ctx.hctx, d_low.contiguous().float(), p, grads, need_feat_grads=any(ctx.req[:ctx.nfeat]))
        ctx.hctx = None
        dfeat = tuple(None if (d is None or not r) else d.permute(0, 3, 1, 2) for d, r in zip(dfe, ctx.req[:ctx.nfeat]))
        dpar = tuple(grads[k] if r else None for k, r in zip(ctx.names, ctx.req[ctx.nfeat:]))
        return (None, None, None, None, None) + dfeat + dpar


class _UpsampleFunction(torch.autograd.Function):
    """F.interpolate(x, size, mode='bilinear', align_corners=False) (vlg_head.py:247-248) on the hand-written kernels."""

    @staticmethod
    def forward(ctx, low, H, W):
        low = low.contiguous().float()
        R, N, hl, wl = low.shape
        out = torch.empty(R, N, H, W, device=low.device, dtype=torch.float32)
        L.call("svl_upsample_bilinear", low, out, R * N, hl, wl, H, W)
        ctx.shape = (R, N, hl, wl, H, W)
        return out

    @staticmethod
    def backward(ctx, dout):
        R, N, hl, wl, H, W = ctx.shape
        dlow = torch.zeros(R, N, hl, wl, device=dout.device, dtype=torch.float32)
        L.call("svl_upsample_bilinear_bwd", dout.contiguous().float(), dlow, R * N, hl, wl, H, W)
        return dlow, None, None


def upsample_bilinear(low, size):
    return _UpsampleFunction.apply(low, int(size[0]), int(size[1]))


@HEADS.register_module()
class VLGHead(nn.Module):
    def __init__(self, img_size, num_classes, text_in_channels, text_channels, up_channels, skip_in_channels, skip_channels,
                 skip_from_conv_feat, num_layers, num_heads, channels, pool_size, conv1_ksize, loss_decode, align_corners, precise=False):
        super().__init__()
        assert loss_decode is None
        assert not align_corners, "semivl_b200 implements align_corners=False for the final resize (the SemiVL configuration)"
        assert pool_size is not None and pool_size[0] == pool_size[1]
        assert channels + text_channels == num_heads * 64, "attention kernels are specialised for head_dim 64"
        assert len(skip_in_channels) == 2 and len(skip_channels) == 2 and len(up_channels) == 2
        assert all(c % 16 == 0 for c in skip_in_channels), "skip inputs are GEMM operands: channel counts must be multiples of 16"
        self.image_size, self.num_classes, self.align_corners = img_size, num_classes, align_corners
        self.text_in_channels, self.num_layers, self.channels, self.skip_from_conv_feat = text_in_channels, num_layers, channels, skip_from_conv_feat
        self.load_text_embedding = None
        self.conv1 = nn.Conv2d(1, channels, kernel_size=conv1_ksize, stride=1, padding=(conv1_ksize - 1) // 2)
        self.aspp = _ASPP(channels)
        self.layers = nn.ModuleList([_SemanticTransformer(channels, text_channels) for _ in range(num_layers)])
        self.text_proj = nn.Sequential(nn.Linear(text_in_channels, text_channels), nn.ReLU())
        self.skip_proj = nn.ModuleList([nn.Sequential(nn.Conv2d(sic, sc, 3, 1, 1), nn.ReLU()) for sic, sc in zip(skip_in_channels, skip_channels)])
        self.up1 = _Up(channels, up_channels[0], skip_channels[0])
        self.up2 = _Up(up_channels[0], up_channels[1], skip_channels[1])
        self.head = nn.Conv2d(up_channels[1], 1, kernel_size=3, stride=1, padding=1)
        self.engine = HeadEngine(HeadCfg(channels=channels, text_channels=text_channels, up_channels=up_channels, skip_channels=skip_channels,
                                         num_layers=num_layers, num_heads=num_heads, pool=pool_size[0], conv1_ksize=conv1_ksize,
                                         in_dim=text_in_channels, skip_dim=skip_in_channels[0], align_corners=align_corners), precise=precise)

    def init_weights(self):
        pass          # PyTorch default initialisation of the container modules == the reference's (plain nn modules, no init_cfg)

    def set_precise(self, precise):
        self.engine.precise = bool(precise)

    def forward_lowres(self, inputs):
        """[B, N, 4h, 4w] class maps (vlg_head.py:192-244)."""
        pyramid = inputs[0][0]
        text = inputs[1]
        assert text.shape[0] == self.num_classes and text.shape[1] == self.text_in_channels, \
            "concept-expanded text tables for the head are not implemented (SemiVL uses them for the MaskCLIP guidance only)"
        names = tuple(n for n, _ in self.named_parameters())
        params = tuple(p for _, p in self.named_parameters())
        conv_feats = ()
        if self.skip_from_conv_feat:                       # vlg_head.py:196-205: the conv encoder's features close the skip list
            conv_feats = tuple(inputs[2])
        assert len(pyramid) - 1 + len(conv_feats) == len(self.skip_proj), "number of skip features != number of skip projections"
        return _HeadFunction.apply(self, text, (len(pyramid), len(conv_feats)), torch.is_grad_enabled(), names, *pyramid, *conv_feats, *params)

    def forward(self, inputs, force_output_pred_masks=False):
        low = self.forward_lowres(inputs)
        if force_output_pred_masks:
            return {"pred_masks": upsample_bilinear(low, (self.image_size, self.image_size))}
        return low
