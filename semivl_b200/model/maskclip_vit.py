"""`MaskClipVisionTransformer` backbone: the reference's registry type name, constructor keywords, parameter names and
forward contract (third_party/maskclip/models/backbones/maskclip_vit.py:147-603) over the B200 encoder engine.

The torch modules below are parameter containers only (their names produce the reference's state_dict keys, e.g.
`layers.3.attn.attn.in_proj_weight`, `layers.3.ffn.layers.0.0.weight`, `patch_embed.projection.weight`, pinned by
third_party/maskclip/convert_clip_weights.py:27-64); all arithmetic runs in semivl_b200.engine.vit through the C ABI.
"""
import torch
import torch.nn as nn

from ..engine.vit import VitCfg, VitEngine
from ..registry import BACKBONES


class _MHAParams(nn.Module):
    """attribute path attn.attn.{in_proj_weight,in_proj_bias,out_proj.{weight,bias}} (mmcv MultiheadAttention wrapping nn.MultiheadAttention)"""

    def __init__(self, embed_dims, bias=True):
        super().__init__()
        self.attn = nn.Module()
        self.attn.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dims, embed_dims))
        self.attn.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dims))
        self.attn.out_proj = nn.Linear(embed_dims, embed_dims, bias=True)
        nn.init.xavier_uniform_(self.attn.in_proj_weight)        # nn.MultiheadAttention._reset_parameters
        nn.init.zeros_(self.attn.out_proj.bias)


class _FFNParams(nn.Module):
    """key names ffn.layers.0.0.* / ffn.layers.1.* (mmcv FFN)"""

    def __init__(self, embed_dims, hidden):
        super().__init__()
        self.layers = nn.Sequential(nn.Sequential(nn.Linear(embed_dims, hidden), nn.GELU(), nn.Dropout(0.0)), nn.Linear(hidden, embed_dims),
                                    nn.Dropout(0.0))


class EncoderLayerParams(nn.Module):
    def __init__(self, embed_dims, hidden, eps):
        super().__init__()
        self.ln1 = nn.LayerNorm(embed_dims, eps=eps)
        self.attn = _MHAParams(embed_dims)
        self.ln2 = nn.LayerNorm(embed_dims, eps=eps)
        self.ffn = _FFNParams(embed_dims, hidden)


class _PatchEmbedParams(nn.Module):
    def __init__(self, in_channels, embed_dims, patch, bias):
        super().__init__()
        self.projection = nn.Conv2d(in_channels, embed_dims, kernel_size=patch, stride=patch, bias=bias)


class _VitFunction(torch.autograd.Function):
    """Whole-encoder autograd node: forward and backward are the engine's hand-scheduled kernel sequences."""

    @staticmethod
    def forward(ctx, module, img, want_global, grad_mode, names, *params):
        p = dict(zip(names, params))
        need_grad = grad_mode and any(t.requires_grad for t in params)
        if need_grad:
            bad = [k for k, t in p.items() if t.requires_grad and not _is_trainable_name(k)]
            if bad:
                raise NotImplementedError(f"semivl_b200 computes backbone weight gradients for attn.* and pos_embed only (the SemiVL "
                                          f"fine-tuning set); freeze {bad[:3]}... or run under torch.no_grad()")
        feats, glob, ectx = module.engine.forward(img, {k: v.detach() for k, v in p.items()}, need_grad=need_grad, want_global=want_global)
        ctx.module, ctx.names, ctx.ectx = module, names, ectx
        ctx.req = [t.requires_grad for t in params]
        ctx.save_for_backward(*params)
        outs = tuple(feats) + ((glob,) if glob is not None else ())
        if glob is not None:
            ctx.mark_non_differentiable(glob)
        return outs

    @staticmethod
    def backward(ctx, *douts):
        params = ctx.saved_tensors
        p = {k: v.detach() for k, v in zip(ctx.names, params)}
        nfe = len(douts) - (1 if ctx.module._last_want_global else 0)
        dfe = [None if d is None else d.contiguous() for d in douts[:nfe]]      # the Function's outputs are NHWC already
        grads = {k: torch.zeros_like(v) for k, v in zip(ctx.names, params) if _is_trainable_name(k)}
        ctx.module.engine.backward(ctx.ectx, dfe, p, grads)
        ctx.ectx = None
        return (None, None, None, None, None) + tuple(grads[k] if (r and k in grads) else None for k, r in zip(ctx.names, ctx.req))


def _is_trainable_name(name):
    """The engine's backward produces weight gradients for exactly the tensors SemiVL fine-tunes (model/vlm.py:80-88,
    exclude_keys=['attn', 'pos_embed']); the FFN / LayerNorm / patch-embed / proj weights only get data gradients."""
    return "attn" in name or "pos_embed" in name


@BACKBONES.register_module()
class MaskClipVisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, patch_bias=True, in_channels=3, embed_dims=768, num_layers=12, num_heads=12, mlp_ratio=4,
                 out_indices=-1, qkv_bias=True, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, with_cls_token=True,
                 output_cls_token=False, norm_cfg=dict(type='LN'), act_cfg=dict(type='GELU'), patch_norm=False, pre_norm=False, final_norm=False,
                 return_clip_embed=False, return_qkv=False, skip_last_attn=False, interpolate_mode='bicubic', num_fcs=2, norm_eval=False,
                 with_cp=False, pretrained=None, init_cfg=None, precise=False, **unsupported):
        super().__init__()
        if isinstance(img_size, int):
            img_size = (img_size, img_size)
        # the B200 path implements the configuration SemiVL uses; everything else is rejected loudly
        assert pre_norm and final_norm and return_clip_embed and return_qkv and with_cls_token and not output_cls_token, \
            "semivl_b200 implements the pre_norm/final_norm/return_clip_embed/return_qkv/with_cls_token configuration only"
        assert not skip_last_attn and not patch_norm and num_fcs == 2 and qkv_bias and act_cfg.get('type') == 'GELU'
        assert drop_rate == 0.0 and attn_drop_rate == 0.0 and drop_path_rate == 0.0, "dropout inside the encoder is not supported"
        assert embed_dims == num_heads * 64, "attention kernels are specialised for head_dim 64"
        assert interpolate_mode == 'bicubic'
        self.interpolate_mode = interpolate_mode
        for k, v in unsupported.items():
            assert not v, f"unsupported backbone option {k}={v}"
        self.img_size, self.patch_size, self.pretrained, self.norm_eval = img_size, patch_size, pretrained, norm_eval
        self.embed_dims, self.num_layers = embed_dims, num_layers
        if out_indices is None:
            self.out_indices = [num_layers]
        elif isinstance(out_indices, int):
            self.out_indices = [num_layers - 1 if out_indices == -1 else out_indices]
        else:
            self.out_indices = list(out_indices)
        eps = norm_cfg.get('eps', 1e-5)
        self.patch_embed = _PatchEmbedParams(in_channels, embed_dims, patch_size, patch_bias)
        num_patches = (img_size[0] // patch_size) * (img_size[1] // patch_size)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dims))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dims))
        self.ln0 = nn.LayerNorm(embed_dims, eps=eps)
        self.layers = nn.ModuleList([EncoderLayerParams(embed_dims, mlp_ratio * embed_dims, eps) for _ in range(num_layers)])
        self.ln1 = nn.LayerNorm(embed_dims, eps=eps)
        self.proj = nn.Conv2d(embed_dims, 512, kernel_size=1, bias=False)
        self.engine = VitEngine(VitCfg(embed=embed_dims, heads=num_heads, layers=num_layers, patch=patch_size, out_indices=self.out_indices,
                                       eps=eps, proj_dim=512, img_size=img_size[0]), precise=precise)
        self._last_want_global = True

    def init_weights(self):
        """maskclip_vit.py:378-429.  `pretrained` (a converted CLIP checkpoint, see semivl_b200/convert_clip_weights.py): strip the
        'backbone.' prefix, resize the position table to this model's token grid (bicubic, cls entry kept), give the CLIP projection
        its 1x1-conv shape (or drop it when the model does not return the CLIP embedding), load non-strictly -- the reference's
        Pretrained branch.  Otherwise the random-init branch: trunc-normal(0.02) for pos/cls and Linear weights."""
        if isinstance(self.pretrained, str):
            sd = torch.load(self.pretrained, map_location='cpu')
            sd = sd.get('state_dict', sd)
            sd = {k.replace('backbone.', ''): v for k, v in sd.items()}
            if 'pos_embed' in sd and sd['pos_embed'].shape != self.pos_embed.shape:          # maskclip_vit.py:395-408
                g = int(round((sd['pos_embed'].shape[1] - 1) ** 0.5))
                hw = (self.img_size[0] // self.patch_size, self.img_size[1] // self.patch_size)
                grid = sd['pos_embed'][:, 1:].reshape(1, g, g, -1).permute(0, 3, 1, 2)
                grid = torch.nn.functional.interpolate(grid, size=hw, mode=self.interpolate_mode, align_corners=False)
                sd['pos_embed'] = torch.cat((sd['pos_embed'][:, :1], grid.flatten(2).transpose(1, 2)), dim=1)
            if 'proj.weight' in sd:
                own = dict(self.named_parameters()).get('proj.weight')
                if own is None:
                    sd.pop('proj.weight')
                elif sd['proj.weight'].dim() == 2 and own.dim() == 4:
                    sd['proj.weight'] = sd['proj.weight'][:, :, None, None]
            self.load_report = self.load_state_dict(sd, strict=False)
            return
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        for n, m in self.named_modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    if 'ffn' in n:
                        nn.init.normal_(m.bias, mean=0.0, std=1e-6)
                    else:
                        nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_in', nonlinearity='relu')
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def set_precise(self, precise):
        self.engine.precise = bool(precise)

    def forward(self, x, want_global=True):
        """x [B,3,H,W] -> [tuple(features [B,C,h,w] ...), global_embedding [B,512]]   (maskclip_vit.py:591-596)."""
        names = tuple(n for n, _ in self.named_parameters())
        params = tuple(p for _, p in self.named_parameters())
        self._last_want_global = want_global
        outs = _VitFunction.apply(self, x, want_global, torch.is_grad_enabled(), names, *params)
        nfe = len(outs) - (1 if want_global else 0)
        feats = tuple(o.permute(0, 3, 1, 2) for o in outs[:nfe])       # NHWC storage, NCHW view (values identical to the reference)
        return [feats, outs[nfe] if want_global else None]
