"""TEST INFRASTRUCTURE ONLY (checker, never the product path).

CPU fp32 restatement of the SemiVL training hot path, written from the
reference's behaviour (file:line citations into /root/reference) as plain
functional PyTorch over a state-dict that uses the REFERENCE'S parameter names
(SURVEY.md Appendix C).  It contains no mmcv/mmseg dependency, so it travels to
the GPU box, where `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline / `--impl reference` legs use it as the checker / CPU baseline.

PINNING: the reference ships no tests or golden vectors, so
this oracle is pinned against outputs of the unmodified reference itself, run in
the authoring container through oracle/ref_shim.py; the vectors are committed
under tests/golden/ by oracle/make_golden.py (tests/test_oracle_golden.py checks
them on every run; `python -m oracle.make_golden` regenerates them from
/root/reference in the authoring container).

Nothing under semivl_b200/ may import this module.
"""
import math

import torch
import torch.nn.functional as F

VIT_LAYERS = 12


# ----------------------------------------------------------------------------- configuration
class ModelCfg:
    """Static hyper-parameters of the sk04 model (configs/_base_/models/vlm-vlg-aspp-s2p4-sk04-ftap-mcvitb.py:16-69)."""

    def __init__(self, img_size=512, num_classes=21, patch=16, embed=768, heads=12, layers=12,
                 out_indices=(0, 4, 12), ln_eps=1e-6, channels=128, text_channels=128,
                 up_channels=(64, 32), skip_channels=(32, 16), head_layers=2, head_heads=4,
                 pool=4, conv1_ksize=7, align_corners=False):
        self.img_size, self.num_classes, self.patch, self.embed = img_size, num_classes, patch, embed
        self.heads, self.layers, self.out_indices, self.ln_eps = heads, layers, tuple(out_indices), ln_eps
        self.channels, self.text_channels = channels, text_channels
        self.up_channels, self.skip_channels = tuple(up_channels), tuple(skip_channels)
        self.head_layers, self.head_heads, self.pool = head_layers, head_heads, pool
        self.conv1_ksize, self.align_corners = conv1_ksize, align_corners


# ----------------------------------------------------------------------------- ViT (maskclip_vit.py)
def _mha(x, p, pre, heads):
    """nn.MultiheadAttention self-attention as wrapped by mmcv (maskclip_vit.py:77-84,141): returns attn output (no residual)."""
    B, L, E = x.shape
    qkv = F.linear(x, p[pre + "attn.attn.in_proj_weight"], p[pre + "attn.attn.in_proj_bias"])
    q, k, v = qkv.view(B, L, 3, heads, E // heads).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(E // heads)
    a = s.softmax(-1) @ v
    a = a.transpose(1, 2).reshape(B, L, E)
    return F.linear(a, p[pre + "attn.attn.out_proj.weight"], p[pre + "attn.attn.out_proj.bias"])


def _ffn(x, p, pre):
    """mmcv FFN without the identity (maskclip_vit.py:94-100): Linear -> exact GELU -> Linear."""
    h = F.gelu(F.linear(x, p[pre + "ffn.layers.0.0.weight"], p[pre + "ffn.layers.0.0.bias"]))
    return F.linear(h, p[pre + "ffn.layers.1.weight"], p[pre + "ffn.layers.1.bias"])


def encoder_layer(x, p, pre, heads, eps, want_v):
    """TransformerEncoderLayer.forward / forward_qkv (maskclip_vit.py:110-144).

    Only the v branch of forward_qkv is materialised: q and k are returned by the
    reference but never consumed on the hot path (maskclip_vit.py:576-585 keeps o[3]).
    """
    E = x.shape[-1]
    v = None
    y = F.layer_norm(x, (E,), p[pre + "ln1.weight"], p[pre + "ln1.bias"], eps)
    if want_v:
        wv = p[pre + "attn.attn.in_proj_weight"][2 * E:]
        bv = p[pre + "attn.attn.in_proj_bias"][2 * E:]
        v = F.linear(F.linear(y, wv, bv), p[pre + "attn.attn.out_proj.weight"], p[pre + "attn.attn.out_proj.bias"]) + x
        v = v + _ffn(F.layer_norm(v, (E,), p[pre + "ln2.weight"], p[pre + "ln2.bias"], eps), p, pre)
    x = x + _mha(y, p, pre, heads)
    x = x + _ffn(F.layer_norm(x, (E,), p[pre + "ln2.weight"], p[pre + "ln2.bias"], eps), p, pre)
    return x, v


def resize_pos_embed(pos, hw, pos_hw):
    """maskclip_vit.py:462-490 -- bicubic, align_corners=False, cls row kept."""
    C = pos.shape[2]
    grid = pos[:, -pos_hw[0] * pos_hw[1]:].reshape(1, pos_hw[0], pos_hw[1], C).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=hw, mode="bicubic", align_corners=False)
    return torch.cat((pos[:, :1], grid.flatten(2).transpose(1, 2)), dim=1)


def vit_forward(img, p, cfg, pre="backbone.", out_indices=None, pos_img_size=None):
    """MaskClipVisionTransformer.forward (maskclip_vit.py:492-596) for the in-scope config
    (pre_norm, final_norm, return_clip_embed, return_qkv, with_cls_token).

    Returns ([features..., clip_embedding] as NCHW fp32, global_embedding[B,512]).
    `pos_img_size`: the img_size the module was BUILT with (pos_embed rows); differs from
    the input size for the frozen clip_encoder (builder.py:141-145).
    """
    out_indices = cfg.out_indices if out_indices is None else out_indices
    B, _, H, W = img.shape
    ps, E = cfg.patch, cfg.embed
    ph, pw = (-H) % ps, (-W) % ps          # mmseg PatchEmbed padding='corner' (bottom/right zero pad)
    x = F.pad(img, [0, pw, 0, ph]) if (ph or pw) else img
    x = F.conv2d(x, p[pre + "patch_embed.projection.weight"], p.get(pre + "patch_embed.projection.bias"), stride=ps)
    h, w = x.shape[-2:]
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat((p[pre + "cls_token"].expand(B, -1, -1), x), dim=1)
    pos = p[pre + "pos_embed"]
    if pos.shape[1] != x.shape[1]:          # maskclip_vit.py:448-459
        s = pos_img_size if pos_img_size is not None else cfg.img_size
        pos = resize_pos_embed(pos, (h, w), (s // ps, s // ps))
    x = x + pos
    x = F.layer_norm(x, (E,), p[pre + "ln0.weight"], p[pre + "ln0.bias"], cfg.ln_eps)

    feats = []
    v_last = None
    for i in range(cfg.layers):
        want_v = (i in out_indices) or i == cfg.layers - 1
        x, v = encoder_layer(x, p, f"{pre}layers.{i}.", cfg.heads, cfg.ln_eps, want_v)
        if i == cfg.layers - 1:
            x = F.layer_norm(x, (E,), p[pre + "ln1.weight"], p[pre + "ln1.bias"], cfg.ln_eps)
            v = F.layer_norm(v, (E,), p[pre + "ln1.weight"], p[pre + "ln1.bias"], cfg.ln_eps)
            v_last = v
        if i in out_indices:
            feats.append(v[:, 1:].reshape(B, h, w, E).permute(0, 3, 1, 2).contiguous())
    emb = v_last[:, 1:].reshape(B, h, w, E).permute(0, 3, 1, 2).contiguous()
    emb = F.conv2d(emb, p[pre + "proj.weight"])
    emb = emb / emb.norm(dim=1, keepdim=True)
    if cfg.layers in out_indices:
        feats.append(emb)
    glob = F.conv2d(x[:, 0][:, :, None, None], p[pre + "proj.weight"])[:, :, 0, 0]
    glob = glob / glob.norm(dim=1, keepdim=True)
    return feats, glob


# ----------------------------------------------------------------------------- VLG head (vlg_head.py)
def _gn_relu(x, p, name, groups):
    return F.relu(F.group_norm(x, groups, p[name + ".weight"], p[name + ".bias"], 1e-5))


def aspp_forward(x, p, pre):
    """ASPPModule.forward (vlg_head.py:84-113), 128 ch, GroupNorm(8)."""
    C = x.shape[1]
    g = C // 16
    outs = [_gn_relu(F.conv2d(x, p[pre + "aspp_convs.0.0.weight"]), p, pre + "aspp_convs.0.1", g)]
    for j, d in ((1, 6), (2, 12), (3, 18)):
        outs.append(_gn_relu(F.conv2d(x, p[pre + f"aspp_convs.{j}.0.weight"], padding=d, dilation=d),
                             p, pre + f"aspp_convs.{j}.1", g))
    pool = x.mean(dim=(2, 3), keepdim=True)                        # ASPPPooling (vlg_head.py:70-81)
    pool = _gn_relu(F.conv2d(pool, p[pre + "aspp_convs.4.gap.1.weight"]), p, pre + "aspp_convs.4.gap.2", g)
    outs.append(pool.expand(-1, -1, x.shape[2], x.shape[3]))        # bilinear(align_corners=True) of a 1x1 map == broadcast
    y = _gn_relu(F.conv2d(torch.cat(outs, 1), p[pre + "project.0.weight"]), p, pre + "project.1", g)
    return x + y


def semantic_transformer(x, text, p, pre, heads, pool):
    """SemanticTransformer.forward (vlg_head.py:39-67). x: [B,C,N,H,W]; text: [B,N,Ct]."""
    B, C, N, H, W = x.shape
    xp = F.avg_pool2d(x.permute(0, 2, 1, 3, 4).reshape(B * N, C, H, W), pool)
    Hp, Wp = xp.shape[-2:]
    xp = xp.reshape(B, N, C, Hp, Wp).permute(0, 3, 4, 1, 2).reshape(B * Hp * Wp, N, C)
    t = text[:, None, None].expand(B, Hp, Wp, N, text.shape[-1]).reshape(B * Hp * Wp, N, -1)
    tok = torch.cat([xp, t], dim=-1)
    tok, _ = encoder_layer(tok, p, pre + "transformer.", heads, 1e-5, False)
    tok = tok[..., :C]
    tok = tok.reshape(B, Hp, Wp, N, C).permute(0, 3, 4, 1, 2).reshape(B * N, C, Hp, Wp)
    tok = F.interpolate(tok, size=(H, W), mode="bilinear", align_corners=True)
    return x + tok.reshape(B, N, C, H, W).permute(0, 2, 1, 3, 4)


def up_forward(x, skip, p, pre, n_per_img):
    """Up.forward (vlg_head.py:116-137)."""
    x = F.conv_transpose2d(x, p[pre + "up.weight"], p[pre + "up.bias"], stride=2)
    skip = F.interpolate(skip, size=x.shape[-2:], mode="bilinear", align_corners=True)
    skip = skip.repeat_interleave(n_per_img, dim=0)                 # einops repeat 'b c h w -> (b n) c h w'
    x = torch.cat([x, skip], dim=1)
    g = p[pre + "conv.0.weight"].shape[0] // 16
    x = _gn_relu(F.conv2d(x, p[pre + "conv.0.weight"], padding=1), p, pre + "conv.1", g)
    x = _gn_relu(F.conv2d(x, p[pre + "conv.3.weight"], padding=1), p, pre + "conv.4", g)
    return x


def vlg_head_forward(feats, text, p, cfg, pre="decode_head.", class_to_concept=None, conv_feats=None):
    """VLGHead.forward (vlg_head.py:192-251) up to the 4x-resolution logits [B,N,4h,4w]; `conv_feats` = the conv encoder's feature
    list of the `skip_from_conv_feat` branch (vlg_head.py:196-205)."""
    img = feats[-1]
    skips = list(feats[:-1])[::-1] + list(conv_feats or [])[::-1]
    B, C, H, W = img.shape
    text = text.to(img.dtype)[None].expand(B, -1, -1)
    N = text.shape[1]
    x = torch.einsum("bchw,bnc->bnhw", F.normalize(img, dim=1), F.normalize(text, dim=-1))
    x = x.reshape(B * N, 1, H, W)
    k = cfg.conv1_ksize
    x = F.conv2d(x, p[pre + "conv1.weight"], p[pre + "conv1.bias"], padding=(k - 1) // 2)
    x = aspp_forward(x, p, pre + "aspp.")
    x = x.reshape(B, N, -1, H, W).permute(0, 2, 1, 3, 4)
    t = F.relu(F.linear(F.normalize(text, dim=-1), p[pre + "text_proj.0.weight"], p[pre + "text_proj.0.bias"]))
    for l in range(cfg.head_layers):
        x = semantic_transformer(x, t, p, f"{pre}layers.{l}.", cfg.head_heads, cfg.pool)
    sk = [F.relu(F.conv2d(f, p[f"{pre}skip_proj.{j}.0.weight"], p[f"{pre}skip_proj.{j}.0.bias"], padding=1))
          for j, f in enumerate(skips)]
    x = x.permute(0, 2, 1, 3, 4).reshape(B * N, -1, H, W)
    x = up_forward(x, sk[0], p, pre + "up1.", N)
    x = up_forward(x, sk[1], p, pre + "up2.", N)
    x = F.conv2d(x, p[pre + "head.weight"], p[pre + "head.bias"], padding=1)
    x = x.reshape(B, N, x.shape[-2], x.shape[-1])
    if class_to_concept is not None:                                 # text_embeddings.py:188-193
        x = aggregate_concepts(x, class_to_concept)
    return x


def aggregate_concepts(pred, class_to_concept):
    return torch.stack([pred[:, idx].max(dim=1).values for idx in class_to_concept], dim=1)


# ----------------------------------------------------------------------------- segmentor surface (builder.py, vlm.py)
def model_forward(img, p, text, cfg, need_fp=False, drop_masks=None, class_to_concept=None, return_lowres=False):
    """forward_wrapper (builder.py:56-102) + VLM.extract_feat (vlm.py:112-123).

    need_fp: every tap is concatenated with a channel-dropout copy (builder.py:78-85).
    `drop_masks` are the injected {0,1} keep masks [B,C,1,1] per tap (scale 1/(1-p) applied here),
    because Philox streams cannot be matched (SURVEY.md §8c caveat ii).
    """
    feats, _ = vit_forward(img, p, cfg)
    if need_fp:
        feats = [torch.cat((f, f * m * 2.0)) for f, m in zip(feats, drop_masks)]   # fp_rate 0.5 -> scale 2
    low = vlg_head_forward(feats, text, p, cfg, class_to_concept=class_to_concept)
    out = F.interpolate(low, size=(cfg.img_size, cfg.img_size), mode="bilinear", align_corners=cfg.align_corners)
    out = F.interpolate(out, size=img.shape[2:], mode="bilinear", align_corners=cfg.align_corners)  # builder.py:93-97
    if return_lowres:
        return out, low
    return out.chunk(2) if need_fp else out


def forward_maskclip(img, p, mcc_text, cfg, conf_thresh, pos_img_size=512, class_to_concept=None):
    """VLM.forward_maskclip (vlm.py:90-110): frozen clip_encoder -> labels int64, 255 = low confidence."""
    with torch.no_grad():
        feats, _ = vit_forward(img, p, cfg, pre="clip_encoder.", out_indices=(cfg.layers,), pos_img_size=pos_img_size)
        dense = F.conv2d(feats[-1], mcc_text.to(feats[-1].dtype)[:, :, None, None])
        if class_to_concept is not None:
            dense = aggregate_concepts(dense, class_to_concept)
        dense = F.interpolate(dense, size=img.shape[-2:], mode="bilinear", align_corners=cfg.align_corners)
        conf, lab = (100.0 * dense).softmax(dim=1).max(dim=1)
        lab = lab.clone()
        lab[conf < conf_thresh] = 255
    return lab


# ----------------------------------------------------------------------------- losses (semivl.py, utils/train_utils.py)
def cutmix_img(img, img_mix, box):
    """cutmix_img_ (train_utils.py:19-21), out of place."""
    m = (box.unsqueeze(1) == 1).expand(img.shape)
    return torch.where(m, img_mix, img)


def cutmix_mask(mask, mask_mix, box):
    """cutmix_mask (train_utils.py:24-27)."""
    return torch.where(box == 1, mask_mix, mask)


def confidence_weighted_loss(loss, conf, ignore, conf_thresh=0.95, conf_mode="pixelwise"):
    """train_utils.py:30-49."""
    valid = ignore != 255
    if conf_mode == "pixelwise":
        return (loss * ((conf >= conf_thresh) & valid)).sum() / valid.sum()
    if conf_mode == "pixelratio":
        ratio = ((conf >= conf_thresh) & valid).sum(dim=(1, 2), keepdim=True) / valid.sum(dim=(1, 2), keepdim=True)
        return (loss * ratio).sum() / valid.sum()
    if conf_mode == "pixelavg":
        avg = (conf * valid).sum(dim=(1, 2), keepdim=True) / valid.sum(dim=(1, 2), keepdim=True)
        return (loss.sum() * avg).sum() / valid.sum()
    raise ValueError(conf_mode)


def mc_loss(pred, lab, ign, reduce="mean_all"):
    """compute_mc_loss (semivl.py:52-58)."""
    if reduce == "mean":
        return F.cross_entropy(pred, lab, ignore_index=255)
    l = F.cross_entropy(pred, lab, ignore_index=255, reduction="none").sum()
    return l / ((ign != 255).sum() if reduce == "mean_valid" else ign.numel())


def supervised_step_loss(img, mask, p, text, cfg):
    """Shape of BASELINE config 2 (third_party/unimatch/supervised.py:273-289): model(img) -> CE(ignore 255)."""
    return F.cross_entropy(model_forward(img, p, text, cfg), mask, ignore_index=255)


def semivl_step_losses(batch, p, text, mcc_text, cfg, hp, drop_masks, mcc_class_to_concept=None, clip_pos_img_size=512):
    """One SemiVL iteration, semivl.py:224-323, as a function.  `batch` keys follow semivl.py:203-221.

    hp: dict(conf_thresh, conf_mode, mcc_conf_thresh, mcc_loss_reduce, mcc_lambda (already scheduled scalar)).
    Returns (loss, dict of the individual terms).
    """
    b = batch
    img_s1 = cutmix_img(b["img_s1"], b["img_s1_other"], b["mix1"])
    img_s2 = cutmix_img(b["img_s2"], b["img_s2_other"], b["mix2"])
    with torch.no_grad():
        pred_w_other = model_forward(b["img_w_other"], p, text, cfg)
        conf_w_other, mask_w_other = pred_w_other.softmax(dim=1).max(dim=1)
        lam = hp["mcc_lambda"]
        if lam != 0:
            mclip = forward_maskclip(torch.cat((b["img_w"], b["img_w_other"])), p, mcc_text, cfg,
                                     hp["mcc_conf_thresh"], clip_pos_img_size, mcc_class_to_concept)
            mclip, mclip_other = mclip.split([b["img_w"].shape[0], b["img_w_other"].shape[0]])
            mclip = torch.where(b["ignore_mask"] == 255, torch.full_like(mclip, 255), mclip)
            mclip_other = torch.where(b["ignore_mask_other"] == 255, torch.full_like(mclip_other, 255), mclip_other)
    preds, preds_fp = model_forward(torch.cat((b["img_x"], b["img_w"])), p, text, cfg, need_fp=True, drop_masks=drop_masks)
    pred_x, pred_w = preds.chunk(2)
    _, pred_w_fp = preds_fp.chunk(2)
    pred_s1, pred_s2 = model_forward(torch.cat((img_s1, img_s2)), p, text, cfg).chunk(2)
    conf_w, mask_w = pred_w.detach().softmax(dim=1).max(dim=1)

    mix1, mix2 = b["mix1"], b["mix2"]
    mask_m1, mask_m2 = cutmix_mask(mask_w, mask_w_other, mix1), cutmix_mask(mask_w, mask_w_other, mix2)
    conf_m1, conf_m2 = cutmix_mask(conf_w, conf_w_other, mix1), cutmix_mask(conf_w, conf_w_other, mix2)
    ign_m1 = cutmix_mask(b["ignore_mask"], b["ignore_mask_other"], mix1)
    ign_m2 = cutmix_mask(b["ignore_mask"], b["ignore_mask_other"], mix2)

    ce = lambda a, t: F.cross_entropy(a, t, reduction="none")
    cw = lambda l, c, i: confidence_weighted_loss(l, c, i, hp["conf_thresh"], hp["conf_mode"])
    terms = dict(loss_x=F.cross_entropy(pred_x, b["mask_x"], ignore_index=255),
                 loss_s1=cw(ce(pred_s1, mask_m1), conf_m1, ign_m1),
                 loss_s2=cw(ce(pred_s2, mask_m2), conf_m2, ign_m2),
                 loss_fp=cw(ce(pred_w_fp, mask_w), conf_w, b["ignore_mask"]))
    loss = (terms["loss_x"] + 0.25 * terms["loss_s1"] + 0.25 * terms["loss_s2"] + 0.5 * terms["loss_fp"]) / 2.0
    if lam != 0:
        mc1, mc2 = cutmix_mask(mclip, mclip_other, mix1), cutmix_mask(mclip, mclip_other, mix2)
        terms["loss_mc_s1"] = mc_loss(pred_s1, mc1, ign_m1, hp["mcc_loss_reduce"])
        terms["loss_mc_s2"] = mc_loss(pred_s2, mc2, ign_m2, hp["mcc_loss_reduce"])
        terms["loss_mc_fp"] = mc_loss(pred_w_fp, mclip, b["ignore_mask"], hp["mcc_loss_reduce"])
        loss = loss + lam * (0.25 * terms["loss_mc_s1"] + 0.25 * terms["loss_mc_s2"] + 0.5 * terms["loss_mc_fp"])
    return loss, terms


# ----------------------------------------------------------------------------- optimizer (experiments.py:246-255, semivl.py:326-345)
def param_group_hparams(name, base_lr, base_wd, custom_keys):
    """mmcv DefaultOptimizerConstructor (1.4.4) rule restated: keys sorted alphabetically then by
    descending length; the FIRST key that is a substring of the parameter name sets lr_mult / decay_mult."""
    lr, wd = base_lr, base_wd
    for key in sorted(sorted(custom_keys.keys()), key=len, reverse=True):
        if key in name:
            lr = base_lr * custom_keys[key].get("lr_mult", 1.0)
            wd = base_wd * custom_keys[key].get("decay_mult", 1.0)
            break
    return lr, wd


def adamw_step(p, g, m, v, step, lr, wd, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.AdamW single-tensor update (decoupled weight decay), in place on p, m, v."""
    p.mul_(1 - lr * wd)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


def poly_lr(initial_lr, iters, max_iters, power=0.9):
    """semivl.py:343-344 (applied AFTER the step of iteration `iters`)."""
    return initial_lr * (1 - iters / max_iters) ** power


# ----------------------------------------------------------------------------- deterministic fixture weights
def param_shapes(cfg, with_clip_encoder=True, clip_pos_img_size=512):
    """Name -> shape of every tensor of the sk04 VLM."""
    E, ps = cfg.embed, cfg.patch
    def vit(pre, img_size):
        s = {pre + "cls_token": (1, 1, E), pre + "pos_embed": (1, (img_size // ps) ** 2 + 1, E),
             pre + "patch_embed.projection.weight": (E, 3, ps, ps),
             pre + "ln0.weight": (E,), pre + "ln0.bias": (E,), pre + "ln1.weight": (E,), pre + "ln1.bias": (E,),
             pre + "proj.weight": (512, E, 1, 1)}
        for i in range(cfg.layers):
            s.update(_layer_shapes(f"{pre}layers.{i}.", E, 4 * E))
        return s
    shapes = vit("backbone.", cfg.img_size)
    C, Ct, h = cfg.channels, cfg.text_channels, "decode_head."
    k = cfg.conv1_ksize
    shapes.update({h + "conv1.weight": (C, 1, k, k), h + "conv1.bias": (C,)})
    for j in range(4):
        kk = 1 if j == 0 else 3
        shapes.update({h + f"aspp.aspp_convs.{j}.0.weight": (C, C, kk, kk),
                       h + f"aspp.aspp_convs.{j}.1.weight": (C,), h + f"aspp.aspp_convs.{j}.1.bias": (C,)})
    shapes.update({h + "aspp.aspp_convs.4.gap.1.weight": (C, C, 1, 1), h + "aspp.aspp_convs.4.gap.2.weight": (C,),
                   h + "aspp.aspp_convs.4.gap.2.bias": (C,), h + "aspp.project.0.weight": (C, 5 * C, 1, 1),
                   h + "aspp.project.1.weight": (C,), h + "aspp.project.1.bias": (C,)})
    for l in range(cfg.head_layers):
        shapes.update(_layer_shapes(f"{h}layers.{l}.transformer.", C + Ct, 4 * C))
    shapes.update({h + "text_proj.0.weight": (Ct, 512), h + "text_proj.0.bias": (Ct,)})
    for j, sc in enumerate(cfg.skip_channels):
        shapes.update({h + f"skip_proj.{j}.0.weight": (sc, E, 3, 3), h + f"skip_proj.{j}.0.bias": (sc,)})
    cin = C
    for name, cout, sc in (("up1.", cfg.up_channels[0], cfg.skip_channels[0]), ("up2.", cfg.up_channels[1], cfg.skip_channels[1])):
        shapes.update({h + name + "up.weight": (cin, cin - sc, 2, 2), h + name + "up.bias": (cin - sc,),
                       h + name + "conv.0.weight": (cout, cin, 3, 3), h + name + "conv.1.weight": (cout,), h + name + "conv.1.bias": (cout,),
                       h + name + "conv.3.weight": (cout, cout, 3, 3), h + name + "conv.4.weight": (cout,), h + name + "conv.4.bias": (cout,)})
        cin = cout
    shapes.update({h + "head.weight": (1, cin, 3, 3), h + "head.bias": (1,)})
    if with_clip_encoder:
        shapes.update(vit("clip_encoder.", clip_pos_img_size))
    return shapes


def _layer_shapes(pre, E, ff):
    return {pre + "ln1.weight": (E,), pre + "ln1.bias": (E,), pre + "ln2.weight": (E,), pre + "ln2.bias": (E,),
            pre + "attn.attn.in_proj_weight": (3 * E, E), pre + "attn.attn.in_proj_bias": (3 * E,),
            pre + "attn.attn.out_proj.weight": (E, E), pre + "attn.attn.out_proj.bias": (E,),
            pre + "ffn.layers.0.0.weight": (ff, E), pre + "ffn.layers.0.0.bias": (ff,),
            pre + "ffn.layers.1.weight": (E, ff), pre + "ffn.layers.1.bias": (E,)}


def fixture_state_dict(shapes, seed=0, gain=1.0):
    """Deterministic 'alive' weights: one generator per tensor, seeded from (seed, name), so the
    same numbers can be loaded into the reference modules, this oracle and the CUDA path regardless
    of construction order.  Matrices ~ N(0, gain/sqrt(fan_in)); norm scales 1+0.1 N; biases 0.02 N.
    (trunc-normal 0.02 init, maskclip_vit.py:416-429, gives class-degenerate logits -- SURVEY.md §0 fact 6.)
    """
    import zlib
    sd = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31))
        r = torch.randn(shp, generator=g)
        leaf = name.split(".")[-1]
        is_norm = any(t in name for t in (".ln0.", ".ln1.", ".ln2.", "aspp_convs.0.1.", "aspp_convs.1.1.", "aspp_convs.2.1.",
                                            "aspp_convs.3.1.", "gap.2.", "project.1.", "conv.1.", "conv.4."))
        if "pos_embed" in name or "cls_token" in name:
            t = 0.02 * r
        elif is_norm and leaf == "weight":
            t = 1.0 + 0.1 * r
        elif leaf == "bias" or leaf.endswith("_bias"):
            t = 0.02 * r
        elif name.endswith("up.weight"):
            t = r * (gain / math.sqrt(shp[0]))                     # ConvTranspose: (Cin, Cout, 2, 2), one tap per output
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            t = r * (gain / math.sqrt(fan_in))
        sd[name] = t.contiguous()
    return sd


# ----------------------------------------------------------------------------- evaluation (third_party/unimatch/supervised.py:40-164)
def predict_reference(model, img, mask, mode, cfg):
    """CPU restatement of `predict` for any callable `model(img) -> logits [b, nclass, h, w]`; returns (labels, stitched scores)."""
    n = cfg["nclass"]
    b, _, h, w = img.shape
    if mode == "padded_sliding_window":                                      # supervised.py:41-65
        grid, stride = cfg["crop_size"], cfg["stride"]
        stride = int(grid * stride) if stride < 1 else stride
        final = torch.zeros(b, n, h, w)
        for row in range(0, h, stride):
            for col in range(0, w, stride):
                y2, x2 = min(h, row + grid), min(w, col + grid)
                crop = torch.zeros(b, 3, grid, grid)
                crop[:, :, :y2 - row, :x2 - col] = img[:, :, row:y2, col:x2]
                final[:, :, row:y2, col:x2] += model(crop).softmax(dim=1)[:, :, :y2 - row, :x2 - col]
    elif mode == "zegclip_sliding_window":                                   # supervised.py:67-103
        s, c = cfg["stride"], cfg["crop_size"]
        hg, wg = max(h - c + s - 1, 0) // s + 1, max(w - c + s - 1, 0) // s + 1
        final, count = torch.zeros(b, n, h, w), torch.zeros(b, 1, h, w)
        for hi in range(hg):
            for wi in range(wg):
                y2, x2 = min(hi * s + c, h), min(wi * s + c, w)
                y1, x1 = max(y2 - c, 0), max(x2 - c, 0)
                final += F.pad(model(img[:, :, y1:y2, x1:x2]), (x1, w - x2, y1, h - y2))
                count[:, :, y1:y2, x1:x2] += 1
        assert (count == 0).sum() == 0
        final = F.interpolate(final / count, size=mask.shape[-2:], mode="bilinear", align_corners=True)
    elif mode == "sliding_window":                                           # supervised.py:105-117
        grid = cfg["crop_size"]
        step = int(grid * 2 / 3)
        final = torch.zeros(b, n, h, w)
        for row in range(0, h, step):
            for col in range(0, w, step):
                y2, x2 = min(h, row + grid), min(w, col + grid)
                final[:, :, row:y2, col:x2] += model(img[:, :, row:y2, col:x2]).softmax(dim=1)
    else:
        if mode == "center_crop":                                            # supervised.py:120-124
            c = cfg["crop_size"]
            sh, sw = (h - c) // 2, (w - c) // 2
            img = img[:, :, sh:sh + c, sw:sw + c]
        final = model(img)
    return final.argmax(dim=1), final


def intersection_and_union(output, target, K, ignore_index=255):
    """third_party/unimatch/util/utils.py:91-103 (numpy): (area_intersection, area_union, area_target), each int64 [K]."""
    import numpy as np
    output = np.asarray(output).reshape(-1).copy()
    target = np.asarray(target).reshape(-1)
    output[target == ignore_index] = ignore_index
    inter = output[output == target]
    bins = np.arange(K + 1)
    ai = np.histogram(inter, bins=bins)[0]
    ao = np.histogram(output, bins=bins)[0]
    at = np.histogram(target, bins=bins)[0]
    return ai, ao + at - ai, at


class StubSegModel(torch.nn.Module):
    """Deterministic stand-in for a segmentor in the evaluation fixtures: logits = 3x3 conv of the image + a class-dependent
    horizontal ramp (so that windows cut at different offsets disagree near their borders, like a real model does)."""

    def __init__(self, nclass, weight=None, seed=0):
        super().__init__()
        if weight is None:
            weight = torch.randn(nclass, 3, 3, 3, generator=torch.Generator().manual_seed(seed))
        self.w = torch.nn.Parameter(torch.as_tensor(weight).float(), requires_grad=False)

    def forward(self, x):
        y = F.conv2d(x, self.w.to(x.device), padding=1)
        ramp = torch.linspace(0, 1, x.shape[-1], device=x.device)[None, None, None, :] * \
            torch.arange(y.shape[1], device=x.device, dtype=y.dtype)[None, :, None, None]
        return y + 0.1 * ramp


# ----------------------------------------------------------------------------- input stage (third_party/unimatch/dataset/transform.py)
def crop_flip_normalize_reference(img_u8, mask_u8, size, x0, y0, flip, pad_value=255):
    """numpy/torch restatement of transform.crop (9-24) + hflip (27-31) + normalize (34-41) for GIVEN draws (x0, y0, flip), and the
    ignore mask of semi.py:99-103.  img_u8 [h,w,3] uint8, mask_u8 [h,w] uint8 -> (f32 [3,size,size], int64 labels, int64 ignore mask)."""
    import numpy as np
    h, w = mask_u8.shape
    ph, pw = max(h, size), max(w, size)
    pi = np.zeros((ph, pw, 3), np.uint8)
    pi[:h, :w] = img_u8
    pm = np.full((ph, pw), pad_value, np.uint8)
    pm[:h, :w] = mask_u8
    pi, pm = pi[y0:y0 + size, x0:x0 + size], pm[y0:y0 + size, x0:x0 + size]
    if flip:
        pi, pm = pi[:, ::-1], pm[:, ::-1]
    t = torch.from_numpy(np.ascontiguousarray(pi)).permute(2, 0, 1).float().div(255)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
    t = t.sub(mean).div(std)
    lab = torch.from_numpy(np.ascontiguousarray(pm)).long()
    ign = torch.zeros_like(lab)
    ign[lab == 254] = 255
    return t, lab, ign
