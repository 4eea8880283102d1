# Frozen MaskCLIP image encoder that produces the guidance pseudo-labels (the reference's configs/_base_/models/mcvit16.py).
# It is the same CLIP ViT-B/16 tower as the trainable backbone; the only difference is `out_indices=None`: no intermediate taps,
# just the final dense CLIP embedding.  The shared tower settings live in one place so the two configs cannot drift apart.
img_size = 512

CLIP_VIT_B16 = dict(
    # geometry
    patch_size=16, in_channels=3, embed_dims=768, num_layers=12, num_heads=12, mlp_ratio=4, num_fcs=2,
    # CLIP specifics: no patch bias, pre-LN before the blocks, LN after them, exact GELU, biased QKV
    patch_bias=False, patch_norm=False, pre_norm=True, final_norm=True, qkv_bias=True,
    norm_cfg=dict(type='LN', eps=1e-6), act_cfg=dict(type='GELU'),
    # tokens
    with_cls_token=True, output_cls_token=False, interpolate_mode='bicubic',
    # regularisation is off everywhere in SemiVL
    drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, norm_eval=False,
    # MaskCLIP outputs
    return_clip_embed=True, return_qkv=True,
)

backbone = dict(type='MaskClipVisionTransformer', pretrained='pretrained/clip2mmseg_ViT16_clip_backbone.pth',
                img_size=(img_size, img_size), out_indices=None, **CLIP_VIT_B16)
