# Frozen MaskCLIP image encoder used for the guidance pseudo-labels (reference configs/_base_/models/mcvit16.py:16-45):
# out_indices=None -> only the final dense CLIP embedding.
img_size = 512
backbone = dict(type='MaskClipVisionTransformer', pretrained='pretrained/clip2mmseg_ViT16_clip_backbone.pth',
                img_size=(img_size, img_size), patch_size=16, patch_bias=False, in_channels=3, embed_dims=768, num_layers=12,
                num_heads=12, mlp_ratio=4, out_indices=None, qkv_bias=True, drop_rate=0.0, attn_drop_rate=0.0,
                drop_path_rate=0.0, with_cls_token=True, output_cls_token=False, norm_cfg=dict(type='LN', eps=1e-6),
                act_cfg=dict(type='GELU'), patch_norm=False, pre_norm=True, final_norm=True, return_clip_embed=True,
                return_qkv=True, interpolate_mode='bicubic', num_fcs=2, norm_eval=False)
