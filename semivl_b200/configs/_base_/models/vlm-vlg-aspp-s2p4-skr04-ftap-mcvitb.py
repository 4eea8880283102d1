# SemiVL Cityscapes model: CLIP ViT-B/16 (MaskCLIP v-path tap at layer 4 and the final embedding) + ResNetV1c stem/layer1 conv encoder
# (SyncBN) supplying the stride-4 skip + VLG head.
# Same keys and values as the reference's configs/_base_/models/vlm-vlg-aspp-s2p4-skr04-ftap-mcvitb.py:16-81.
norm_cfg = dict(type='SyncBN', requires_grad=True)
img_size = 512
_vit = dict(patch_size=16, patch_bias=False, in_channels=3, embed_dims=768, num_layers=12, num_heads=12, mlp_ratio=4,
            qkv_bias=True, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, with_cls_token=True, output_cls_token=False,
            norm_cfg=dict(type='LN', eps=1e-6), act_cfg=dict(type='GELU'), patch_norm=False, pre_norm=True, final_norm=True,
            return_clip_embed=True, return_qkv=True, interpolate_mode='bicubic', num_fcs=2, norm_eval=False)
model = dict(
    type='VLM',
    pretrained='pretrained/clip2mmseg_ViT16_clip_backbone.pth',
    backbone=dict(type='MaskClipVisionTransformer', img_size=(img_size, img_size), out_indices=[4, 12], **_vit),
    conv_encoder=dict(type='ResNetV1c', pretrained='pretrained/resnet101_v1c-e67eebb6.pth', depth=101, num_stages=1, out_indices=[0],
                      dilations=[1], strides=[1], norm_cfg=norm_cfg, style='pytorch', contract_dilation=True),
    decode_head=dict(type='VLGHead', img_size=img_size, num_classes=19, text_in_channels=512, text_channels=128,
                     up_channels=(64, 32), skip_in_channels=(768, 256), skip_channels=(32, 32), skip_from_conv_feat=True,
                     num_layers=2, num_heads=4, channels=128, pool_size=(4, 4), conv1_ksize=7, align_corners=False,
                     loss_decode=None),
    freeze_backbone=True,
    exclude_keys=['attn', 'pos_embed'],
)
