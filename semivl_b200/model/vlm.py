"""`VLM` segmentor (model/vlm.py:27-127): backbone + VLG head (+ frozen MaskCLIP `clip_encoder`), backbone freezing,
MaskCLIP pseudo-labels and feature extraction, with the reference's attribute surface."""
import os

import numpy as np
import torch
import torch.nn as nn

from .. import lib as L
from .. import ops
from ..registry import SEGMENTORS, build_backbone, build_head
from ..text_embeddings import concept_offsets

_CFG_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def resolve_path(path):
    """The reference uses paths relative to its repo root (builder.py:126-141); resolve them inside this package too."""
    if path is None or os.path.isabs(path) or os.path.exists(path):
        return path
    cand = os.path.join(_CFG_ROOT, path)
    return cand if os.path.exists(cand) else path


@SEGMENTORS.register_module()
class VLM(nn.Module):
    def __init__(self, backbone, decode_head, freeze_backbone=False, exclude_keys=None, load_text_embedding=None,
                 load_mcc_text_embedding=None, load_pl_text_embedding=None, clip_encoder=None, conv_encoder=None,
                 maskclip_class_filter=None, maskclip_trust_head=None, renorm_clip_img=False, neck=None, auxiliary_head=None,
                 train_cfg=None, test_cfg=None, pretrained=None, init_cfg=None):
        super().__init__()
        assert load_text_embedding == load_pl_text_embedding
        assert maskclip_class_filter is None and maskclip_trust_head is None and neck is None and auxiliary_head is None
        backbone = dict(backbone)
        if pretrained is not None:                        # mmseg EncoderDecoder pushes `pretrained` into the backbone cfg
            backbone["pretrained"] = pretrained
        self.backbone = build_backbone(backbone)
        self.decode_head = build_head(decode_head)
        self.align_corners = self.decode_head.align_corners
        self.num_classes = self.decode_head.num_classes
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.local_iter = 0
        self.clip_encoder = build_backbone(clip_encoder) if clip_encoder is not None else None
        self.conv_encoder = build_backbone(conv_encoder) if conv_encoder is not None else None          # model/vlm.py:50-52
        self.load_text_embedding = load_text_embedding
        self.decode_head.load_text_embedding = load_text_embedding
        self.load_mcc_text_embedding = load_mcc_text_embedding
        self.renorm_clip_img = renorm_clip_img
        if not self.load_mcc_text_embedding:
            raise NotImplementedError
        self.loaded_mcc_text_feat = torch.from_numpy(np.load(resolve_path(self.load_mcc_text_embedding))).float()
        self._text_cache = {}
        if freeze_backbone:
            self.freeze(self.backbone, exclude_keys=exclude_keys)

    # ------------------------------------------------------------------ reference surface
    def init_weights(self):
        for m in (self.backbone, self.decode_head, self.clip_encoder, self.conv_encoder):
            if m is not None:
                m.init_weights()

    def set_precise(self, precise):
        for m in (self.backbone, self.decode_head, self.clip_encoder, self.conv_encoder):
            if m is not None:
                m.set_precise(precise)

    def renormalize_img_for_clip(self, img):
        """model/vlm.py:69-78: undo the ImageNet normalisation, apply CLIP's.  The six constants live on the device (cached per device on
        first use, so that a captured CUDA graph of the step contains no host -> device copy)."""
        if not self.renorm_clip_img:
            return img
        key = ("renorm", str(img.device))
        if key not in self._text_cache:
            t = lambda v: torch.tensor(v, device=img.device).view(1, -1, 1, 1)
            self._text_cache[key] = tuple(t(v) for v in ([0.229, 0.224, 0.225], [0.485, 0.456, 0.406], [0.48145466, 0.4578275, 0.40821073],
                                                        [0.26862954, 0.26130258, 0.27577711]))
        s_in, m_in, m_clip, s_clip = self._text_cache[key]
        return (img * s_in + m_in - m_clip) / s_clip

    def freeze(self, model, exclude_keys=None):
        for n, m in model.named_parameters():
            m.requires_grad = False
            if exclude_keys is not None:
                assert isinstance(exclude_keys, list)
                if any(str(k) in n for k in exclude_keys):
                    m.requires_grad = True

    def _text(self, device):
        """The reference re-reads the .npy on every forward (vlm.py:116); the table is immutable, so it is cached per device."""
        key = (self.load_text_embedding, str(device))
        if key not in self._text_cache:
            self._text_cache[key] = torch.from_numpy(np.load(resolve_path(self.load_text_embedding))).to(device)
        return self._text_cache[key]

    def extract_feat(self, img):
        orig_img = img
        img = self.renormalize_img_for_clip(img)
        visual_feat = self.backbone(img)
        self.decode_head.load_text_embedding = self.load_text_embedding
        conv_feat = None
        if self.conv_encoder is not None:                       # the conv encoder sees the ImageNet-normalised image (model/vlm.py:113,120-121)
            conv_feat = list(self.conv_encoder(orig_img))
        return [visual_feat, self._text(img.device), conv_feat]

    def _decode_head_forward_test(self, x, img_metas):
        return self.decode_head.forward(x, force_output_pred_masks=True)["pred_masks"]

    def maskclip_lowres(self, img):
        """Class-major MaskCLIP scores [B, N, h, w] of the frozen clip_encoder (vlm.py:96-102), before upsampling."""
        img = self.renormalize_img_for_clip(img)
        enc = self.clip_encoder
        enc.eval()
        with torch.no_grad():
            dev = img.device
            p = {n: t.detach() for n, t in enc.named_parameters()}
            feats, _, _ = enc.engine.forward(img, p, need_grad=False, want_global=False)
            emb = feats[-1]                                           # [B,h,w,512] unit-norm
            B, h, w, D = emb.shape
            pr = enc.engine.precise
            text = self.loaded_mcc_text_feat.to(dev)
            K = text.shape[0]
            Kp = (K + 7) // 8 * 8
            scores = torch.empty(B * h * w, Kp, device=dev, dtype=torch.float32)
            ops.gemm(ops.to_act(emb.view(B * h * w, D), pr), ops.prep_weight(text, pr), scores, n=K, k=D, precise=pr)
            offs = concept_offsets(self.load_mcc_text_embedding, K, self.num_classes, dev)
            low = torch.empty(B, self.num_classes, h, w, device=dev, dtype=torch.float32)
            L.call("svl_group_max", scores, Kp, offs, low, B, self.num_classes, h * w)
        return low

    def forward_maskclip(self, img, conf_tresh):
        low = self.maskclip_lowres(img)
        B, N, h, w = low.shape
        H, W = img.shape[-2:]
        lab = torch.empty(B, H, W, device=img.device, dtype=torch.int64)
        L.call("svl_softmax_max", low, None, lab, B, N, h, w, H, W, 100.0, float(conf_tresh))
        return lab
