"""`build_model(cfg)` and the patched `model(img, need_fp=, only_fp=)` forward of the reference (model/builder.py:56-159)."""
import os
import types

import torch
from torch.nn import functional as F

from ..registry import Config, build_segmentor
from . import maskclip_vit, resnet, vlg_head, vlm  # noqa: F401  (register the types)
from .vlg_head import upsample_bilinear
from .vlm import VLM

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def nested_set(dic, key, value):
    keys = key.split('.')
    for k in keys[:-1]:
        dic = dic.setdefault(k, {})
    dic[keys[-1]] = value


def is_vlm(obj):
    return isinstance(obj, VLM)


def forward_wrapper(self, img, gt=None, need_fp=False, only_fp=False, forward_mode='default', drop_masks=None):
    """builder.py:56-102.  `drop_masks` (optional, list of {0,1} keep masks [B,C,1,1] per feature) replaces F.dropout2d's
    Philox draw so that parity tests can share the perturbation with the oracle (SURVEY.md §8c caveat ii)."""
    if forward_mode != 'default':
        raise ValueError(forward_mode)

    def drop(f, i):
        if drop_masks is not None:
            return f * (drop_masks[i].to(f.device) / (1.0 - self.fp_rate))
        return F.dropout2d(f, self.fp_rate)

    x = self.extract_feat(img)
    nf = len(x[0][0])
    if only_fp:
        x[0][0] = [drop(f, i) for i, f in enumerate(x[0][0])]
        if x[2] is not None:                                          # conv-encoder features are perturbed too (builder.py:69-71)
            x[2] = [drop(f, nf + i) for i, f in enumerate(x[2])]
    elif need_fp:
        x[0][0] = [torch.cat((f, drop(f, i))) for i, f in enumerate(x[0][0])]
        if x[0][1] is not None:
            x[0][1] = torch.cat((x[0][1], x[0][1]))
        if x[2] is not None:                                          # builder.py:83-85
            x[2] = [torch.cat((f, drop(f, nf + i))) for i, f in enumerate(x[2])]
    out = self._decode_head_forward_test(x, img_metas=None)
    if tuple(out.shape[2:]) != tuple(img.shape[2:]):           # identity resize otherwise (builder.py:93-97)
        out = upsample_bilinear(out, img.shape[2:])
    if need_fp:
        out = out.chunk(2)
    return out


def forward_lowres(self, img, need_fp=False, drop_masks=None):
    """Same as forward_wrapper but stops at the head's 4x-resolution class maps (input of the fused loss kernels)."""
    x = self.extract_feat(img)
    if need_fp:
        def drop(f, i):
            if drop_masks is not None:
                return f * (drop_masks[i].to(f.device) / (1.0 - self.fp_rate))
            return F.dropout2d(f, self.fp_rate)
        nf = len(x[0][0])
        x[0][0] = [torch.cat((f, drop(f, i))) for i, f in enumerate(x[0][0])]
        if x[2] is not None:
            x[2] = [torch.cat((f, drop(f, nf + i))) for i, f in enumerate(x[2])]
    return self.decode_head.forward_lowres(x)


def build_model(cfg):
    model_type = cfg['model']
    if 'mmseg.' not in model_type:
        raise ValueError(f"semivl_b200 implements the 'mmseg.vlm-vlg-*' SemiVL models, got {model_type}")
    model_type = model_type.replace('mmseg.', '')
    model_cfg_file = f'configs/_base_/models/{model_type}.py'
    if not os.path.exists(model_cfg_file):
        model_cfg_file = os.path.join(_PKG, model_cfg_file)
    mmseg_cfg = Config.fromfile(model_cfg_file)
    mmseg_cfg['model']['decode_head']['num_classes'] = cfg['nclass']
    if mmseg_cfg['img_size'] != cfg['crop_size']:
        nested_set(mmseg_cfg, 'img_size', cfg['crop_size'])
        nested_set(mmseg_cfg, 'model.backbone.img_size', (cfg['crop_size'], cfg['crop_size']))
        nested_set(mmseg_cfg, 'model.decode_head.img_size', cfg['crop_size'])
    prefix = {'pascal': 'voc12_wbg', 'cityscapes': 'cityscapes', 'coco': 'coco', 'ade': 'ade'}[cfg['dataset']]
    emb = 'configs/_base_/datasets/text_embedding/{}_{}.npy'
    nested_set(mmseg_cfg, 'model.load_text_embedding', emb.format(prefix, cfg['text_embedding_variant']))
    nested_set(mmseg_cfg, 'model.load_mcc_text_embedding', emb.format(prefix, cfg['mcc_text']))
    nested_set(mmseg_cfg, 'model.load_pl_text_embedding', emb.format(prefix, cfg['pl_text']))
    if cfg.get('clip_encoder') is not None:
        ce_file = f'configs/_base_/models/{cfg["clip_encoder"]}.py'
        if not os.path.exists(ce_file):
            ce_file = os.path.join(_PKG, ce_file)
        clip_encoder_cfg = Config.fromfile(ce_file)
        clip_encoder_cfg['img_size'] = mmseg_cfg['img_size']
        if cfg.get('mcc_fix_resize_pos'):
            clip_encoder_cfg['backbone']['img_size'] = mmseg_cfg['img_size']
        if 'clip_encoder_args' in cfg:           # extension: e.g. dict(pretrained=None) for random-init benchmarking
            clip_encoder_cfg['backbone'].update(cfg['clip_encoder_args'])
        mmseg_cfg['model']['clip_encoder'] = clip_encoder_cfg['backbone']
    if 'model_args' in cfg:
        mmseg_cfg['model'].update(cfg['model_args'])
    # extension: arithmetic mode per engine.  False = bf16 operands everywhere (throughput), True = split-bf16 x3 everywhere (parity),
    # 'head' / 'encoder' = split operands in that engine only (the frozen clip_encoder follows the encoder)
    pmode = cfg.get('precise', False)
    if pmode not in (False, True, 'head', 'encoder'):
        raise ValueError(f"cfg['precise'] must be False, True, 'head' or 'encoder', got {pmode!r}")
    for part in ('backbone', 'decode_head', 'clip_encoder', 'conv_encoder'):
        if mmseg_cfg['model'].get(part) is not None:
            mmseg_cfg['model'][part]['precise'] = pmode is True or pmode == ('head' if part == 'decode_head' else 'encoder')
    if 'conv_encoder_args' in cfg and mmseg_cfg['model'].get('conv_encoder') is not None:      # extension: e.g. dict(pretrained=None)
        mmseg_cfg['model']['conv_encoder'].update(cfg['conv_encoder_args'])
    model = build_segmentor(mmseg_cfg.model, train_cfg=mmseg_cfg.get('train_cfg'), test_cfg=mmseg_cfg.get('test_cfg'))
    model.disable_dropout = cfg['disable_dropout']
    model.fp_rate = cfg['fp_rate']
    model.forward = types.MethodType(forward_wrapper, model)
    model.forward_lowres = types.MethodType(forward_lowres, model)
    model.init_weights()
    return model
