"""OpenAI-CLIP ViT visual-tower state dict -> the key names `MaskClipVisionTransformer` loads.

Restates the ViT branch of third_party/maskclip/convert_clip_weights.py:27-64 of the reference (the tool that produces
`clip2mmseg_ViT16_clip_backbone.pth`, configs/_base_/models/*-mcvitb.py `pretrained=`): same renames, same transposed
projection, same optional 'backbone.' prefix.  It works on a plain dict of tensors, so it needs neither the `clip` package
nor mmcv.

  python -m semivl_b200.convert_clip_weights clip_visual.pt out.pth [--backbone]
"""
import re
import sys

import torch

_BLOCK = re.compile(r"^transformer\.resblocks\.(\d+)\.(.+)$")
_IN_BLOCK = (
    ("attn.", "attn.attn."),            # nn.MultiheadAttention sits inside mmcv's MultiheadAttention wrapper
    ("ln_1.", "ln1."), ("ln_2.", "ln2."),
    ("mlp.c_fc.", "ffn.layers.0.0."),   # FFN = Sequential(Sequential(Linear, GELU, Dropout), Linear, Dropout)
    ("mlp.c_proj.", "ffn.layers.1."),
)


def rename_visual_key(key):
    """'class_embedding' -> 'cls_token', 'transformer.resblocks.3.mlp.c_fc.weight' -> 'layers.3.ffn.layers.0.0.weight', ..."""
    fixed = {"class_embedding": "cls_token", "positional_embedding": "pos_embed", "conv1.weight": "patch_embed.projection.weight",
             "proj": "proj.weight"}
    if key in fixed:
        return fixed[key]
    for old, new in (("ln_pre.", "ln0."), ("ln_post.", "ln1.")):
        if key.startswith(old):
            return new + key[len(old):]
    m = _BLOCK.match(key)
    if m:
        rest = m.group(2)
        for old, new in _IN_BLOCK:
            if rest.startswith(old):
                rest = new + rest[len(old):]
                break
        return f"layers.{m.group(1)}.{rest}"
    return key


def clip_visual_to_mmseg(state_dict, backbone_prefix=False, visual_prefix="visual."):
    """state_dict: a CLIP model state dict (keys 'visual.*' are used, the rest ignored).  Returns {'meta': {}, 'state_dict': {...}}."""
    out = {}
    for key, val in state_dict.items():
        if not key.startswith(visual_prefix):
            continue
        k = key[len(visual_prefix):]
        v = val.float()
        if k == "proj":
            v = v.t().contiguous()                      # [width, embed] -> Linear-style [embed, width]
        elif k == "class_embedding":
            v = v[None, None, :]
        elif k == "positional_embedding":
            v = v[None, :, :]
        name = rename_visual_key(k)
        out[("backbone." + name) if backbone_prefix and k != "proj" else name] = v
    return {"meta": {}, "state_dict": out}


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    sd = torch.load(src, map_location="cpu")
    sd = sd.get("state_dict", sd) if isinstance(sd, dict) else sd.state_dict()
    torch.save(clip_visual_to_mmseg(sd, backbone_prefix="--backbone" in sys.argv), dst)
