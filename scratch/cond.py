import sys, numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from oracle import semivl_oracle as O
torch.set_num_threads(8)
g = dict(np.load("tests/golden/fwd_c64_b2.npz", allow_pickle=False))
crop = 64
mc = O.ModelCfg(img_size=crop, num_classes=21)
sd = O.fixture_state_dict(O.param_shapes(mc, with_clip_encoder=False), seed=0)
img = torch.from_numpy(g["img"]); lab = torch.from_numpy(g["label"].astype(np.int64))
text = torch.from_numpy(np.load("semivl_b200/configs/_base_/datasets/text_embedding/voc12_wbg_single.npy"))
def run(dt, noise):
    gen = torch.Generator().manual_seed(5)
    p = {}
    for k, v in sd.items():
        v = v.clone().to(dt)
        if noise and v.dim() > 1:
            v = v * (1 + noise * torch.randn(v.shape, generator=gen).to(dt))
        p[k] = v.requires_grad_(True)
    y = O.model_forward(img.to(dt), p, text, mc)
    F.cross_entropy(y, lab, ignore_index=255).backward()
    return y.detach(), p
y64, p64 = run(torch.float64, 0)
for noise in (1e-6, 1e-5, 3e-5):
    yn, pn = run(torch.float64, noise)
    rel = lambda a, b: ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()
    ks = ["decode_head.up2.conv.4.bias", "decode_head.up2.conv.4.weight", "decode_head.up2.conv.3.weight", "decode_head.conv1.weight", "decode_head.aspp.project.0.weight", "backbone.layers.0.attn.attn.in_proj_weight", "backbone.pos_embed"]
    print(noise, "logits", f"{rel(yn, y64):.1e}", " ".join(f"{k.split('.',1)[1][:22]}={rel(pn[k].grad, p64[k].grad):.1e}" for k in ks))
