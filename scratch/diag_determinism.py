"""Run-to-run determinism of the bf16 head forward, stage by stage (walks the saved context of two identical forward calls)."""
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
from oracle import semivl_oracle as O
from semivl_b200.model import build_model
crop, b = 512, 2
cfg = dict(model='mmseg.vlm-vlg-aspp-s2p4-sk04-ftap-mcvitb', nclass=21, crop_size=crop, dataset='pascal', text_embedding_variant='single',
           mcc_text='single', pl_text='single', clip_encoder=None, disable_dropout=True, fp_rate=0.5, model_args=dict(pretrained=None), precise=False)
mc = O.ModelCfg(img_size=crop, num_classes=21)
sd = O.fixture_state_dict(O.param_shapes(mc, with_clip_encoder=False), seed=0)
m = build_model(cfg); m.load_state_dict(sd); m = m.cuda()
img = torch.randn(b, 3, crop, crop, generator=torch.Generator().manual_seed(17)).cuda()
eng = m.decode_head.engine
ph = {n: p.data for n, p in m.decode_head.named_parameters()}
text = m._text(img.device)
with torch.no_grad():
    x = m.extract_feat(img)
    feats = [f.permute(0, 2, 3, 1).contiguous() for f in x[0][0]]
def walk(a, b_, path, out):
    if isinstance(a, torch.Tensor):
        if a.dtype in (torch.float32, torch.bfloat16) and a.numel() and a.shape == b_.shape:
            af, bf = a.float(), b_.float()
            d = (af - bf).abs().max().item(); r = af.abs().max().item()
            out.append((path, tuple(a.shape), str(a.dtype)[6:], d, r))
    elif isinstance(a, dict):
        for k in a:
            walk(a[k], b_[k], f"{path}.{k}", out)
    elif isinstance(a, (list, tuple)):
        for i, (u, v) in enumerate(zip(a, b_)):
            walk(u, v, f"{path}[{i}]", out)
with torch.no_grad():
    low1, c1 = eng.forward(feats, text, ph, need_grad=True)
    low2, c2 = eng.forward(feats, text, ph, need_grad=True)
out = []
walk(c1, c2, "ctx", out)
for path, shp, dt, d, r in out:
    print(f"{path:60s} {str(shp):28s} {dt:9s} maxdiff {d:.3e} / range {r:.3e}  {'DIFF' if d > 0 else ''}")
print("low", (low1 - low2).abs().max().item(), low1.abs().max().item())
