import sys, numpy as np, torch
sys.path.insert(0, ".")
from oracle import semivl_oracle as O
from semivl_b200 import lib
from semivl_b200.engine.head import HeadCfg, HeadEngine
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
def rel(a, b): return ((a.float().cpu() - b.float().cpu()).abs().max() / (b.float().abs().max() + 1e-12)).item()
hw, b, n, precise = int(sys.argv[1]), 2, 21, sys.argv[2] == "1"
mc = O.ModelCfg(img_size=hw * 16, num_classes=n)
sd = O.fixture_state_dict(O.param_shapes(mc, with_clip_encoder=False), seed=0)
pcpu = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("decode_head.")}
g = torch.Generator().manual_seed(hw)
feats = [torch.randn(b, 768, hw, hw, generator=g).requires_grad_(True), torch.randn(b, 768, hw, hw, generator=g).requires_grad_(True)]
emb = torch.randn(b, 512, hw, hw, generator=g); emb = (emb / emb.norm(dim=1, keepdim=True)).requires_grad_(True)
text = torch.from_numpy(np.load("semivl_b200/configs/_base_/datasets/text_embedding/voc12_wbg_single.npy"))
low_ref = O.vlg_head_forward(feats + [emb], text, pcpu, mc)
import torch.nn.functional as F
lab = torch.randint(0, n, (b, hw * 16, hw * 16), generator=g)
up = F.interpolate(low_ref, size=(hw * 16, hw * 16), mode="bilinear", align_corners=False)
loss = F.cross_entropy(up, lab)
low_ref.retain_grad()
loss.backward()
wgt = low_ref.grad.clone()
eng = HeadEngine(HeadCfg(), precise=precise)
p = {k[len("decode_head."):]: v.detach().cuda() for k, v in pcpu.items()}
fin = [f.detach().permute(0, 2, 3, 1).contiguous().cuda() for f in feats + [emb]]
low, ctx = eng.forward(fin, text.cuda(), p, need_grad=True)
print("logits rel", rel(low, low_ref.detach()))
grads = {k: torch.zeros_like(v) for k, v in p.items()}
dfe = eng.backward(ctx, wgt.cuda(), p, grads)
for k, gv in grads.items():
    gr = pcpu["decode_head." + k].grad
    print(f"{k:50s} {rel(gv, gr) if gr is not None else -1:.2e}  norm {gr.norm().item() if gr is not None else 0:.3e}")
for i, (d, fr) in enumerate(zip(dfe, feats + [emb])):
    print(f"feat{i} {rel(d.permute(0, 3, 1, 2), fr.grad):.2e}")
