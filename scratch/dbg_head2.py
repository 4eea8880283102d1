import sys, numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from oracle import semivl_oracle as O
from semivl_b200 import lib
from semivl_b200.engine.head import HeadCfg, HeadEngine
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
def rel(a, b): return ((a.double().cpu() - b.double().cpu()).abs().max() / (b.double().abs().max() + 1e-30)).item()
hw, b, n, precise = int(sys.argv[1]), 2, 21, sys.argv[2] == "1"
mc = O.ModelCfg(img_size=hw * 16, num_classes=n)
sd = O.fixture_state_dict(O.param_shapes(mc, with_clip_encoder=False), seed=0)
g = torch.Generator().manual_seed(hw)
f0 = [torch.randn(b, 768, hw, hw, generator=g), torch.randn(b, 768, hw, hw, generator=g)]
e0 = torch.randn(b, 512, hw, hw, generator=g); e0 = e0 / e0.norm(dim=1, keepdim=True)
text = torch.from_numpy(np.load("semivl_b200/configs/_base_/datasets/text_embedding/voc12_wbg_single.npy"))
lab = torch.randint(0, n, (b, hw * 16, hw * 16), generator=g)
res = {}
for dt in (torch.float64, torch.float32):
    pc = {k: v.clone().to(dt).requires_grad_(True) for k, v in sd.items() if k.startswith("decode_head.")}
    feats = [f.clone().to(dt).requires_grad_(True) for f in f0 + [e0]]
    low = O.vlg_head_forward(feats, text, pc, mc)
    lab_low = lab[:, ::4, ::4]
    wl = (F.one_hot(lab_low, n).permute(0, 3, 1, 2) * (0.5 + torch.rand(lab_low.shape, generator=torch.Generator().manual_seed(1)))[:, None]).to(dt)
    loss = (low * wl).sum()
    low.retain_grad(); loss.backward()
    res[dt] = (low.detach(), low.grad.clone(), {k: v.grad for k, v in pc.items()}, [f.grad for f in feats])
l64, dl64, g64, fg64 = res[torch.float64]
l32, dl32, g32, fg32 = res[torch.float32]
eng = HeadEngine(HeadCfg(), precise=precise)
p = {k[len("decode_head."):]: v.cuda() for k, v in sd.items() if k.startswith("decode_head.")}
fin = [f.permute(0, 2, 3, 1).contiguous().cuda() for f in f0 + [e0]]
low, ctx = eng.forward(fin, text.cuda(), p, need_grad=True)
print("logits rel: mine", rel(low, l64), " oracle32", rel(l32, l64))
grads = {k: torch.zeros_like(v) for k, v in p.items()}
dfe = eng.backward(ctx, dl64.float().cuda(), p, grads)
for k, gv in grads.items():
    r64 = g64["decode_head." + k]
    print(f"{k:50s} mine {rel(gv, r64):.2e}  oracle32 {rel(g32['decode_head.' + k], r64):.2e}")
for i, d in enumerate(dfe):
    print(f"feat{i} mine {rel(d.permute(0, 3, 1, 2), fg64[i]):.2e} oracle32 {rel(fg32[i], fg64[i]):.2e}")
