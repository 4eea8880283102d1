import os, sys, numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_model_gpu import _build
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
name = sys.argv[1]
g = dict(np.load(f"tests/golden/{name}.npz", allow_pickle=False))
crop = int(g["crop"])
m, mc, sd = _build(crop, True, int(g["nclass"]))
img = torch.from_numpy(g["img"]).cuda(); lab = torch.from_numpy(g["label"].astype(np.int64)).cuda()
m.train()
y = m(img)
loss = F.cross_entropy(y, lab, ignore_index=255)
loss.backward()
named = dict(m.named_parameters())
# oracle grads in fp32 and fp64
from oracle import semivl_oracle as O
text = torch.from_numpy(np.load("semivl_b200/configs/_base_/datasets/text_embedding/voc12_wbg_single.npy"))
res = {}
for dt in (torch.float32, torch.float64):
    p = {k: v.clone().to(dt).requires_grad_(True) for k, v in sd.items()}
    yo = O.model_forward(img.cpu().to(dt), p, text, mc)
    F.cross_entropy(yo, lab.cpu(), ignore_index=255).backward()
    res[dt] = p
for nme, norm in zip(g["grad_names"], g["grad_norms"]):
    k = str(nme)
    gr = named[k].grad
    o32, o64 = res[torch.float32][k].grad, res[torch.float64][k].grad
    rel = lambda a, b: ((a.double().cpu() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()
    print(f"{k:55s} refnorm {norm:.4e} mine {gr.double().norm().item():.4e} o32 {o32.double().norm().item():.4e} o64 {o64.norm().item():.4e}  rel(mine,o64) {rel(gr, o64):.2e} rel(o32,o64) {rel(o32, o64):.2e}")
