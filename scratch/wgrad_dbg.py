import sys, torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from semivl_b200 import lib as L, ops
L.check_device()
nb, h, w, cin, cout = [int(a) for a in sys.argv[1:6]] if len(sys.argv) > 5 else (1, 4, 64, 32, 32)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(nb, cin, h, w, device="cuda", generator=g).bfloat16().float()
dy = torch.randn(nb, cout, h, w, device="cuda", generator=g).bfloat16().float()
wt = torch.zeros(cout, cin, 3, 3, device="cuda", dtype=torch.float64, requires_grad=True)
F.conv2d(x.double(), wt, None, padding=1).backward(dy.double())
ref = wt.grad.permute(2, 3, 0, 1).reshape(9, cout, cin).float()
xa = x.permute(0, 2, 3, 1).contiguous().bfloat16(); dya = dy.permute(0, 2, 3, 1).contiguous().bfloat16()
filt = [(i - 1, j - 1) for i in range(3) for j in range(3)]
dw = torch.zeros(9, cout, cin, device="cuda")
ops.wgrad(dya, xa, dw, m=cout, n=cin, conv=(nb, h, w), filt=filt)
torch.cuda.synchronize()
for s in range(9):
    e = (dw[s] - ref[s]).abs().max().item() / ref.abs().max().item()
    # does dw[s] match some other reference slot?
    best = min(range(9), key=lambda k: (dw[s] - ref[k]).abs().max().item())
    print(f"slot {s} (dy {filt[s][0]:+d} dx {filt[s][1]:+d}): rel err {e:.2e}; closest ref slot {best} err {(dw[s]-ref[best]).abs().max().item()/ref.abs().max().item():.2e}; |dw| {dw[s].abs().max().item():.3f} |ref| {ref[s].abs().max().item():.3f}")
# which x rows / pixel ranges contribute?  least squares of dw[slot] against per-(row, 16-pixel block) partial references
import itertools
xa32, dya32 = xa.float(), dya.float()
for s in (4, 7, 1):
    fy, fx = filt[s]
    parts = []
    for r in range(h):
        for pb in range(0, w, 16):
            acc = torch.zeros(cout, cin, device="cuda")
            for px in range(pb, min(pb + 16, w)):
                yy, xx = r + fy, px + fx
                if 0 <= yy < h and 0 <= xx < w:
                    acc += torch.outer(dya32[0, r, px], xa32[0, yy, xx])
            parts.append(acc.reshape(-1))
    A = torch.stack(parts, 1)
    coef = torch.linalg.lstsq(A, dw[s].reshape(-1, 1)).solution.reshape(h, -1)
    print(f"slot {s}: contribution coefficient per (dy row, 16-pixel block):")
    print(coef.cpu().numpy().round(2))
