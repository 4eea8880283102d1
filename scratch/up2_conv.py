"""up2-sized 3x3 convs (336 maps of 128x128, 32/64 channels): forward gemm + weight gradient, timed alone."""
import sys, torch
sys.path.insert(0, ".")
from semivl_b200 import lib as L, ops
L.check_device()
nb, h, w = 336, 128, 128
cin = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cout = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
x = torch.randn(nb, h, w, cin, device="cuda").to(torch.bfloat16)
dy = torch.randn(nb, h, w, cout, device="cuda").to(torch.bfloat16)
wt = (torch.randn(9 * cout, cin, device="cuda") / (9 * cin) ** 0.5).to(torch.bfloat16)
out = torch.empty(nb, h, w, cout, device="cuda", dtype=torch.bfloat16)
dw = torch.zeros(9, cout, cin, device="cuda")
filt = [(i - 1, j - 1) for i in range(3) for j in range(3)]
def t(f):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
fl = 2.0 * nb * h * w * cin * cout * 9
tg = t(lambda: ops.gemm(x, wt, out, n=cout, k=cin, conv=(nb, h, w), filt=filt, b_row_stride=cout))
tw = t(lambda: ops.wgrad(dy, x, dw, m=cout, n=cin, conv=(nb, h, w), filt=filt))
print(f"conv fwd cin{cin} cout{cout}: {tg:8.1f} us {fl / tg / 1e6:7.1f} TF/s   wgrad: {tw:8.1f} us {fl / tw / 1e6:7.1f} TF/s   (HBM floor fwd {nb*h*w*(cin+cout)*2/6.4e6:.0f} us)")
