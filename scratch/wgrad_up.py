"""Weight gradients of the 32-channel Up convolutions at config-2 size (336 maps of 128 x 128)."""
import sys, torch
sys.path.insert(0, ".")
from semivl_b200 import lib as L, ops
L.check_device()
nb, h, w = 336, 128, 128
filt = [(i - 1, j - 1) for i in range(3) for j in range(3)]
for cout, cin in ((32, 32), (32, 64)):
    dy = torch.randn(nb * h * w, cout, device="cuda").bfloat16(); x = torch.randn(nb * h * w, cin, device="cuda").bfloat16()
    dw = torch.zeros(9, cout, cin, device="cuda")
    f = lambda: ops.wgrad(dy, x, dw, m=cout, n=cin, conv=(nb, h, w), filt=filt)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    gb = nb * h * w * (cout + cin) * 2 / 1e9
    print(f"wgrad cout{cout} cin{cin}: {us:7.1f} us  {2.0 * nb * h * w * cout * cin * 9 / us / 1e6:7.1f} TF/s  {gb / us * 1e6 / 1e3:.2f} TB/s of operand bytes")
