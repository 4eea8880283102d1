"""Transposed 2x2 convolutions of the Up blocks at config-2 size (forward GEMM with the 2x2 scatter epilogue)."""
import sys, torch
sys.path.insert(0, ".")
from semivl_b200 import lib as L, ops
L.check_device()
for (nb, h, w, cin, cup, ldc) in ((336, 64, 64, 64, 48, 64), (336, 32, 32, 128, 96, 128)):
    x = torch.randn(nb * h * w, cin, device="cuda").bfloat16(); wt = torch.randn(4 * cup, cin, device="cuda").bfloat16()
    bias = torch.randn(4 * cup, device="cuda")
    out = torch.empty(nb * 4 * h * w, ldc, device="cuda", dtype=torch.bfloat16)
    f = lambda: ops.gemm(x, wt, out, n=4 * cup, k=cin, bias=bias, out_mode=L.OUT_CONVT2X2, out_hw=(h, w), out_dtype=L.BF16, m=nb * h * w)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"convT {h}x{w} cin{cin} cup{cup}: {us:7.1f} us  {(nb*h*w*cin*2 + nb*4*h*w*cup*2)/us/1e6:.2f} TB/s in+out")
