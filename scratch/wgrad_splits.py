"""Split-K sweep of the ViT weight-gradient shapes (rows = 16400): how much of a launch is the red.add epilogue of its partial tiles?"""
import sys, torch
sys.path.insert(0, ".")
from semivl_b200 import lib as L, ops
L.check_device()
rows = 16400
for (m, n) in ((2304, 768), (768, 768), (768, 2304), (3072, 768)):
    dy = torch.randn(rows, m, device="cuda").bfloat16(); x = torch.randn(rows, n, device="cuda").bfloat16()
    dw = torch.zeros(m, n, device="cuda")
    out = []
    for s in (0, 1, 2, 3, 4, 5, 6, 8, 10, 12, 16):
        f = lambda: ops.wgrad(dy, x, dw, m=m, n=n, splits=s)
        for _ in range(3): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        out.append(f"s{s}: {e0.elapsed_time(e1) * 100:.1f}")
    print(f"wgrad m{m} n{n}: " + "  ".join(out) + f"   ({2.0 * rows * m * n / 1e6:.0f} MF)")
