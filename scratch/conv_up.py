"""Forward / data-gradient convolutions of the Up blocks at config-2 size (336 maps): 3x3, 32/64 channels."""
import os, sys, torch
sys.path.insert(0, ".")
from semivl_b200 import lib as L, ops
L.check_device()
filt = [(i - 1, j - 1) for i in range(3) for j in range(3)]
for (nb, h, w, cin, cout) in ((336, 128, 128, 32, 32), (336, 128, 128, 64, 32), (336, 128, 128, 32, 64), (336, 64, 64, 64, 64), (336, 64, 64, 128, 64)):
    x = torch.randn(nb * h * w, cin, device="cuda").bfloat16(); wt = torch.randn(9 * cout, cin, device="cuda").bfloat16()
    out = torch.empty(nb * h * w, cout, device="cuda", dtype=torch.bfloat16)
    f = lambda: ops.gemm(x, wt, out, n=cout, k=cin, conv=(nb, h, w), filt=filt, b_row_stride=cout)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    if os.environ.get("TRACE"):
        tr = torch.zeros(8, dtype=torch.int64, device="cuda"); os.environ["SVL_ROLL_TRACE"] = str(tr.data_ptr()); f(); torch.cuda.synchronize(); os.environ.pop("SVL_ROLL_TRACE")
        t = tr.cpu().tolist()
        if t[3]: print(f"   CTA 0: {t[3]} strips, {t[0] / t[3]:.0f} clk per strip in the MMA warp (waiting: {t[1] / t[3]:.0f} for a free accumulator block, {t[2] / t[3]:.0f} for the strip); producer waits {t[4] / t[3]:.0f} per strip for a free slot; quartet 0 waits {t[5] / max(t[6], 1):.0f} per block for a full block")
    print(f"conv {h}x{w} cin{cin} cout{cout}: {us:7.1f} us  {2.0 * nb * h * w * cout * cin * 9 / us / 1e6:7.1f} TF/s  {nb*h*w*(cin+cout)*2/us/1e6:.2f} TB/s in+out")
