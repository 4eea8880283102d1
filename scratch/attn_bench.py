import sys, torch
sys.path.insert(0, ".")
from semivl_b200 import lib as L, ops
L.check_device()
b, Lq, H = 16, 1025, 12
E = H * 64
qkv = torch.randn(b * Lq, 3 * E, device="cuda").to(torch.bfloat16)
dout = torch.randn(b * Lq, E, device="cuda").to(torch.bfloat16)
def t(fn, name, flops, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:20s} {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TF/s")
unit = 2.0 * Lq * Lq * 64 * H * b
out, lse = ops.attention_fwd(qkv, b, Lq, H, False)
t(lambda: ops.attention_fwd(qkv, b, Lq, H, False), "attn fwd", 2 * unit)
t(lambda: ops.attention_bwd(qkv, out, dout, lse, b, Lq, H, False), "attn bwd", 5 * unit)
