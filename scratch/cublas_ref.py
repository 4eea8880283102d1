"""What the library gets on the encoder's GEMM shapes (torch -> cuBLASLt), for orientation only (never on the product path)."""
import torch, torch.nn.functional as F
M = 16400
def t(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for name, n, k in (("qkv", 2304, 768), ("ffn1", 3072, 768), ("n768 k768", 768, 768), ("n768 k3072", 768, 3072), ("n768 k2304", 768, 2304)):
    a = torch.randn(M, k, device="cuda").bfloat16(); w = torch.randn(n, k, device="cuda").bfloat16(); b = torch.randn(n, device="cuda").bfloat16()
    us = t(lambda: F.linear(a, w, b))
    print(f"cublas {name:12s} bf16 linear+bias : {us:6.1f} us {2.0*M*n*k/us/1e6:7.1f} TF/s")
    if name == "ffn1":
        us = t(lambda: F.gelu(F.linear(a, w, b)))
        print(f"cublas {name:12s} linear + gelu kernel: {us:6.1f} us")
