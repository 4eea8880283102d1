"""Per-tile epilogue cost: plain bf16 GEMM (M = 148 * 128 * 4, N = 2048: 32 tiles per SM) at K = 64 .. 768, SVL_GEMM_CLUSTER=0."""
import sys, torch
sys.path.insert(0, ".")
from semivl_b200 import lib as L, ops
L.check_device()
m, n = 148 * 128 * 4, 2048
for k in (64, 128, 256, 512, 768):
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16); w = torch.randn(n, k, device="cuda").to(torch.bfloat16)
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(a, w, out, n=n, k=k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.gemm(a, w, out, n=n, k=k)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    tiles_per_sm = (m // 128) * (n // 256) / 148
    print(f"K={k:4d}: {us:8.1f} us  {2.0*m*n*k/us/1e6:7.1f} TF/s  {us*1.965e3/tiles_per_sm:8.0f} cycles per 128x256 tile ({k//64} k-blocks)  out {m*n*2/us/1e6:.2f} TB/s")
