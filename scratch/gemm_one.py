import sys, torch
sys.path.insert(0, ".")
from semivl_b200 import lib as L, ops
L.check_device()
m, n, k = 148 * 128, 2048, int(sys.argv[1]) if len(sys.argv) > 1 else 768
a = torch.randn(m, k, device="cuda").to(torch.bfloat16); w = torch.randn(n, k, device="cuda").to(torch.bfloat16)
out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
for _ in range(5): ops.gemm(a, w, out, n=n, k=k)
torch.cuda.synchronize()
