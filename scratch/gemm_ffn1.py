import sys, torch
sys.path.insert(0, ".")
from semivl_b200 import lib as L, ops
L.check_device()
m, n, k = 16400, 3072, 768
a = torch.randn(m, k, device="cuda").to(torch.bfloat16); w = torch.randn(n, k, device="cuda").to(torch.bfloat16)
bias = torch.randn(n, device="cuda")
out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16); pre = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
for _ in range(5): ops.gemm(a, w, out, n=n, k=k, bias=bias, act=L.ACT_GELU, preact_out=pre)
torch.cuda.synchronize()
