import sys, torch
sys.path.insert(0, ".")
from semivl_b200 import lib as L, ops
L.check_device()
def timeit(fn, nbytes, name, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:38s} {ms*1e3:9.1f} us  {nbytes/ms/1e6:8.1f} GB/s")
rows, c = 16400 * 4, 768
x = torch.randn(rows, c, device="cuda"); xb = x.to(torch.bfloat16)
y = torch.empty_like(x); yb = torch.empty_like(xb)
timeit(lambda: y.copy_(x), rows*c*8, "torch copy f32")
timeit(lambda: ops.cast(x, L.F32, y, L.F32, rows, c), rows*c*8, "svl_cast f32->f32")
timeit(lambda: ops.cast(x, L.F32, yb, L.BF16, rows, c), rows*c*6, "svl_cast f32->bf16")
g = torch.ones(c, device="cuda"); bt = torch.zeros(c, device="cuda")
mean = torch.empty(rows, device="cuda"); rstd = torch.empty(rows, device="cuda")
timeit(lambda: L.call("svl_layernorm_fwd", x, c, g, bt, yb, L.BF16, c, mean, rstd, rows, c, 1e-6), rows*c*6, "layernorm_fwd f32->bf16")
dx = torch.empty_like(x); dxa = torch.empty_like(xb)
timeit(lambda: L.call("svl_layernorm_bwd", xb, L.BF16, c, x, c, g, mean, rstd, y, None, dx, dxa, L.BF16, c, None, None, rows, c), rows*c*(2+4+4+4+2), "layernorm_bwd (dy bf16,x,dres->dx,act)")
out = torch.zeros(c, device="cuda")
timeit(lambda: ops.colsum(xb, L.BF16, rows, c, out), rows*c*2, "colsum bf16")
maps, hw, C, G = 336, 16384, 32, 2
xm = torch.randn(maps*hw, C, device="cuda").to(torch.bfloat16); ym = torch.empty_like(xm); dm = torch.randn(maps*hw, C, device="cuda").to(torch.bfloat16)
ga = torch.ones(C, device="cuda"); be = torch.zeros(C, device="cuda")
def gnf():
    return ops.gn_relu_fwd(xm, L.BF16, ga, be, ym, L.BF16, maps, hw, C, G)
m_, r_ = gnf()
timeit(gnf, maps*hw*C*(2+2+2), "gn_relu_fwd 336x128x128x32 (3 kernels)")
dg = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda")
timeit(lambda: ops.gn_relu_bwd(dm, L.BF16, xm, L.BF16, ga, be, m_, r_, ym, L.BF16, dg, db, maps, hw, C, G), maps*hw*C*(4+4+2), "gn_relu_bwd (2 kernels)")
n = 31_350_000
p = torch.randn(n, device="cuda"); gg = torch.randn(n, device="cuda"); m1 = torch.zeros(n, device="cuda"); v1 = torch.zeros(n, device="cuda")
timeit(lambda: L.call("svl_adamw", p, gg, m1, v1, n, 1e-4, 0.9, 0.999, 1e-8, 0.01, 1, 1.0), n*28, "adamw 31.35M")
