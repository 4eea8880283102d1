import sys, torch
sys.path.insert(0, ".")
from semivl_b200 import lib as L, ops
L.check_device()
def t(m, n, k, odt=torch.bfloat16, bn=0, reps=10, **kw):
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16); w = torch.randn(n, k, device="cuda").to(torch.bfloat16)
    out = torch.empty(m, n, device="cuda", dtype=odt)
    for _ in range(3): ops.gemm(a, w, out, n=n, k=k, block_n=bn, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.gemm(a, w, out, n=n, k=k, block_n=bn, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tiles = ((m + 127) // 128) * ((n + (bn or 256) - 1) // (bn or 256))
    print(f"m{m} n{n} k{k} {str(odt)[6:]} bn{bn or 'auto'}: {ms*1e3:8.1f} us {2*m*n*k/ms/1e9:7.1f} TF/s  tiles {tiles} rounds {-(-tiles//148)} us/round {ms*1e3/(-(-tiles//148)):.1f}")
M = 148 * 128
for k in (64, 256, 768, 3072):
    t(M, 256, k)            # exactly one tile per SM
for k in (64, 768):
    t(M, 256, k, torch.float32)
    t(M, 128, k, bn=128)
    t(M, 2048, k)           # 8 tiles per SM
t(16400, 2304, 768); t(16400, 2304, 768, bn=128); t(16400, 768, 3072); t(16400, 768, 3072, bn=128); t(16400, 768, 3072, bn=192)
