"""block_n sweep of the K = 768 GEMM shapes of the encoder (out-proj with fp32 residual, QKV, FFN1 + GELU, FFN2 with residual)."""
import sys, torch
sys.path.insert(0, ".")
from semivl_b200 import lib as L, ops
L.check_device()
M = 16400
def t(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
g = torch.Generator(device="cuda").manual_seed(0)
def case(name, n, k, out_dtype, residual=False, act=L.ACT_NONE, preact=False):
    a = torch.randn(M, k, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(n, k, device="cuda", generator=g).to(torch.bfloat16)
    bias = torch.randn(n, device="cuda", generator=g)
    out = torch.empty(M, n, device="cuda", dtype=out_dtype)
    res = torch.randn(M, n, device="cuda", generator=g) if residual else None
    pre = torch.empty(M, n, device="cuda", dtype=torch.bfloat16) if preact else None
    for bn in (0, 64, 96, 128, 192, 256):
        try:
            us = t(lambda: ops.gemm(a, w, out, n=n, k=k, bias=bias, residual=res, act=act, preact_out=pre, block_n=bn,
                                    out_dtype=L.BF16 if out_dtype == torch.bfloat16 else None))
            print(f"{name:28s} block_n={bn:3d}: {us:7.1f} us  {2.0 * M * n * k / us / 1e6:7.1f} TF/s")
        except Exception as e:
            print(f"{name:28s} block_n={bn:3d}: failed {str(e)[:80]}")
case("out-proj n768 k768 f32+res", 768, 768, torch.float32, residual=True)
case("qkv n2304 k768 bf16", 2304, 768, torch.bfloat16)
case("ffn1 n3072 k768 gelu dsave", 3072, 768, torch.bfloat16, act=L.ACT_GELU_DSAVE, preact=True)
case("ffn2 n768 k3072 f32+res", 768, 3072, torch.float32, residual=True)
case("dgrad n768 k2304 bf16", 768, 2304, torch.bfloat16)
case("dgrad n768 k768 bf16", 768, 768, torch.bfloat16)
