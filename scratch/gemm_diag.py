"""Where does a GEMM tile's time go?  Needs a -DSVL_GEMM_DIAG build (SVL_NVCC_EXTRA=-DSVL_GEMM_DIAG python semivl_b200/build.py --force).
SVL_GEMM_DBG bits: 1 no global stores, 2 no epilogue body (accumulator released at once), 4 no side loads."""
import os, sys, torch
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
def case(name, n, k, out_dtype, residual=False, act=L.ACT_NONE, preact=False, dact=False, bn=0):
    a = torch.randn(M, k, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(n, k, device="cuda", generator=g).to(torch.bfloat16)
    bias = torch.randn(n, device="cuda", generator=g)
    out = torch.empty(M, n, device="cuda", dtype=out_dtype)
    res = torch.randn(M, n, device="cuda", generator=g) if residual else None
    pre = torch.empty(M, n, device="cuda", dtype=torch.bfloat16) if preact else None
    ds = torch.randn(M, n, device="cuda", generator=g).to(torch.bfloat16) if dact else None
    row = []
    for dbg in [int(x) for x in os.environ.get('DBGS', '0,1,4,5,2').split(',')]:
        os.environ["SVL_GEMM_DBG"] = str(dbg)
        kw = dict(dact_src=ds, dact_kind=L.ACT_SAVED) if dact else {}
        us = t(lambda: ops.gemm(a, w, out, n=n, k=k, bias=None if dact else bias, residual=res, act=act, preact_out=pre, block_n=bn,
                                out_dtype=L.BF16 if out_dtype == torch.bfloat16 else None, **kw))
        row.append(f"dbg{dbg}: {us:6.1f}")
    print(f"{name:30s} bn={bn:3d} " + "  ".join(row) + f"   ({2.0 * M * n * k / 1e6:.0f} MF)", flush=True)
for bn in (0, 128) if len(sys.argv) < 2 else (int(sys.argv[1]),):
    case("qkv n2304 k768 bf16", 2304, 768, torch.bfloat16, bn=bn)
    case("ffn1 n3072 k768 gelu dsave", 3072, 768, torch.bfloat16, act=L.ACT_GELU_DSAVE, preact=True, bn=bn)
    case("ffn2-dgrad n3072 k768 saved", 3072, 768, torch.bfloat16, dact=True, bn=bn)
    case("out-proj n768 k768 f32+res", 768, 768, torch.float32, residual=True, bn=bn)
    case("dgrad n768 k768 bf16", 768, 768, torch.bfloat16, bn=bn)
    case("ffn2 n768 k3072 f32+res", 768, 3072, torch.float32, residual=True, bn=bn)
    case("dgrad n768 k3072 bf16", 768, 3072, torch.bfloat16, bn=bn)
