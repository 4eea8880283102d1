"""Per-tile timeline of CTA 0 of one GEMM launch (diag build): MMA warp (tile start / accumulator free / last MMA issued) and three
epilogue warps (loop top / accumulator full / tile done), in cycles relative to the first event."""
import os, sys, torch
sys.path.insert(0, ".")
from semivl_b200 import lib as L, ops
L.check_device()
M = 16400
g = torch.Generator(device="cuda").manual_seed(0)
def case(name, n, k, out_dtype, residual=False, act=L.ACT_NONE, preact=False, dact=False, bn=0):
    a = torch.randn(M, k, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(n, k, device="cuda", generator=g).to(torch.bfloat16)
    bias = torch.randn(n, device="cuda", generator=g)
    out = torch.empty(M, n, device="cuda", dtype=out_dtype)
    res = torch.randn(M, n, device="cuda", generator=g) if residual else None
    pre = torch.empty(M, n, device="cuda", dtype=torch.bfloat16) if preact else None
    ds = torch.randn(M, n, device="cuda", generator=g).to(torch.bfloat16) if dact else None
    kw = dict(dact_src=ds, dact_kind=L.ACT_SAVED) if dact else {}
    fn = lambda: ops.gemm(a, w, out, n=n, k=k, bias=None if dact else bias, residual=res, act=act, preact_out=pre, block_n=bn,
                          out_dtype=L.BF16 if out_dtype == torch.bfloat16 else None, **kw)
    os.environ.pop("SVL_GEMM_TRACE", None)
    for _ in range(3): fn()
    tr = torch.zeros(32 * 16, dtype=torch.int64, device="cuda")
    os.environ["SVL_GEMM_TRACE"] = str(tr.data_ptr())
    fn(); torch.cuda.synchronize()
    os.environ.pop("SVL_GEMM_TRACE", None)
    t = tr.cpu().view(32, 16)
    t0 = int(t[0, 0])
    print(f"== {name} (dbg {os.environ.get('SVL_GEMM_DBG', '0')})")
    print("tile | mma: top  accfree  issued | epi w2: top full done | w9: top full done | w17: top full done")
    for i in range(32):
        if t[i, 0] == 0 and t[i, 4] == 0: break
        r = [(int(x) - t0) if x else -1 for x in t[i]]
        if os.environ.get("DETAIL"): print(f"{i:4d} | w2: top {r[4]} full {r[5]} side0+ld {r[3]} staged0 {r[7]} side1+ld {r[11]} done {r[6]}")
        print(f"{i:4d} | {r[0]:7d} {r[1]:7d} {r[2]:7d} | {r[4]:7d} {r[5]:7d} {r[6]:7d} | {r[8]:7d} {r[9]:7d} {r[10]:7d} | {r[12]:7d} {r[13]:7d} {r[14]:7d} | mma waited {int(t[i, 15])}")
def convt():
    global M
    nb, h, w, cin, cup, ldc = 336, 64, 64, 64, 48, 64
    x = torch.randn(nb * h * w, cin, device="cuda").bfloat16(); wt = torch.randn(4 * cup, cin, device="cuda").bfloat16()
    bias = torch.randn(4 * cup, device="cuda")
    out = torch.empty(nb * 4 * h * w, ldc, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.gemm(x, wt, out, n=4 * cup, k=cin, bias=bias, out_mode=L.OUT_CONVT2X2, out_hw=(h, w), out_dtype=L.BF16, m=nb * h * w)
    for _ in range(3): fn()
    tr = torch.zeros(32 * 16, dtype=torch.int64, device="cuda")
    os.environ["SVL_GEMM_TRACE"] = str(tr.data_ptr())
    fn(); torch.cuda.synchronize()
    os.environ.pop("SVL_GEMM_TRACE", None)
    t = tr.cpu().view(32, 16)
    t0 = int(t[0, 0])
    print("== convT 64x64 cin64 cup48 (first 32 tiles of CTA 0)")
    print("tile | mma: top  accfree  issued | epi w2: top full done | w9: top full done | w17: top full done")
    for i in range(20):
        r = [(int(x_) - t0) if x_ else -1 for x_ in t[i]]
        print(f"{i:4d} | {r[0]:7d} {r[1]:7d} {r[2]:7d} | {r[4]:7d} {r[5]:7d} {r[6]:7d} | {r[8]:7d} {r[9]:7d} {r[10]:7d} | {r[12]:7d} {r[13]:7d} {r[14]:7d} | mma waited {int(t[i, 15])}")
which = sys.argv[1] if len(sys.argv) > 1 else "qkv"
if which == "convt": convt()
if which == "qkv": case("qkv n2304 k768 bf16", 2304, 768, torch.bfloat16)
if which == "ffn1": case("ffn1 n3072 k768 gelu dsave", 3072, 768, torch.bfloat16, act=L.ACT_GELU_DSAVE, preact=True)
if which == "ffn2d": case("ffn2-dgrad n3072 k768 saved", 3072, 768, torch.bfloat16, dact=True)
if which == "outproj": case("out-proj n768 k768 f32+res", 768, 768, torch.float32, residual=True)
if which == "ffn2": case("ffn2 n768 k3072 f32+res", 768, 3072, torch.float32, residual=True)
