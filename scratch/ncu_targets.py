"""Standalone launches of the step's dominant kernels at BASELINE config-2 shapes, for `ncu --set full --profile-from-start off`.
Each target runs 3x unprofiled (warm-up) and then 2x inside cudaProfilerStart/Stop.  Also prints CUDA-event timings."""
import sys, torch
sys.path.insert(0, ".")
from semivl_b200 import lib as L, ops
L.check_device()
dev = "cuda"
M, E = 16400, 768
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
x = bf(M, E); h = bf(M, 4 * E)
w_qkv, w_out, w1, w2 = bf(3 * E, E), bf(E, E), bf(4 * E, E), bf(E, 4 * E)
res = torch.randn(M, E, device=dev)
b3, b1, b4 = torch.randn(3 * E, device=dev), torch.randn(E, device=dev), torch.randn(4 * E, device=dev)
o_qkv = torch.empty(M, 3 * E, device=dev, dtype=torch.bfloat16)
o_f32 = torch.empty(M, E, device=dev)
o_h = torch.empty(M, 4 * E, device=dev, dtype=torch.bfloat16); pre = torch.empty_like(o_h)
o_bf = torch.empty(M, E, device=dev, dtype=torch.bfloat16)
dw = torch.zeros(3 * E, E, device=dev)
targets = {
    "gemm_qkv": lambda: ops.gemm(x, w_qkv, o_qkv, n=3 * E, k=E, bias=b3),
    "gemm_outproj_f32res": lambda: ops.gemm(x, w_out, o_f32, n=E, k=E, bias=b1, residual=res),
    "gemm_ffn1_gelu": lambda: ops.gemm(x, w1, o_h, n=4 * E, k=E, bias=b4, act=L.ACT_GELU_DSAVE, preact_out=pre),
    "gemm_ffn2_f32res": lambda: ops.gemm(h, w2, o_f32, n=E, k=4 * E, bias=b1, residual=res),
    "gemm_ffn2_dgrad_dact": lambda: ops.gemm(x, w1, o_h, n=4 * E, k=E, dact_src=pre, dact_kind=L.ACT_SAVED),
    "gemm_ffn1_dgrad": lambda: ops.gemm(h, w2, o_bf, n=E, k=4 * E),
    "wgrad_inproj": lambda: ops.wgrad(o_qkv, x, dw, m=3 * E, n=E),
}
qkv = bf(M, 3 * E) * 0.5
att, lse = ops.attention_fwd(qkv, 16, 1025, 12, False)
datt = bf(M, E)
targets["attn_fwd"] = lambda: ops.attention_fwd(qkv, 16, 1025, 12, False)
targets["attn_bwd"] = lambda: ops.attention_bwd(qkv, att, datt, lse, 16, 1025, 12, False)
g = torch.ones(E, device=dev); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
xf = torch.randn(M, E, device=dev); dx = torch.empty_like(xf); dxa = torch.empty(M, E, device=dev, dtype=torch.bfloat16)
L.call("svl_layernorm_fwd", xf, E, g, b1, o_bf, L.BF16, E, mean, rstd, M, E, 1e-6)
targets["layernorm_bwd"] = lambda: L.call("svl_layernorm_bwd", x, L.BF16, E, xf, E, g, mean, rstd, res, None, dx, dxa, L.BF16, E, None, None, M, E)
cs = torch.zeros(3 * E, device=dev)
targets["colsum_qkv"] = lambda: ops.colsum(o_qkv, L.BF16, M, 3 * E, cs)
maps, hw, C, G = 336, 16384, 32, 2
xm = bf(maps * hw, C); ym = torch.empty_like(xm); dm = bf(maps * hw, C)
ga = torch.ones(C, device=dev); be = torch.zeros(C, device=dev)
m_, r_ = ops.gn_relu_fwd(xm, L.BF16, ga, be, ym, L.BF16, maps, hw, C, G)
dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
targets["gn_fwd_336x128x128x32"] = lambda: ops.gn_relu_fwd(xm, L.BF16, ga, be, ym, L.BF16, maps, hw, C, G)
targets["gn_bwd_336x128x128x32"] = lambda: ops.gn_relu_bwd(dm, L.BF16, xm, L.BF16, ga, be, m_, r_, ym, L.BF16, dg, db, maps, hw, C, G)
u2 = bf(maps * hw, C); low = torch.empty(16, 21, 128, 128, device=dev); wk = torch.randn(9 * C, device=dev); ob = torch.zeros(1, device=dev)
dlow = torch.randn(16, 21, 128, 128, device=dev); du2 = torch.empty_like(u2); dw9 = torch.zeros(9 * C, device=dev); dbo = torch.zeros(1, device=dev)
targets["conv_out1_fwd"] = lambda: L.call("svl_conv_out1_fwd", u2, L.BF16, C, wk, ob, low, maps, 128, 128, C)
targets["conv_out1_bwd"] = lambda: L.call("svl_conv_out1_bwd", dlow, u2, L.BF16, C, wk, du2, L.BF16, C, dw9, dbo, maps, 128, 128, C, n_launch=2)
import ctypes as Cc
lab = torch.randint(0, 21, (16, 512, 512), device=dev); coef = torch.full((1,), 1e-6, device=dev); lossb = torch.zeros(3, device=dev); dl2 = torch.zeros_like(low)
arr = lambda t: (Cc.c_void_p * 3)(t.data_ptr() if t is not None else None, None, None)
targets["upsample_ce"] = lambda: L.call("svl_upsample_ce", low, dl2, 16, 21, 128, 128, 512, 512, 1, arr(lab), arr(None), arr(coef), lossb, 1.0, 255)
f3 = [(i - 1, j - 1) for i in range(3) for j in range(3)]
xc32 = bf(336 * 128 * 128, 32); wc32 = bf(9 * 32, 32); oc32 = torch.empty(336 * 128 * 128, 32, device=dev, dtype=torch.bfloat16)
targets["conv_roll_128x128_c32_c32"] = lambda: ops.gemm(xc32, wc32, oc32, n=32, k=32, conv=(336, 128, 128), filt=f3, b_row_stride=32)
xc64 = bf(336 * 64 * 64, 64); wc64 = bf(9 * 64, 64); oc64 = torch.empty(336 * 64 * 64, 64, device=dev, dtype=torch.bfloat16)
targets["conv_roll_dual_64x64_c64_c64"] = lambda: ops.gemm(xc64, wc64, oc64, n=64, k=64, conv=(336, 64, 64), filt=f3, b_row_stride=64)
dwc = torch.zeros(9, 32, 32, device=dev)
targets["wgrad_rowstack_c32_c32"] = lambda: ops.wgrad(oc32, xc32, dwc, m=32, n=32, conv=(336, 128, 128), filt=f3)
sel = sys.argv[1:] or list(targets)
for name in sel:
    fn = targets[name]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1) / 5 * 1e3:9.1f} us", flush=True)
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
