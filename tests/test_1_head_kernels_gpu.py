"""Unit parity of the VLG head's HBM-bound kernels against plain PyTorch fp32 ops on the GPU."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12)).item()


@pytest.fixture(scope="module")
def L():
    from semivl_b200 import lib
    lib.check_device()
    return lib


@pytest.mark.parametrize("maps,hw,C,G", [(5, 64, 128, 8), (3, 4096, 32, 2), (7, 1, 128, 8), (2, 256, 64, 4)])
def test_groupnorm_relu(L, maps, hw, C, G):
    from semivl_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(maps)
    x = torch.randn(maps * hw, C, device="cuda", generator=g) * 1.5 + 0.3
    gamma = torch.randn(C, device="cuda", generator=g)
    beta = torch.randn(C, device="cuda", generator=g) * 0.3
    res = torch.randn(maps * hw, C, device="cuda", generator=g)
    dy = torch.randn(maps * hw, C, device="cuda", generator=g)
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    xin = xr.view(maps, hw, C).permute(0, 2, 1)                       # [maps, C, hw]
    ref = F.relu(F.group_norm(xin, G, gr, br, 1e-5)).permute(0, 2, 1).reshape(maps * hw, C) + res
    ref.backward(dy)
    out = torch.empty(maps * hw, C, device="cuda")
    mean, rstd = ops.gn_relu_fwd(x, L.F32, gamma, beta, out, L.F32, maps, hw, C, G, res=res, res_dtype=L.F32)
    assert _rel(out, ref) < 1e-5
    dx = torch.empty(maps * hw, C, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ops.gn_relu_bwd(dy, L.F32, x, L.F32, gamma, beta, mean, rstd, dx, L.F32, dg, db, maps, hw, C, G)
    assert _rel(dx, xr.grad) < 1e-4, _rel(dx, xr.grad)
    assert _rel(dg, gr.grad) < 1e-4 and _rel(db, br.grad) < 1e-4, (_rel(dg, gr.grad), _rel(db, br.grad))


def test_conv_out1(L):
    maps, h, w, C = 3, 20, 24, 32
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(maps, h, w, C, device="cuda", generator=g)
    wt = torch.randn(1, C, 3, 3, device="cuda", generator=g)
    bias = torch.randn(1, device="cuda", generator=g)
    dout = torch.randn(maps, h, w, device="cuda", generator=g)
    xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    ref = F.conv2d(xr.permute(0, 3, 1, 2), wr, br, padding=1)[:, 0]
    ref.backward(dout)
    wk = wt[0].permute(1, 2, 0).reshape(-1).contiguous()
    out = torch.empty(maps, h, w, device="cuda")
    L.call("svl_conv_out1_fwd", x, L.F32, C, wk, bias, out, maps, h, w, C)
    assert _rel(out, ref) < 1e-5
    dx = torch.empty_like(x)
    dw, dbias = torch.zeros(9 * C, device="cuda"), torch.zeros(1, device="cuda")
    L.call("svl_conv_out1_bwd", dout, x, L.F32, C, wk, dx, L.F32, C, dw, dbias, maps, h, w, C)
    assert _rel(dx, xr.grad) < 1e-5
    assert _rel(dw.view(3, 3, C).permute(2, 0, 1)[None], wr.grad) < 1e-4
    assert _rel(dbias, br.grad) < 1e-4


@pytest.mark.parametrize("B,N,h,w", [(2, 5, 8, 8), (1, 3, 14, 14), (2, 4, 5, 5)])
def test_pool_unpool_tokens(L, B, N, h, w):
    C, Ct, pool = 128, 128, 4
    hp, wp = h // pool, w // pool
    g = torch.Generator(device="cuda").manual_seed(h)
    x = torch.randn(B * N, h, w, C, device="cuda", generator=g)
    text = torch.randn(N, Ct, device="cuda", generator=g)
    tok = torch.empty(B * hp * wp * N, C + Ct, device="cuda")
    L.call("svl_pool_tokens", x, L.F32, C, text, tok, B, N, h, w, C, Ct, pool)
    xp = F.avg_pool2d(x.permute(0, 3, 1, 2), pool)                                      # [BN, C, hp, wp]
    ref = xp.reshape(B, N, C, hp, wp).permute(0, 3, 4, 1, 2).reshape(B * hp * wp, N, C)
    ref = torch.cat((ref, text[None].expand(B * hp * wp, N, Ct)), -1).reshape(-1, C + Ct)
    assert _rel(tok, ref) < 1e-6
    # bf16 input (vectorised kernel): exact average of the bf16-rounded values
    xb = x.to(torch.bfloat16)
    tok_b = torch.full_like(tok, 3.0)
    L.call("svl_pool_tokens", xb, L.BF16, C, text, tok_b, B, N, h, w, C, Ct, pool)
    xpb = F.avg_pool2d(xb.float().permute(0, 3, 1, 2), pool)
    refb = xpb.reshape(B, N, C, hp, wp).permute(0, 3, 4, 1, 2).reshape(B * hp * wp, N, C)
    refb = torch.cat((refb, text[None].expand(B * hp * wp, N, Ct)), -1).reshape(-1, C + Ct)
    assert _rel(tok_b, refb) < 1e-6
    # unpool + add
    tk = torch.randn(B * hp * wp * N, C + Ct, device="cuda", generator=g).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    t2 = tk[:, :C].reshape(B, hp, wp, N, C).permute(0, 3, 4, 1, 2).reshape(B * N, C, hp, wp)
    refo = xr + F.interpolate(t2, size=(h, w), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    dout = torch.randn(B * N, h, w, C, device="cuda", generator=g)
    refo.backward(dout)
    out = torch.empty(B * N, h, w, C, device="cuda")
    L.call("svl_unpool_add", x, L.F32, C, tk.detach(), C + Ct, out, L.F32, C, B, N, h, w, C, hp, wp)
    assert _rel(out, refo) < 1e-5
    out_b = torch.empty(B * N, h, w, C, device="cuda", dtype=torch.bfloat16)
    L.call("svl_unpool_add", xb, L.BF16, C, tk.detach(), C + Ct, out_b, L.BF16, C, B, N, h, w, C, hp, wp)
    refo_b = xb.float() + F.interpolate(t2.detach(), size=(h, w), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    assert _rel(out_b, refo_b) < 5e-3                      # bf16 output rounding
    dtok = torch.empty_like(tk)
    L.call("svl_unpool_bwd", dout, L.F32, C, dtok, C + Ct, B, N, h, w, C, hp, wp)
    assert _rel(dtok, tk.grad) < 1e-5
    # bf16 gradient input (vectorised kernel): exact on the bf16-rounded values
    dtok_b = torch.full_like(tk, 7.0)
    L.call("svl_unpool_bwd", dout.to(torch.bfloat16), L.BF16, C, dtok_b, C + Ct, B, N, h, w, C, hp, wp)
    tk.grad = None
    (xr + F.interpolate(t2, size=(h, w), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)).backward(dout.to(torch.bfloat16).float())
    assert _rel(dtok_b, tk.grad) < 1e-5
    # pool backward
    dtk = torch.randn(B * hp * wp * N, C + Ct, device="cuda", generator=g)
    xr2 = x.clone().requires_grad_(True)
    xp2 = F.avg_pool2d(xr2.permute(0, 3, 1, 2), pool).reshape(B, N, C, hp, wp).permute(0, 3, 4, 1, 2).reshape(-1, C)
    xp2.backward(dtk[:, :C].contiguous())
    dx = torch.ones(B * N, h, w, C, device="cuda")
    L.call("svl_pool_tokens_bwd", dtk, C + Ct, dx, B, N, h, w, C, pool)
    assert _rel(dx, xr2.grad + 1) < 1e-5
    # fused variant: dx = src + pooling gradient, src f32 or bf16, written once
    src = torch.randn(B * N, h, w, C, device="cuda", generator=g)
    for sv, dt in ((src, L.F32), (src.to(torch.bfloat16), L.BF16)):
        dx2 = torch.full((B * N, h, w, C), 5.0, device="cuda")
        L.call("svl_pool_tokens_bwd_from", sv, dt, C, dtk, C + Ct, dx2, B, N, h, w, C, pool)
        assert _rel(dx2, sv.float() + xr2.grad) < 1e-5


@pytest.mark.parametrize("B,N,h,w,Cs", [(2, 3, 8, 8, 32), (1, 4, 5, 5, 16)])
def test_skip_fill_grad(L, B, N, h, w, Cs):
    H2, W2, cup = 2 * h, 2 * w, 96
    g = torch.Generator(device="cuda").manual_seed(h)
    skip = torch.relu(torch.randn(B, h, w, Cs, device="cuda", generator=g))
    sr = skip.clone().requires_grad_(True)
    up = F.interpolate(sr.permute(0, 3, 1, 2), size=(H2, W2), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    ref = up.repeat_interleave(N, dim=0)
    cat = torch.zeros(B * N, H2, W2, cup + Cs, device="cuda")
    L.call("svl_skip_fill", skip, L.F32, Cs, cat, L.F32, cup + Cs, cup, B, N, h, w, Cs, H2, W2)
    assert _rel(cat[..., cup:], ref) < 1e-5 and cat[..., :cup].abs().max() == 0
    cat_b = torch.zeros(B * N, H2, W2, cup + Cs, device="cuda", dtype=torch.bfloat16)
    L.call("svl_skip_fill", skip, L.F32, Cs, cat_b, L.BF16, cup + Cs, cup, B, N, h, w, Cs, H2, W2)
    assert _rel(cat_b[..., cup:], ref) < 5e-3 and cat_b[..., :cup].abs().max() == 0
    dcat = torch.randn(B * N, H2, W2, cup + Cs, device="cuda", generator=g)
    ref.backward(dcat[..., cup:])
    dsk = torch.empty(B, h, w, Cs, device="cuda")
    L.call("svl_skip_grad", dcat, L.F32, cup + Cs, cup, skip, L.F32, Cs, dsk, L.F32, Cs, B, N, h, w, Cs, H2, W2)
    assert _rel(dsk, sr.grad * (skip > 0)) < 1e-5


def test_sim_im2col(L):
    B, N, h, w, ks = 2, 5, 9, 9, 7
    g = torch.Generator(device="cuda").manual_seed(0)
    sim = torch.randn(B * h * w, 8, device="cuda", generator=g)
    col = torch.empty(B * N * h * w, 64, device="cuda")
    L.call("svl_sim_im2col", sim, 8, col, L.F32, 64, B, N, h, w, ks, 64)
    maps = sim[:, :N].reshape(B, h, w, N).permute(0, 3, 1, 2).reshape(B * N, 1, h, w)
    ref = F.unfold(maps, ks, padding=ks // 2).transpose(1, 2).reshape(B * N * h * w, ks * ks)
    assert _rel(col[:, :49], ref) < 1e-6 and col[:, 49:].abs().max() == 0
    dcol = torch.randn(B * N * h * w, 64, device="cuda", generator=g)
    dsim = torch.empty(B * h * w, 16, device="cuda")
    L.call("svl_sim_col2im", dcol, L.F32, 64, dsim, L.F32, 16, 16, B, N, h, w, ks)
    refd = F.fold(dcol[:, :49].reshape(B * N, h * w, 49).transpose(1, 2), (h, w), ks, padding=ks // 2)     # [BN,1,h,w]
    refd = refd.reshape(B, N, h, w).permute(0, 2, 3, 1).reshape(B * h * w, N)
    assert _rel(dsim[:, :N], refd) < 1e-5 and dsim[:, N:].abs().max() == 0


def test_map_sum_bcast(L):
    maps, hw, C = 5, 100, 128
    x = torch.randn(maps, hw, C, device="cuda")
    out = torch.empty(maps, C, device="cuda")
    L.call("svl_map_sum", x, L.F32, C, out, maps, hw, C, 0.01)
    assert _rel(out, x.sum(1) * 0.01) < 1e-5
    y = x.clone()
    L.call("svl_map_bcast_add", y, out, L.F32, C, maps, hw, C, 2.0)
    assert _rel(y, x + 2 * out[:, None]) < 1e-6
    # bf16 input (vectorised, 16 pixel lanes x 16 channel vectors), ragged pixel counts
    for hw2 in (100, 1024, 7):
        xb = torch.randn(maps, hw2, C, device="cuda").to(torch.bfloat16)
        out_b = torch.full((maps, C), 9.0, device="cuda")
        L.call("svl_map_sum", xb, L.BF16, C, out_b, maps, hw2, C, 0.5)
        assert _rel(out_b, xb.float().sum(1) * 0.5) < 1e-5


@pytest.mark.parametrize("maps,h,w", [(3, 20, 24), (2, 128, 128), (5, 9, 70)])
def test_conv_out1_bf16_fast_path(L, maps, h, w):
    """the all-bf16 kernels of the 32 -> 1 output conv (ragged tiles, full-size maps) against F.conv2d on the bf16-rounded input"""
    C = 32
    g = torch.Generator(device="cuda").manual_seed(1)
    xb = torch.randn(maps, h, w, C, device="cuda", generator=g).to(torch.bfloat16)
    x = xb.float()
    wt = torch.randn(1, C, 3, 3, device="cuda", generator=g)
    bias = torch.randn(1, device="cuda", generator=g)
    dout = torch.randn(maps, h, w, device="cuda", generator=g)
    xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    ref = F.conv2d(xr.permute(0, 3, 1, 2), wr, br, padding=1)[:, 0]
    ref.backward(dout)
    wk = wt[0].permute(1, 2, 0).reshape(-1).contiguous()
    out = torch.empty(maps, h, w, device="cuda")
    L.call("svl_conv_out1_fwd", xb, L.BF16, C, wk, bias, out, maps, h, w, C)
    assert _rel(out, ref) < 1e-5
    dx = torch.empty(maps, h, w, C, device="cuda", dtype=torch.bfloat16)
    dw, dbias = torch.zeros(9 * C, device="cuda"), torch.zeros(1, device="cuda")
    L.call("svl_conv_out1_bwd", dout, xb, L.BF16, C, wk, dx, L.BF16, C, dw, dbias, maps, h, w, C)
    assert _rel(dx, xr.grad) < 5e-3                         # bf16 output rounding
    assert _rel(dw.view(3, 3, C).permute(2, 0, 1)[None], wr.grad) < 1e-4
    assert _rel(dbias, br.grad) < 1e-4


@pytest.mark.parametrize("B,N,P,Cs,c0,ld", [(2, 21, 300, 32, 96, 128), (1, 5, 77, 16, 48, 64), (3, 4, 64, 8, 0, 8)])
def test_class_sum(L, B, N, P, Cs, c0, ld):
    """out[b, p, c] = sum_n x[(b, n), p, c0 + c]: generic (f32) and the batched-load bf16 path"""
    g = torch.Generator(device="cuda").manual_seed(P)
    x = torch.randn(B * N, P, ld, device="cuda", generator=g)
    for t, dt, tol in ((x, L.F32, 1e-6), (x.to(torch.bfloat16), L.BF16, 1e-6)):
        out = torch.full((B, P, Cs), 4.0, device="cuda")
        L.call("svl_class_sum", t, dt, ld, c0, out, B, N, P, Cs)
        ref = t.float().reshape(B, N, P, ld)[..., c0:c0 + Cs].sum(1)
        assert _rel(out, ref) < tol
