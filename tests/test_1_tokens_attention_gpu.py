"""GPU parity of the token / normalisation kernels and of flash attention (fwd + bwd) against plain PyTorch fp32."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="module")
def ops():
    from semivl_b200 import lib, ops
    lib.check_device()
    return ops


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12)).item()


def _val(t, precise):
    """operand tensor -> f32 value"""
    if precise:
        c = t.shape[-1] // 2
        return t[..., :c].float() + t[..., c:].float()
    return t.float()


@pytest.mark.parametrize("precise", [False, True])
@pytest.mark.parametrize("H,W", [(64, 64), (72, 72), (224, 224)])
def test_patchify_assemble(ops, precise, H, W):
    g = torch.Generator(device="cuda").manual_seed(0)
    b = 2
    img = torch.randn(b, 3, H, W, device="cuda", generator=g)
    out, hp, wp = ops.patchify(img, precise)
    pad = F.pad(img, [0, wp * 16 - W, 0, hp * 16 - H])
    ref = F.unfold(pad, 16, stride=16).transpose(1, 2).reshape(b * hp * wp, 768)
    assert _rel(_val(out, precise), ref) < (1e-5 if precise else 5e-3)
    patches = torch.randn(b * hp * wp, 768, device="cuda", generator=g)
    cls = torch.randn(768, device="cuda", generator=g)
    pos = torch.randn(hp * wp + 1, 768, device="cuda", generator=g)
    x = ops.assemble_tokens(patches, cls, pos, b, hp * wp)
    refx = torch.cat((cls.expand(b, 1, 768), patches.view(b, hp * wp, 768)), 1) + pos
    assert torch.equal(x, refx)


@pytest.mark.parametrize("rows,c", [(37, 768), (4100, 768), (501, 256)])
def test_layernorm(ops, rows, c):
    from semivl_b200 import lib as L
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(rows, c, device="cuda", generator=g) * 2 + 0.5
    gamma = torch.randn(c, device="cuda", generator=g)
    beta = torch.randn(c, device="cuda", generator=g)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (c,), gr, br, 1e-6)
    y, mean, rstd = ops.layernorm_fwd(x, gamma, beta, 1e-6, out_dtype=L.F32)
    assert _rel(y, ref) < 1e-5
    yp, _, _ = ops.layernorm_fwd(x, gamma, beta, 1e-6, precise=True)
    assert _rel(_val(yp, True), ref) < 2e-5
    yb, _, _ = ops.layernorm_fwd(x, gamma, beta, 1e-6, precise=False)
    assert _rel(yb, ref) < 5e-3
    dy = torch.randn(rows, c, device="cuda", generator=g)
    r1 = torch.randn(rows, c, device="cuda", generator=g)
    ref.backward(dy)
    dgamma, dbeta = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    dx, dxa = ops.layernorm_bwd(dy, L.F32, x, gamma, mean, rstd, dres1=r1, act_precise=True, dgamma=dgamma, dbeta=dbeta)
    assert _rel(dx, xr.grad + r1) < 2e-5
    assert _rel(_val(dxa, True), xr.grad + r1) < 3e-5
    assert _rel(dgamma, gr.grad) < 1e-4 and _rel(dbeta, br.grad) < 1e-4
    # bf16 dy input, no weight grads
    dx2, _ = ops.layernorm_bwd(dy.to(torch.bfloat16), L.BF16, x, gamma, mean, rstd)
    xr.grad = None
    F.layer_norm(xr, (c,), gamma, beta, 1e-6).backward(dy.to(torch.bfloat16).float())
    assert _rel(dx2, xr.grad) < 2e-5
    # the encoder's common call (lean kernel for c = 768): bf16 / f32 dy, two residual gradients, bf16 operand copy, no weight grads
    r2 = torch.randn(rows, c, device="cuda", generator=g)
    for dyv, dt in ((dy.to(torch.bfloat16), L.BF16), (dy, L.F32)):
        dx3, dxa3 = ops.layernorm_bwd(dyv, dt, x, gamma, mean, rstd, dres1=r1, dres2=r2, act_precise=False)
        xr.grad = None
        F.layer_norm(xr, (c,), gamma, beta, 1e-6).backward(dyv.float())
        assert _rel(dx3, xr.grad + r1 + r2) < 2e-5
        assert _rel(dxa3, xr.grad + r1 + r2) < 5e-3
    _, dxa4 = ops.layernorm_bwd(dy, L.F32, x, gamma, mean, rstd, dres1=r1, want_dx=False, act_precise=False)
    xr.grad = None
    F.layer_norm(xr, (c,), gamma, beta, 1e-6).backward(dy)
    assert _rel(dxa4, xr.grad + r1) < 5e-3


@pytest.mark.parametrize("rows,cols,ld", [(16400, 2304, 2304), (5000, 32, 32), (3001, 48, 128), (777, 768, 768), (9, 8, 8), (2000, 200, 200)])
def test_colsum_shapes(ops, rows, cols, ld):
    """bias gradients: wide token matrices, 32-column conv gradients (all 256 threads busy), strided views, ragged row counts"""
    from semivl_b200 import lib as L
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(rows, ld, device="cuda", generator=g)
    for t, dt, tol in ((x, L.F32, 1e-5), (x.to(torch.bfloat16), L.BF16, 1e-5)):
        out = torch.full((cols,), 2.0, device="cuda")
        ops.colsum(t, dt, rows, cols, out, ld=ld)
        assert _rel(out, 2.0 + t[:, :cols].float().sum(0)) < tol


def test_l2norm_cast_colsum(ops):
    from semivl_b200 import lib as L
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(1000, 512, device="cuda", generator=g)
    xr = x.clone().requires_grad_(True)
    ref = F.normalize(xr, dim=1)
    y, ya, inv = ops.l2norm_fwd(x, act_precise=True)
    assert _rel(y, ref) < 1e-6 and _rel(_val(ya, True), ref) < 2e-5
    dy = torch.randn(1000, 512, device="cuda", generator=g)
    ref.backward(dy)
    dx = torch.zeros(1000, 512, device="cuda")
    ops.l2norm_bwd(dy, L.F32, y, inv, dx)
    assert _rel(dx, xr.grad) < 1e-5
    a = ops.to_act(x, True)
    assert _rel(_val(a, True), x) < 2e-5
    out = torch.ones(512, device="cuda")
    ops.colsum(a, L.BF16X2, 1000, 512, out)
    assert _rel(out, 1 + _val(a, True).sum(0)) < 1e-5
    acc = torch.ones(3, 7, device="cuda")
    ops.batch_sum(torch.arange(2 * 21, device="cuda", dtype=torch.float32).view(2, 3, 7), acc, accumulate=True)
    assert torch.equal(acc, 1 + torch.arange(21, device="cuda").view(3, 7) * 2.0 + 21)


def _attn_ref(qkv, b, L_, heads):
    E = heads * 64
    q, k, v = qkv.view(b, L_, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    o = s.softmax(-1) @ v
    return o.transpose(1, 2).reshape(b * L_, E), torch.logsumexp(s, -1)


@pytest.mark.parametrize("b,L_,heads", [(2, 197, 12), (1, 1025, 12), (3, 21, 4), (2, 64, 2), (1, 130, 1), (2, 1682, 3), (1, 2602, 2), (2, 128, 2),
                                        (1, 193, 2)])
@pytest.mark.parametrize("precise", [False, True])
def test_attention(ops, b, L_, heads, precise):
    from semivl_b200 import lib as L
    E = heads * 64
    g = torch.Generator(device="cuda").manual_seed(L_)
    qkv = torch.randn(b * L_, 3 * E, device="cuda", generator=g)
    dout = torch.randn(b * L_, E, device="cuda", generator=g)
    dvadd = torch.randn(b * L_, E, device="cuda", generator=g)
    if not precise:
        qkv, dout = qkv.to(torch.bfloat16).float(), dout.to(torch.bfloat16).float()
    qr = qkv.clone().requires_grad_(True)
    ref, ref_lse = _attn_ref(qr, b, L_, heads)
    ref.backward(dout)
    ref_d = qr.grad.clone()
    ref_d[:, 2 * E:] += dvadd
    a = ops.split_bf16(qkv) if precise else qkv.to(torch.bfloat16)
    out, lse = ops.attention_fwd(a, b, L_, heads, precise)
    tol = 3e-5 if precise else 1e-2
    assert _rel(_val(out, precise), ref) < tol
    assert _rel(lse, ref_lse) < (1e-5 if precise else 1e-3)
    da = ops.split_bf16(dout) if precise else dout.to(torch.bfloat16)
    dqkv = ops.attention_bwd(a, out, da, lse, b, L_, heads, precise, dv_add=dvadd, dv_add_dtype=L.F32)
    assert _rel(_val(dqkv, precise), ref_d) < (1e-4 if precise else 2e-2)


@pytest.mark.parametrize("L_", [1025, 300])
def test_attention_forward_rescales_when_the_running_maximum_grows(ops, L_):
    """The tcgen05 forward keeps O in TMEM and moves a row's reference maximum only when the running maximum exceeds it by more than 8 (log2
    units), rescaling O in place.  Keys whose scores GROW along the sequence (several crossings of the threshold per row, at different tiles
    for different rows) against the fp32 reference; also rows whose scores shrink (never rescale after the first tile)."""
    b, heads = 2, 3
    E = heads * 64
    g = torch.Generator(device="cuda").manual_seed(L_ + 5)
    qkv = torch.randn(b, L_, 3, heads, 64, device="cuda", generator=g)
    ramp = torch.linspace(0.0, 1.0, L_, device="cuda").view(1, L_, 1, 1)
    u = torch.randn(1, 1, heads, 64, device="cuda", generator=g)
    u = u / u.norm(dim=-1, keepdim=True)
    sign = torch.where(torch.arange(L_, device="cuda") % 3 == 0, -1.0, 1.0).view(1, L_, 1, 1)       # a third of the queries see shrinking scores
    qkv[:, :, 0] = qkv[:, :, 0] * 0.3 + 12.0 * u * sign              # queries aligned (or anti-aligned) with u
    qkv[:, :, 1] = qkv[:, :, 1] * 0.3 + 30.0 * u * ramp              # key component along u grows with the position: logits up to ~45
    qkv = qkv.reshape(b * L_, 3 * E).to(torch.bfloat16)
    ref, ref_lse = _attn_ref(qkv.float(), b, L_, heads)
    out, lse = ops.attention_fwd(qkv, b, L_, heads, False)
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
    assert _rel(out, ref) < 1e-2
    assert (lse - ref_lse).abs().max().item() < 2e-2              # absolute: lse reaches ~45 here (bf16 P rounding does not enter lse)


@pytest.mark.parametrize("g,oh,ow", [(4, 5, 5), (50, 51, 51), (40, 41, 41), (14, 32, 32), (32, 14, 20)])
def test_pos_resize_matches_torch_bicubic(ops, g, oh, ow):
    """svl_pos_resize_fwd / _bwd against F.interpolate(mode='bicubic', align_corners=False) and its autograd on the position grid
    (maskclip_vit.py:448-459,462-490): 50^2 -> 51^2 (801^2 crops), 40^2 -> 41^2 (641^2), the clip_encoder's 32^2 table on other crops."""
    from semivl_b200 import lib as L
    E = 768
    gen = torch.Generator(device="cuda").manual_seed(g * 100 + oh)
    pos = torch.randn(1, g * g + 1, E, device="cuda", generator=gen)
    out = torch.empty(oh * ow + 1, E, device="cuda")
    L.call("svl_pos_resize_fwd", pos, out, g, g, oh, ow, E)
    src = pos.clone().requires_grad_(True)
    grid = src[:, 1:].reshape(1, g, g, E).permute(0, 3, 1, 2)
    ref = torch.cat((src[:, :1], F.interpolate(grid, size=(oh, ow), mode="bicubic", align_corners=False).flatten(2).transpose(1, 2)), dim=1)[0]
    assert _rel(out, ref.detach()) < 2e-6
    dout = torch.randn(oh * ow + 1, E, device="cuda", generator=gen)
    ref.backward(dout)
    dpos = torch.ones(1, g * g + 1, E, device="cuda")                 # accumulates into the parameter gradient
    L.call("svl_pos_resize_bwd", dout, dpos, g, g, oh, ow, E)
    assert _rel(dpos - 1.0, src.grad) < 1e-5
