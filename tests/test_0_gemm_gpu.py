"""GPU parity of the tcgen05 contraction engine (svl_gemm) against plain PyTorch fp32 on the same operands.

fast mode   : operands rounded to bf16, fp32 accumulate -> compared with an fp32 matmul of the rounded operands (tol 2e-3 of scale)
precise mode: split bf16 pairs, 3 taps                  -> compared with an fp32 matmul of the ORIGINAL operands (tol 3e-5 of scale)
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False          # the fp32 references must be fp32
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="module")
def ops():
    from semivl_b200 import lib, ops
    lib.check_device()
    return ops


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12)).item()


def _bf(x):
    return x.to(torch.bfloat16)


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (300, 200, 192), (1000, 768, 768), (4100, 2304, 768), (77, 21, 512), (513, 3072, 768)])
def test_plain_gemm(ops, m, n, k):
    from semivl_b200 import lib as L
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) / k ** 0.5
    bias = torch.randn(n, device="cuda", generator=g)
    out = torch.full((m, n), float("nan"), device="cuda")
    ops.gemm(_bf(a), _bf(w), out, n=n, k=k, bias=bias)
    ref = _bf(a).float() @ _bf(w).float().t() + bias
    assert _rel(out, ref) < 1e-4
    # bf16 output
    out16 = torch.zeros(m, n, device="cuda", dtype=torch.bfloat16)
    ops.gemm(_bf(a), _bf(w), out16, n=n, k=k, bias=bias)
    assert _rel(out16, ref) < 6e-3
    # precise mode (split operands)
    outp = torch.zeros(m, n, device="cuda")
    ops.gemm(ops.split_bf16(a), ops.split_bf16(w), outp, n=n, k=k, bias=bias, precise=True)
    refp = a @ w.t() + bias
    assert _rel(outp, refp) < 3e-5
    # split output of precise mode
    outs = torch.zeros(m, 2 * n, device="cuda", dtype=torch.bfloat16)
    ops.gemm(ops.split_bf16(a), ops.split_bf16(w), outs, n=n, k=k, bias=bias, precise=True, out_dtype=L.BF16X2)
    assert _rel(outs[:, :n].float() + outs[:, n:].float(), refp) < 3e-5


def test_epilogue_gelu_residual_preact(ops):
    from semivl_b200 import lib as L
    m, n, k = 700, 320, 256
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) / k ** 0.5
    bias = torch.randn(n, device="cuda", generator=g)
    res = torch.randn(m, n, device="cuda", generator=g)
    out = torch.zeros(m, n, device="cuda")
    pre = torch.zeros(m, n, device="cuda")
    ops.gemm(_bf(a), _bf(w), out, n=n, k=k, bias=bias, act=L.ACT_GELU, residual=res, preact_out=pre, alpha=0.5)
    z = 0.5 * (_bf(a).float() @ _bf(w).float().t()) + bias
    assert _rel(pre, z) < 1e-4
    assert _rel(out, F.gelu(z) + res) < 1e-4
    # act' multiply (GELU backward) + accumulate
    src = torch.randn(m, n, device="cuda", generator=g)
    acc = torch.randn(m, n, device="cuda", generator=g)
    acc0 = acc.clone()
    ops.gemm(_bf(a), _bf(w), acc, n=n, k=k, dact_src=src, dact_kind=L.ACT_GELU, accumulate=True)
    s = src.clone().requires_grad_(True)
    F.gelu(s).sum().backward()
    ref = acc0 + (_bf(a).float() @ _bf(w).float().t()) * s.grad
    assert _rel(acc, ref) < 1e-4
    # relu' from saved output + row bias
    y = torch.randn(m, n, device="cuda", generator=g)
    rb = torch.randn(7, n, device="cuda", generator=g)
    o2 = torch.zeros(m, n, device="cuda")
    ops.gemm(_bf(a), _bf(w), o2, n=n, k=k, dact_src=_bf(y), dact_kind=L.ACT_RELU, row_bias=rb, row_bias_div=100)
    zz = _bf(a).float() @ _bf(w).float().t() + rb[torch.arange(m, device="cuda") // 100]
    assert _rel(o2, zz * (_bf(y).float() > 0)) < 1e-4


@pytest.mark.parametrize("m,n,k", [(700, 320, 256), (2100, 3072, 768), (130, 96, 64)])
def test_ffn_epilogues_bf16(ops, m, n, k):
    """The specialised throughput-mode epilogues of the transformer FFN: GELU forward saving z or gelu'(z), and the data-gradient
    epilogue multiplying by gelu'(z) (recomputed from z, or read back as saved), plus the fp32 residual epilogue."""
    from semivl_b200 import lib as L
    g = torch.Generator(device="cuda").manual_seed(m + n)
    a = _bf(torch.randn(m, k, device="cuda", generator=g))
    w = _bf(torch.randn(n, k, device="cuda", generator=g) / k ** 0.5)
    bias = torch.randn(n, device="cuda", generator=g)
    z = a.float() @ w.float().t() + bias
    zr = z.clone().requires_grad_(True)
    F.gelu(zr).sum().backward()
    out = torch.zeros(m, n, device="cuda", dtype=torch.bfloat16)
    pre = torch.zeros(m, n, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, out, n=n, k=k, bias=bias, act=L.ACT_GELU, preact_out=pre)
    assert _rel(out, F.gelu(z)) < 6e-3 and _rel(pre, z) < 6e-3
    out2 = torch.zeros_like(out)
    dsave = torch.zeros_like(pre)
    ops.gemm(a, w, out2, n=n, k=k, bias=bias, act=L.ACT_GELU_DSAVE, preact_out=dsave)
    assert torch.equal(out2, out)
    assert (dsave.float() - zr.grad).abs().max().item() < 6e-3
    # backward: dY [m, n2] x W2^T -> [m, n], times gelu'
    n2 = 128
    dy = _bf(torch.randn(m, n2, device="cuda", generator=g))
    w2t = _bf(torch.randn(n, n2, device="cuda", generator=g) / n2 ** 0.5)
    lin = dy.float() @ w2t.float().t()
    d1 = torch.zeros(m, n, device="cuda", dtype=torch.bfloat16)
    ops.gemm(dy, w2t, d1, n=n, k=n2, dact_src=pre, dact_kind=L.ACT_GELU)
    s = pre.float().requires_grad_(True)
    F.gelu(s).sum().backward()
    assert _rel(d1, lin * s.grad) < 6e-3
    d2 = torch.zeros_like(d1)
    ops.gemm(dy, w2t, d2, n=n, k=n2, dact_src=dsave, dact_kind=L.ACT_SAVED)
    assert _rel(d2, lin * dsave.float()) < 6e-3
    assert _rel(d2, lin * zr.grad) < 1.5e-2
    # fp32 residual epilogue
    res = torch.randn(m, n, device="cuda", generator=g)
    o3 = torch.zeros(m, n, device="cuda")
    ops.gemm(a, w, o3, n=n, k=k, bias=bias, residual=res)
    assert _rel(o3, z + res) < 1e-4


@pytest.mark.parametrize("nb,h,w,cin,cout,ks,dil", [(5, 12, 12, 64, 48, 3, 2), (7, 32, 32, 128, 128, 3, 6), (3, 64, 64, 128, 64, 3, 1),
                                                     (2, 128, 128, 32, 32, 3, 1), (2, 5, 5, 128, 128, 3, 1), (3, 51, 51, 64, 16, 3, 1),
                                                     (2, 128, 128, 32, 1, 3, 1), (4, 8, 8, 256, 64, 1, 1), (3, 37, 128, 64, 32, 3, 1),
                                                     (2, 21, 164, 32, 64, 3, 1), (5, 1, 130, 32, 32, 3, 1), (150, 16, 128, 32, 32, 3, 1),
                                                     (2, 40, 204, 64, 64, 3, 1), (5, 64, 64, 64, 64, 3, 1), (3, 40, 48, 32, 32, 3, 1),
                                                     (2, 20, 130, 128, 32, 3, 1), (7, 33, 64, 128, 32, 3, 1)])
@pytest.mark.parametrize("precise", [False, True])
def test_implicit_conv(ops, nb, h, w, cin, cout, ks, dil, precise):
    g = torch.Generator(device="cuda").manual_seed(nb * h + cin + cout)
    x = torch.randn(nb, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, ks, ks, device="cuda", generator=g) / (cin * ks * ks) ** 0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    if not precise:
        x, wt = _bf(x).float(), _bf(wt).float()
    ref = F.conv2d(x, wt, bias, padding=dil * (ks // 2), dilation=dil).permute(0, 2, 3, 1)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    a = ops.split_bf16(x_nhwc) if precise else _bf(x_nhwc)
    wk = wt.permute(2, 3, 0, 1).reshape(ks * ks * cout, cin)          # [tap*cout, cin]
    b = ops.prep_weight(wk, precise)
    filt = [((i - ks // 2) * dil, (j - ks // 2) * dil) for i in range(ks) for j in range(ks)]
    out = torch.full((nb, h, w, cout), float("nan"), device="cuda")
    ops.gemm(a, b, out, n=cout, k=cin, precise=precise, conv=(nb, h, w), filt=filt, b_row_stride=cout, bias=bias)
    assert _rel(out, ref) < (3e-5 if precise else 1e-4)
    if not precise:
        # bf16 output (+ ReLU): the 3 x 3 / 32-64 channel / >= 96-pixel-wide cases take the rolling-accumulator kernel (conv_roll.cu)
        from semivl_b200 import lib as L
        out16 = torch.full((nb, h, w, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.gemm(a, b, out16, n=cout, k=cin, conv=(nb, h, w), filt=filt, b_row_stride=cout, bias=bias)
        assert _rel(out16, ref) < 6e-3
        ops.gemm(a, b, out16, n=cout, k=cin, conv=(nb, h, w), filt=filt, b_row_stride=cout, act=L.ACT_RELU)
        assert _rel(out16, (ref - bias).clamp_min(0)) < 6e-3
        if cout in (32, 64, 128):
            # GroupNorm statistics fused into the epilogue (conv_roll only: `ws` appears in the dict when the library takes that path):
            # same mean / rstd / normalised output as the stand-alone statistics pass over the stored bf16 tensor
            G = cout // 16
            gs = dict(maps=nb, G=G)
            ops.gemm(a, b, out16, n=cout, k=cin, conv=(nb, h, w), filt=filt, b_row_stride=cout, bias=bias, gn_stats=gs)
            gamma, beta = torch.rand(cout, device="cuda") + 0.5, torch.randn(cout, device="cuda")
            raw = out16.view(-1, cout)
            y0, y1 = torch.empty_like(raw), torch.empty_like(raw)
            m0, r0 = ops.gn_relu_fwd(raw, L.BF16, gamma, beta, y0, L.BF16, nb, h * w, cout, G)
            m1, r1 = ops.gn_relu_fwd(raw, L.BF16, gamma, beta, y1, L.BF16, nb, h * w, cout, G, stats=gs)
            assert ("ws" in gs) == (ks == 3 and dil == 1 and cout in (32, 64) and cin in (32, 64, 128) and (w <= 64 and w >= 33 or w >= 96))
            assert (m0 - m1).abs().max() < 1e-5 * (1 + m0.abs().max()) and ((r0 - r1).abs() / r0).max() < 1e-4
            assert (y0.float() - y1.float()).abs().max() <= 2e-2 * y0.float().abs().max()


def test_convtranspose_scatter(ops):
    from semivl_b200 import lib as L
    nb, h, w, cin, cout = 3, 16, 16, 128, 96
    g = torch.Generator(device="cuda").manual_seed(5)
    x = _bf(torch.randn(nb, cin, h, w, device="cuda", generator=g)).float()
    wt = _bf(torch.randn(cin, cout, 2, 2, device="cuda", generator=g) / cin ** 0.5).float()
    bias = torch.randn(cout, device="cuda", generator=g)
    ref = F.conv_transpose2d(x, wt, bias, stride=2).permute(0, 2, 3, 1)
    a = _bf(x.permute(0, 2, 3, 1).contiguous()).reshape(nb * h * w, cin)
    b = _bf(wt.permute(2, 3, 1, 0).reshape(4 * cout, cin).contiguous())          # rows (qy, qx, co)
    ldc = 128
    out = torch.zeros(nb, 2 * h, 2 * w, ldc, device="cuda")
    ops.gemm(a, b, out, n=4 * cout, k=cin, bias=bias.repeat(4), out_mode=L.OUT_CONVT2X2, out_hw=(h, w), ldc=ldc)
    assert _rel(out[..., :cout], ref) < 1e-4
    assert out[..., cout:].abs().max().item() == 0


@pytest.mark.parametrize("rows,m,n", [(64, 128, 64), (1000, 200, 100), (4100, 2304, 768), (16400, 768, 768), (333, 21, 512)])
@pytest.mark.parametrize("precise", [False, True])
def test_wgrad_linear(ops, rows, m, n, precise):
    g = torch.Generator(device="cuda").manual_seed(rows + m)
    dy = torch.randn(rows, m, device="cuda", generator=g)
    x = torch.randn(rows, n, device="cuda", generator=g)
    if not precise:
        dy, x = _bf(dy).float(), _bf(x).float()
    mp, np_ = (m + 7) // 8 * 8, (n + 7) // 8 * 8           # TMA needs 16-byte row strides
    dyp = torch.zeros(rows, mp, device="cuda"); dyp[:, :m] = dy
    xp = torch.zeros(rows, np_, device="cuda"); xp[:, :n] = x
    a = ops.split_bf16(dyp) if precise else _bf(dyp)
    b = ops.split_bf16(xp) if precise else _bf(xp)
    dw = torch.ones(m, n, device="cuda")
    ops.wgrad(a, b, dw, m=m, n=n, precise=precise, alpha=0.5)
    ref = 1.0 + 0.5 * (dy.double().t() @ x.double()).float()
    assert _rel(dw, ref) < (3e-5 if precise else 1e-4)


@pytest.mark.parametrize("nb,h,w,cin,cout,ks,dil", [(5, 12, 12, 64, 48, 3, 2), (7, 32, 32, 128, 128, 3, 6), (3, 64, 64, 128, 64, 3, 1),
                                                     (2, 128, 128, 32, 32, 3, 1), (2, 5, 5, 128, 128, 3, 1), (3, 51, 51, 64, 16, 3, 1),
                                                     (4, 8, 8, 256, 64, 1, 1), (2, 128, 128, 64, 32, 3, 1), (3, 64, 64, 96, 64, 3, 1),
                                                     (2, 32, 32, 256, 128, 3, 1), (2, 164, 164, 32, 32, 3, 1), (3, 82, 82, 64, 64, 3, 1),
                                                     (2, 204, 204, 64, 32, 3, 1), (2, 100, 70, 128, 64, 3, 1), (3, 67, 130, 32, 32, 3, 1), (1, 65, 64, 64, 32, 3, 1)])
@pytest.mark.parametrize("precise", [False, True])
def test_wgrad_conv(ops, nb, h, w, cin, cout, ks, dil, precise):
    g = torch.Generator(device="cuda").manual_seed(nb * h + cin + cout)
    x = torch.randn(nb, cin, h, w, device="cuda", generator=g)
    dy = torch.randn(nb, cout, h, w, device="cuda", generator=g)
    if not precise:
        x, dy = _bf(x).float(), _bf(dy).float()
    wt = torch.zeros(cout, cin, ks, ks, device="cuda", dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), wt, None, padding=dil * (ks // 2), dilation=dil).backward(dy.double())
    ref = wt.grad.permute(2, 3, 0, 1).reshape(ks * ks, cout, cin).float()
    xa = x.permute(0, 2, 3, 1).contiguous()
    dya = dy.permute(0, 2, 3, 1).contiguous()
    a = ops.split_bf16(dya) if precise else _bf(dya)
    b = ops.split_bf16(xa) if precise else _bf(xa)
    filt = [((i - ks // 2) * dil, (j - ks // 2) * dil) for i in range(ks) for j in range(ks)]
    dw = torch.zeros(ks * ks, cout, cin, device="cuda")
    ops.wgrad(a, b, dw, m=cout, n=cin, precise=precise, conv=(nb, h, w), filt=filt)
    assert _rel(dw, ref) < (5e-5 if precise else 1e-4)          # 2 x 204 x 204 rows summed: 3.04e-5 measured in the split-operand mode


def test_throughput_report(ops):
    """Not an assertion on speed: prints achieved TFLOP/s for the ViT-block shapes (used when reading gpurun logs)."""
    for (m, n, k) in [(16400, 2304, 768), (16400, 768, 768), (16400, 3072, 768), (16400, 768, 3072)]:
        a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
        w = torch.randn(n, k, device="cuda").to(torch.bfloat16)
        out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm(a, w, out, n=n, k=k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm(a, w, out, n=n, k=k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"gemm {m}x{n}x{k}: {ms * 1e3:.1f} us  {2 * m * n * k / ms / 1e9:.1f} TFLOP/s")
    for (rows, m, n) in [(16400, 2304, 768), (16400, 768, 768)]:
        dy = torch.randn(rows, m, device="cuda").to(torch.bfloat16)
        x = torch.randn(rows, n, device="cuda").to(torch.bfloat16)
        dw = torch.zeros(m, n, device="cuda")
        for _ in range(3):
            ops.wgrad(dy, x, dw, m=m, n=n)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.wgrad(dy, x, dw, m=m, n=n)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"wgrad {rows}x{m}x{n}: {ms * 1e3:.1f} us  {2 * m * n * rows / ms / 1e9:.1f} TFLOP/s")


@pytest.mark.parametrize("env", [dict(SVL_GEMM_CLUSTER="1"), dict(SVL_GEMM_CLUSTER="2"), dict(SVL_GEMM_STAGED="0"), dict(SVL_GEMM_STAGED="15"),
                                 dict(SVL_GEMM_CLUSTER="2", SVL_GEMM_STAGED="15"), dict(SVL_GEMM_CLUSTER="0", SVL_GEMM_STAGED="15")])
def test_optin_launch_modes_subprocess(env):
    """Launch modes of the contraction engine, read once per process, exercised in a child process on the plain / epilogue GEMM tests:
    SVL_GEMM_CLUSTER=1 (2-CTA clusters, B tile TMA-multicast, 2-arrival stage barriers), =2 (CTA pairs running one cta_group::2 UMMA with
    M = 256, half of B per CTA, leader-side barriers), =0 (never); SVL_GEMM_STAGED = bit mask of the epilogues that go through the per-quartet
    shared-memory buffers and TMA (0: none, the register / direct-store epilogues; 15: all, including every plain bf16 output)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "-x", "-k",
                        "plain_gemm or epilogue or ffn_epilogues or wgrad_linear"], env=dict(os.environ, **env), cwd=root, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("nb,h,w,cin,cout", [(1, 1, 33, 32, 32), (3, 2, 47, 64, 32), (2, 17, 64, 128, 64), (1, 33, 96, 32, 64), (3, 5, 127, 64, 64),
                                             (1, 18, 129, 32, 32), (2, 3, 200, 128, 32), (1, 16, 257, 64, 32), (5, 31, 40, 32, 32),
                                             (4, 16, 64, 64, 64)])
def test_conv_roll_geometry_sweep(ops, nb, h, w, cin, cout):
    """The rolling-accumulator convolution kernel (conv_roll.cu) on awkward geometries: single rows, odd map counts in the dual-map form, rows
    that end inside a 128-pixel tile, units shorter than the ring, every channel combination -- output and fused GroupNorm statistics against
    fp32 PyTorch on the bf16-rounded operands."""
    from semivl_b200 import lib as L
    g = torch.Generator(device="cuda").manual_seed(nb * 1000 + h * 10 + w + cin + cout)
    x = _bf(torch.randn(nb, cin, h, w, device="cuda", generator=g)).float()
    wt = _bf(torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (9 * cin) ** 0.5).float()
    ref = F.conv2d(x, wt, None, padding=1).permute(0, 2, 3, 1)
    a = _bf(x.permute(0, 2, 3, 1).contiguous())
    b = _bf(wt.permute(2, 3, 0, 1).reshape(9 * cout, cin))
    filt = [(i - 1, j - 1) for i in range(3) for j in range(3)]
    G = cout // 16
    gs = dict(maps=nb, G=G)
    out = torch.full((nb, h, w, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, out, n=cout, k=cin, conv=(nb, h, w), filt=filt, b_row_stride=cout, gn_stats=gs)
    assert "ws" in gs, "this geometry should take the rolling kernel"
    assert _rel(out, ref) < 6e-3
    # the mirrored taps of the data gradient (weights [tap, ci, co]) through the same kernel
    dy = _bf(torch.randn(nb, cout, h, w, device="cuda", generator=g)).float()
    refd = F.conv_transpose2d(dy, wt, None, padding=1).permute(0, 2, 3, 1)
    bt = _bf(wt.permute(2, 3, 1, 0).reshape(9 * cin, cout))
    if cin in (32, 64):
        dx = torch.full((nb, h, w, cin), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.gemm(_bf(dy.permute(0, 2, 3, 1).contiguous()), bt, dx, n=cin, k=cout, conv=(nb, h, w), filt=[(-fy, -fx) for fy, fx in filt], b_row_stride=cin)
        assert _rel(dx, refd) < 6e-3
    # fused statistics == statistics of the stored tensor
    stored = out.float().view(nb, h * w, G, 16)
    mean_ref = stored.mean(dim=(1, 3))
    var_ref = stored.var(dim=(1, 3), unbiased=False)
    gamma, beta = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
    y = torch.empty(nb * h * w, cout, device="cuda", dtype=torch.bfloat16)
    m1, r1 = ops.gn_relu_fwd(out.view(-1, cout), L.BF16, gamma, beta, y, L.BF16, nb, h * w, cout, G, stats=gs)
    assert (m1 - mean_ref).abs().max() < 1e-5 * (1 + mean_ref.abs().max())
    assert ((r1 - (var_ref + 1e-5).rsqrt()).abs() / r1).max() < 1e-3
