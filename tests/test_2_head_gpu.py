"""GPU parity of the VLG decode head engine (forward + backward) and of the fused upsample+CE loss against the CPU oracle.
precise mode: rel 2e-4 (logits) / 2e-2 (gradients); fast (bf16) mode: rel 4e-2 / 2e-1 (measured values are printed).
The gradient tolerances reflect the conditioning of the problem, not kernel error: perturbing the weights of the float64
oracle by 1e-5 relative (logits move 5e-5) already moves these gradients by 2-5e-2, because GroupNorm->ReLU sign flips are
discrete (scratch measurement recorded in DESIGN.md)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).abs().max() / (b.double().abs().max() + 1e-30)).item()


@pytest.mark.parametrize("hw,b,n", [(4, 2, 21), (5, 1, 19), (14, 1, 21)])
@pytest.mark.parametrize("precise", [True, False])
def test_head_forward_backward(text_dir, hw, b, n, precise):
    """Reference = the oracle evaluated in float64; the fp32 oracle's own distance to it is the noise floor of each gradient
    (ReLU/GroupNorm sign flips make some of them ill-conditioned with the random fixture weights).  The upstream gradient is a
    positive per-pixel weight on one class per pixel (coherent, like a real loss), not white noise."""
    from oracle import semivl_oracle as O
    from semivl_b200 import lib
    from semivl_b200.engine.head import HeadCfg, HeadEngine
    lib.check_device()
    mc = O.ModelCfg(img_size=hw * 16, num_classes=n)
    sd = O.fixture_state_dict(O.param_shapes(mc, with_clip_encoder=False), seed=0)
    g = torch.Generator().manual_seed(hw)
    f0 = [torch.randn(b, 768, hw, hw, generator=g), torch.randn(b, 768, hw, hw, generator=g)]
    e0 = torch.randn(b, 512, hw, hw, generator=g)
    f0.append(e0 / e0.norm(dim=1, keepdim=True))
    text = torch.from_numpy(np.load(f"{text_dir}/{'voc12_wbg_single' if n == 21 else 'cityscapes_single'}.npy"))
    lab_low = torch.randint(0, n, (b, 4 * hw, 4 * hw), generator=g)
    wl = F.one_hot(lab_low, n).permute(0, 3, 1, 2) * (0.5 + torch.rand(lab_low.shape, generator=g))[:, None]
    res = {}
    for dt in (torch.float64, torch.float32):
        pc = {k: v.clone().to(dt).requires_grad_(True) for k, v in sd.items() if k.startswith("decode_head.")}
        feats = [f.clone().to(dt).requires_grad_(True) for f in f0]
        low = O.vlg_head_forward(feats, text, pc, mc)
        (low * wl.to(dt)).sum().backward()
        res[dt] = (low.detach(), {k: v.grad for k, v in pc.items()}, [f.grad for f in feats])
    l64, g64, fg64 = res[torch.float64]
    l32, g32, fg32 = res[torch.float32]

    eng = HeadEngine(HeadCfg(), precise=precise)
    p = {k[len("decode_head."):]: v.cuda() for k, v in sd.items() if k.startswith("decode_head.")}
    fin = [f.permute(0, 2, 3, 1).contiguous().cuda() for f in f0]
    low, ctx = eng.forward(fin, text.cuda(), p, need_grad=True)
    tol_f, tol_g = (2e-4, 2e-2) if precise else (4e-2, 2e-1)
    r = _rel(low, l64)
    print(f"head hw {hw} precise {precise}: logits rel {r:.2e}")
    assert r < tol_f
    grads = {k: torch.zeros_like(v) for k, v in p.items()}
    dfe = eng.backward(ctx, wl.float().cuda(), p, grads)
    worst, bad = 0.0, []
    items = [(k, gv, g64["decode_head." + k], g32["decode_head." + k]) for k, gv in grads.items()]
    items += [(f"feat{i}", d.permute(0, 3, 1, 2), fg64[i], fg32[i]) for i, d in enumerate(dfe)]
    for k, mine, r64, r32 in items:
        if r64.abs().max() < 1e-12:
            continue
        rr, floor = _rel(mine, r64), _rel(r32, r64)
        worst = max(worst, rr)
        if rr >= max(tol_g, 8 * floor):
            bad.append((k, round(rr, 5), round(floor, 6)))
    print(f"head hw {hw} precise {precise}: worst grad rel {worst:.2e}")
    assert not bad, bad


@pytest.mark.parametrize("R,N,hl,H", [(2, 21, 16, 64), (1, 19, 20, 72), (1, 150, 32, 128), (1, 5, 41, 164), (2, 81, 9, 36)])
def test_fused_upsample_ce(R, N, hl, H):
    from semivl_b200 import lib as L
    L.check_device()
    g = torch.Generator().manual_seed(R + N)
    low = torch.randn(R, N, hl, hl, generator=g).cuda().requires_grad_(True)
    lab1 = torch.randint(0, N, (R, H, H), generator=g)
    lab1[:, : H // 5] = 255
    lab2 = torch.randint(0, N, (R, H, H), generator=g)
    w2 = (torch.rand(R, H, H, generator=g) > 0.4).float()
    lab1, lab2, w2 = lab1.cuda(), lab2.cuda(), w2.cuda()
    up = F.interpolate(low, size=(H, H), mode="bilinear", align_corners=False)
    l1 = F.cross_entropy(up, lab1, ignore_index=255)
    l2 = (F.cross_entropy(up, lab2, reduction="none") * w2).sum() / w2.numel()
    (l1 + 0.25 * l2).backward()
    # pseudo-labels
    conf_ref, arg_ref = up.detach().softmax(1).max(1)
    conf = torch.empty(R, H, H, device="cuda")
    arg = torch.empty(R, H, H, device="cuda", dtype=torch.int64)
    L.call("svl_softmax_max", low.detach(), conf, arg, R, N, hl, hl, H, H, 1.0, 0.0)
    assert (conf - conf_ref).abs().max() < 1e-5 and (arg == arg_ref).float().mean() > 0.9999
    outf = torch.empty(R, N, H, H, device="cuda")
    L.call("svl_upsample_bilinear", low.detach(), outf, R * N, hl, hl, H, H)
    assert (outf - up.detach()).abs().max() < 1e-5
    # fused loss
    cnt = torch.zeros(1, device="cuda")
    L.call("svl_count_valid", lab1, lab1.numel(), 255, cnt)
    c1 = torch.empty(1, device="cuda")
    L.call("svl_reciprocal", cnt, c1, 1.0, 1.0)
    c2 = torch.full((1,), 0.25 / w2.numel(), device="cuda")
    import ctypes as C
    labels = (C.c_void_p * 3)(lab1.data_ptr(), lab2.data_ptr(), None)
    weights = (C.c_void_p * 3)(None, w2.data_ptr(), None)
    coefs = (C.c_void_p * 3)(c1.data_ptr(), c2.data_ptr(), None)
    loss = torch.zeros(3, device="cuda")
    dlow = torch.zeros_like(low)
    L.call("svl_upsample_ce", low.detach(), dlow, R, N, hl, hl, H, H, 2, labels, weights, coefs, loss, 1.0, 255)
    assert abs(loss[0].item() - l1.item()) < 2e-5 * abs(l1.item()) + 1e-6
    assert abs(loss[1].item() - 0.25 * l2.item()) < 2e-5 * abs(l2.item()) + 1e-6
    assert _rel(dlow, low.grad) < 1e-4


@pytest.mark.parametrize("precise", [True, False])
def test_head_conv_feature_branch_matches_reference_golden(golden_dir, precise):
    """`skip_from_conv_feat=True` (the Cityscapes skr04 head, vlg_head.py:196-205) through the registered VLGHead module against outputs of
    the UNMODIFIED reference head (tests/golden/head_convfeat_b2.npz): the second skip is a 256-channel conv-encoder feature at its own
    16 x 16 resolution.  Logits, input gradients and parameter-gradient norms."""
    import os
    from oracle import semivl_oracle as O
    from oracle.make_golden import HEAD_CONV_KW
    from semivl_b200.model.vlg_head import VLGHead
    g = dict(np.load(os.path.join(golden_dir, "head_convfeat_b2.npz"), allow_pickle=False))
    head = VLGHead(precise=precise, **HEAD_CONV_KW)
    shapes = {"decode_head." + k: tuple(v.shape) for k, v in head.state_dict().items()}
    sd = O.fixture_state_dict(shapes, seed=3)
    head.load_state_dict({k[len("decode_head."):]: v for k, v in sd.items()})
    head = head.cuda()
    v4, emb, conv = (torch.from_numpy(g[k]).cuda().requires_grad_(True) for k in ("v4", "emb", "conv"))
    out = head([[[v4, emb], None], torch.from_numpy(g["text"]).cuda(), [conv]])
    ref = torch.from_numpy(g["out"])
    r = _rel(out.detach(), ref)
    print(f"conv-feature head precise {precise}: logits rel {r:.2e}")
    assert r < (2e-4 if precise else 4e-2)
    (out * torch.from_numpy(g["wgt"]).cuda()).sum().backward()
    for name, t in (("d_v4", v4), ("d_emb", emb), ("d_conv", conv)):
        want = torch.from_numpy(g[name])
        if precise:
            assert _rel(t.grad, want) < 2e-2, name
        else:
            # bf16 operands: individual entries of these gradients are ill-conditioned with the fixture weights (GroupNorm -> ReLU sign
            # flips, see the module docstring); the direction and size of the gradient are what the throughput mode has to keep
            a, b_ = t.grad.double().cpu().flatten(), want.double().flatten()
            cos = (a @ b_ / (a.norm() * b_.norm())).item()
            assert cos > 0.9 and abs(a.norm().item() / b_.norm().item() - 1) < 0.2, (name, cos)
    if precise:
        grads = dict(head.named_parameters())
        for n, norm in zip(g["grad_names"], g["grad_norms"]):
            n = str(n)[len("decode_head."):]
            if norm > 1e-7:
                assert abs(grads[n].grad.double().norm().item() - norm) <= 3e-2 * norm, n
