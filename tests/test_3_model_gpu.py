"""End-to-end GPU parity through the reference-facing API (build_model -> model(img) / forward_maskclip -> CE -> backward)
against the golden vectors recorded from the UNMODIFIED reference (tests/golden, oracle/make_golden.py) and the fused
training steps against the oracle.  Tolerances: north-star 1e-3 relative on logits in precise mode (measured ~3e-5);
bf16 throughput mode is measured and bounded at 4e-2."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _cfg(crop, nclass=21, precise=True, mcc="single"):
    return dict(model='mmseg.vlm-vlg-aspp-s2p4-sk04-ftap-mcvitb', nclass=nclass, crop_size=crop, dataset='pascal',
                text_embedding_variant='single', mcc_text=mcc, pl_text='single', clip_encoder='mcvit16', disable_dropout=True, fp_rate=0.5,
                model_args=dict(pretrained=None), clip_encoder_args=dict(pretrained=None), precise=precise)


def _build(crop, precise, nclass=21):
    from oracle import semivl_oracle as O
    from semivl_b200.model import build_model
    m = build_model(_cfg(crop, nclass, precise))
    mc = O.ModelCfg(img_size=crop, num_classes=nclass)
    sd = O.fixture_state_dict(O.param_shapes(mc), seed=0)
    m.load_state_dict(sd)
    return m.cuda(), mc, sd


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


@pytest.mark.parametrize("name", ["fwd_c64_b2", "fwd_c72_b1", "fwd_c224_b1"])
@pytest.mark.parametrize("precise", [True, False])
def test_forward_matches_reference_golden(golden_dir, name, precise):
    from oracle.make_golden import sample_idx
    g = dict(np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False))
    crop = int(g["crop"])
    m, mc, sd = _build(crop, precise, int(g["nclass"]))
    img = torch.from_numpy(g["img"]).cuda()
    lab = torch.from_numpy(g["label"].astype(np.int64)).cuda()
    m.train()
    feats, glob = m.backbone(img)
    tol = 1e-3 if precise else 4e-2
    assert _rel(feats[-1], g["emb"]) < tol and _rel(glob, g["global_emb"]) < tol
    y = m(img)
    assert y.shape == (img.shape[0], int(g["nclass"]), crop, crop)
    ys = y.detach().flatten().cpu()[sample_idx(y.numel(), 8192)]
    r = ((ys - torch.from_numpy(g["logits_sample"])).abs().max() / np.abs(g["logits_lowres"]).max()).item()
    print(f"{name} precise={precise}: logits rel {r:.2e}")
    assert r < tol
    # argmax masks: bit-exact wherever the reference's top-1/top-2 margin exceeds twice the measured max logit error
    # (random-init fixture weights give near-tied classes on part of the pixels)
    ref_full = F.interpolate(torch.from_numpy(g["logits_lowres"]), size=(crop, crop), mode="bilinear", align_corners=False)
    err = (y.detach().cpu() - ref_full).abs().max().item()
    top2 = ref_full.topk(2, dim=1).values
    decidable = (top2[:, 0] - top2[:, 1]) > 2 * err
    same = y.argmax(1).cpu() == torch.from_numpy(g["argmax"].astype(np.int64))
    agree = same.float().mean().item()
    print(f"{name} precise={precise}: argmax agreement {agree:.5f} overall, decidable pixels {decidable.float().mean().item():.3f}, max |err| {err:.2e}")
    assert bool(same[decidable].all())
    assert agree > (0.9995 if precise else 0.80)
    loss = F.cross_entropy(y, lab, ignore_index=255)
    assert abs(loss.item() - float(g["loss"])) < (1e-4 if precise else 2e-2) * float(g["loss"])
    loss.backward()
    if precise:
        named = dict(m.named_parameters())
        for nme, norm in zip(g["grad_names"], g["grad_norms"]):
            gr = named[str(nme)].grad
            assert gr is not None, nme
            if norm > 1e-7:
                assert abs(gr.double().norm().item() - norm) <= 3e-2 * norm, (nme, gr.norm().item(), norm)
        worst = _check_grad_samples(_sampled_grads({n: q.grad for n, q in named.items()}, g["grad_names"]), g["grad_samples"], g["grad_norms"],
                                    0.995, name)
        print(f"{name}: worst per-tensor cosine of the 64 sampled gradient entries {worst:.6f}")
    for key, th in (("maskclip", 0.9), ("maskclip_lo", float(g["maskclip_lo_thresh"]))):
        mcl = m.forward_maskclip(img, th)
        assert (mcl.cpu().numpy().astype(np.uint8) == g[key]).mean() > (0.999 if precise else 0.97)


@pytest.mark.parametrize("precise", [True, False])
def test_supervised_step_matches_oracle(text_dir, precise):
    """Fused step (engines + fused upsample/CE + flat AdamW) vs oracle loss/grads and torch.optim.AdamW semantics."""
    from oracle import semivl_oracle as O
    from semivl_b200.train import OptimCfg, Trainer
    crop, b = 64, 2
    m, mc, sd = _build(crop, precise)
    g = torch.Generator().manual_seed(3)
    img = torch.randn(b, 3, crop, crop, generator=g)
    mask = torch.randint(0, 21, (b, crop, crop), generator=g)
    mask[:, :9, :20] = 255
    text = torch.from_numpy(np.load(os.path.join(text_dir, "voc12_wbg_single.npy")))
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss_ref = O.supervised_step_loss(img, mask, p, text, mc)
    loss_ref.backward()
    tr = Trainer(m, OptimCfg(lr=1e-4, total_iters=100))
    before = tr.p_flat.clone()
    loss = tr.supervised_step(img.cuda(), mask.cuda(), update=True)
    assert abs(loss.item() - loss_ref.item()) < (1e-4 if precise else 2e-2) * loss_ref.item()
    if precise:
        for prefix, gd in (("backbone.", tr.g_bb), ("decode_head.", tr.g_hd)):
            for k, gv in gd.items():
                gr = p[prefix + k].grad
                nr = gr.double().norm().item()
                if nr > 1e-7:
                    assert abs(gv.double().norm().item() - nr) <= 3e-2 * nr, (k, gv.norm().item(), nr)
    # first AdamW step: |delta| = lr_group * (1 + wd*|p|-ish): sign(g) * lr  (m_hat / sqrt(v_hat) = sign(g) at step 1)
    delta = (tr.p_flat - before)
    gs = tr.g_flat
    nz = gs.abs() > 1e-6
    lr = torch.cat((torch.full((tr.n_bb,), 1e-4 * 0.01), torch.full((tr.n_hd,), 1e-4 * 10.0))).cuda()
    expect = -lr * gs / (gs.abs() + 1e-8) - lr * 0.01 * before           # step 1: m_hat / (sqrt(v_hat) + eps) = g / (|g| + eps)
    assert ((delta - expect).abs()[nz] <= 1e-3 * lr[nz] + 2e-8).all()          # fp32 ulp of the parameters is ~2e-9


@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("name", ["step_c64_b1", "step_c96_b2", "step_c64_b2_pixelavg_mv", "step_c64_b2_pixelratio_mean"])
def test_semivl_step_matches_reference_golden(golden_dir, name, precise):
    """The fused SemiVL step (one 4b encoder pass, 5b head pass, teacher + MaskCLIP passes, fused losses) against the loss terms
    and gradient norms recorded from the UNMODIFIED reference driven in the order of semivl.py:224-323.  Precise mode is the parity mode
    (bounds 3e-3 / 1e-3 / 5e-2); the bf16 throughput mode -- the one bench.py times -- is held to stated, looser bounds: its pseudo-labels are
    arg-maxes of near-tied logits computed with bf16 operands, so a share of them flips and moves every consistency term."""
    T_TERMS, T_LOSS, T_NORM, T_COS = (3e-3, 1e-3, 5e-2, 0.99) if precise else (5e-2, 3e-2, 0.35, 0.80)
    from semivl_b200.train import OptimCfg, Trainer
    g = dict(np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False))
    crop, b = int(g["crop"]), int(g["b"])
    m, mc, sd = _build(crop, precise, int(g["nclass"]))
    lk = ("mask_x", "ignore_mask", "ignore_mask_other")
    batch = {k: torch.from_numpy(g[k].astype(np.int64) if k in lk else g[k]).cuda()
             for k in ("img_x", "img_w", "img_s1", "img_s2", "img_w_other", "img_s1_other", "img_s2_other",
                       "mask_x", "ignore_mask", "ignore_mask_other", "mix1", "mix2")}
    hp = dict(conf_thresh=float(g["hp_conf_thresh"]), conf_mode=str(g["hp_conf_mode"]), mcc_conf_thresh=float(g["hp_mcc_conf_thresh"]),
              mcc_loss_reduce=str(g["hp_mcc_loss_reduce"]), mcc_lambda=float(g["hp_mcc_lambda"]))
    masks = [torch.from_numpy(g[f"drop_mask{i}"])[b:, :, 0, 0].contiguous().cuda() for i in range(3)]
    tr = Trainer(m, OptimCfg(), hp=hp)
    m.train()
    loss, terms = tr.semivl_step(batch, drop_masks=masks, update=False)
    got = np.array([terms[k].item() for k in ("loss_x", "loss_s1", "loss_s2", "loss_fp", "loss_mc_s1", "loss_mc_s2", "loss_mc_fp")])
    print(name, "terms", got, "ref", g["terms"])
    # pseudo-labels are argmaxes / thresholded confidences of near-degenerate logits: a handful may flip -> 3e-3 rel on the terms
    print(f"{name} precise={precise}: terms rel {np.abs(got - g['terms']).max() / np.abs(g['terms']).max():.3e}, "
          f"loss rel {abs(loss.item() - float(g['loss'])) / float(g['loss']):.3e}")
    assert np.abs(got - g["terms"]).max() < T_TERMS * np.abs(g["terms"]).max()
    assert abs(loss.item() - float(g["loss"])) < T_LOSS * float(g["loss"])
    worst_norm = 0.0
    for prefix, gd in (("backbone.", tr.g_bb), ("decode_head.", tr.g_hd)):
        for nme, norm in zip(g["grad_names"], g["grad_norms"]):
            nme = str(nme)
            if nme.startswith(prefix) and norm > 1e-7:
                gv = gd[nme[len(prefix):]]
                worst_norm = max(worst_norm, abs(gv.double().norm().item() - norm) / norm)
                assert abs(gv.double().norm().item() - norm) <= T_NORM * norm, (nme, gv.norm().item(), norm)
    named = {"backbone." + k: v for k, v in tr.g_bb.items()}
    named.update({"decode_head." + k: v for k, v in tr.g_hd.items()})
    names = [n for n in g["grad_names"] if str(n) in named]
    keep = [i for i, n in enumerate(g["grad_names"]) if str(n) in named]
    worst = _check_grad_samples(_sampled_grads(named, names), g["grad_samples"][keep], g["grad_norms"][keep], T_COS, name)
    print(f"{name} precise={precise}: worst gradient-norm deviation {worst_norm:.3e}, worst per-tensor cosine of the 64 sampled gradient entries {worst:.6f}")


def test_graph_replay_matches_eager_steps(text_dir):
    """`graphed_supervised_step` (one CUDA-graph launch per step, AdamW scalars from device memory) against the eager kernel
    sequence over 4 optimizer steps with changing inputs.  The fp32 atomics of the statistics / weight-gradient reductions make
    two EAGER runs differ too (and the random-init head amplifies it), so the bound is calibrated on an eager-vs-eager pair:
    graph-vs-eager must sit within 3x that noise floor (precise mode, where the floor is small)."""
    from semivl_b200.train import OptimCfg, Trainer
    crop, b = 64, 2
    g = torch.Generator().manual_seed(5)
    imgs = [torch.randn(b, 3, crop, crop, generator=g).cuda() for _ in range(4)]
    masks = [torch.randint(0, 21, (b, crop, crop), generator=g).cuda() for _ in range(4)]
    runs = []
    for graphed in (False, False, True):
        m, mc, sd = _build(crop, True)
        tr = Trainer(m, OptimCfg(lr=1e-4, total_iters=10))
        losses = []
        for img, mask in zip(imgs, masks):
            out = tr.graphed_supervised_step(img, mask) if graphed else tr.supervised_step(img, mask)
            losses.append(out.item())
        assert tr.iters == 4
        runs.append((losses, tr.p_flat.clone(), tr.m_flat.clone(), tr.v_flat.clone()))
    (l0, p0, m0, v0), (l1, p1, m1, v1), (l2, p2, m2, v2) = runs
    rel = lambda a, b_: ((a - b_).norm() / a.norm()).item()
    floor_m, floor_v, floor_p = rel(m0, m1), rel(v0, v1), (p0 - p1).abs().max().item()
    print("eager", l0, "graph", l2, "noise floor m/v/p", floor_m, floor_v, floor_p, "graph m/v/p", rel(m0, m2), rel(v0, v2), (p0 - p2).abs().max().item())
    assert np.allclose(l0, l2, rtol=1e-4)
    assert rel(m0, m2) <= max(3 * floor_m, 1e-3)
    assert rel(v0, v2) <= max(3 * floor_v, 1e-3)
    assert (p0 - p2).abs().max().item() <= max(3 * floor_p, 1e-5)


def test_gradients_are_additive_over_the_batch(text_dir):
    """Data-parallel correctness on one GPU (SURVEY.md §4 tier 5): the gradient of the 4-image batch equals the valid-pixel-weighted mean of
    the gradients of its two halves -- what two ranks holding the halves would all-reduce (no cross-sample op in the sk04 model) -- and the
    plain mean of the half-batch gradients is what `GradExchange` + the 1/world factor of AdamW produce for equal valid counts."""
    from semivl_b200.train import OptimCfg, Trainer
    crop, b = 64, 4
    m, mc, sd = _build(crop, True)
    g = torch.Generator().manual_seed(8)
    img = torch.randn(b, 3, crop, crop, generator=g).cuda()
    mask = torch.randint(0, 21, (b, crop, crop), generator=g)
    mask[:2, :10] = 255
    mask[2:, :, :10] = 255                      # same number of ignored pixels in both halves
    mask = mask.cuda()
    tr = Trainer(m, OptimCfg())
    tr.supervised_step(img, mask, update=False)
    g_full = tr.g_flat.clone()
    tr.supervised_step(img, mask, update=False)
    g_again = tr.g_flat.clone()
    tr.supervised_step(img[:2].contiguous(), mask[:2].contiguous(), update=False)
    g_a = tr.g_flat.clone()
    tr.supervised_step(img[2:].contiguous(), mask[2:].contiguous(), update=False)
    g_b = tr.g_flat.clone()
    na, nb = (mask[:2] != 255).sum().item(), (mask[2:] != 255).sum().item()
    assert na == nb
    want = 0.5 * (g_a + g_b)
    # The forward pass is bit-identical per image whatever batch it sits in (fixed-order GroupNorm statistics whose split count does not
    # depend on the number of maps; row-independent GEMM tiles), so the only differences left are the fp32 `red.global.add` orders of the
    # weight-gradient / bias-sum reductions: calibrate on two runs of the SAME batch and bound additivity at 3x that floor.
    floor = (g_full - g_again).norm().item() / g_full.norm().item()
    err = (g_full - want).norm().item() / g_full.norm().item()
    print(f"batch additivity of the gradient: rel {err:.3e} (same-batch run-to-run floor {floor:.3e})")
    assert err <= max(3 * floor, 2e-5)


def _sampled_grads(named, names, k=64):
    """Gradient entries at the positions oracle.make_golden.grad_summary recorded (64 seeded positions per tensor)."""
    from oracle.make_golden import sample_idx
    out = []
    for nme in names:
        gr = named[str(nme)]
        s = gr.detach().flatten()[sample_idx(gr.numel()).to(gr.device)].float().cpu().numpy()
        out.append(np.pad(s, (0, k - len(s))))
    return np.stack(out)


def _check_grad_samples(got, ref, norms, cos_min, what):
    """Element-wise gradient check against the reference golden (norms alone cannot see a permuted or transposed weight gradient):
    per-tensor cosine over the 64 recorded entries, for tensors whose gradient is not numerically zero."""
    worst = 1.0
    for i in range(len(ref)):
        a, b = got[i].astype(np.float64), ref[i].astype(np.float64)
        if norms[i] <= 1e-7 or np.abs(b).max() <= 1e-12:
            continue
        c = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))
        worst = min(worst, c)
        assert c > cos_min, (what, i, c)
    return worst


@pytest.mark.parametrize("precise", [True, False])
def test_fp_branches_of_forward_wrapper_match_reference_golden(golden_dir, precise):
    """`model(img, need_fp=True)` -> (pred, pred_fp) and `model(img, only_fp=True)` (model/builder.py:56-102) through the public call against
    outputs of the UNMODIFIED reference with the same injected dropout2d keep masks; backward of a loss over all three outputs:
    gradient norms and 64 sampled entries per tensor."""
    g = dict(np.load(os.path.join(golden_dir, "fp_c64_b2.npz"), allow_pickle=False))
    crop, b, nclass = int(g["crop"]), int(g["b"]), int(g["nclass"])
    m, mc, sd = _build(crop, precise, nclass)
    m.train()
    img = torch.from_numpy(g["img"]).cuda()
    masks = [torch.from_numpy(g[f"drop_mask{i}"]).cuda() for i in range(3)]
    gen = torch.Generator().manual_seed(int(g["seed"]))
    torch.randn(b, 3, crop, crop, generator=gen)
    for c in (768, 768, 512):
        torch.rand(b, c, 1, 1, generator=gen)
    wgt = [torch.randn(b, nclass, crop, crop, generator=gen).cuda() for _ in range(3)]
    pred, pred_fp = m(img, need_fp=True, drop_masks=masks)
    only = m(img, only_fp=True, drop_masks=masks)
    tol = 1e-3 if precise else 4e-2
    scale = np.abs(g["pred"]).max()
    for name, got in (("pred", pred), ("pred_fp", pred_fp), ("only_fp", only)):
        r = np.abs(got.detach().cpu().numpy() - g[name]).max() / scale
        print(f"{name} precise={precise}: rel {r:.2e}")
        assert r < tol, (name, r)
    # the perturbed branch must really differ from the clean one, and only_fp must equal the fp half of need_fp
    assert (pred - pred_fp).abs().max().item() > 10 * tol * scale or not precise
    assert torch.equal(only, pred_fp) or (only - pred_fp).abs().max().item() <= (1e-5 if precise else 1e-6 + 4e-2) * scale
    loss = (pred * wgt[0]).mean() + (pred_fp * wgt[1]).mean() + (only * wgt[2]).mean()
    assert abs(loss.item() - float(g["loss"])) < 3 * 0.8 * tol * scale          # |mean(d * w)| <= max|d| * E|w| per output
    loss.backward()
    if precise:
        named = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
        for nme, norm in zip(g["grad_names"], g["grad_norms"]):
            assert str(nme) in named, nme
            if norm > 1e-7:
                assert abs(named[str(nme)].double().norm().item() - norm) <= 3e-2 * norm, (nme, named[str(nme)].norm().item(), norm)
        worst = _check_grad_samples(_sampled_grads(named, g["grad_names"]), g["grad_samples"], g["grad_norms"], 0.995, "fp_c64_b2")
        print("worst per-tensor cosine of sampled gradient entries:", worst)


@pytest.mark.parametrize("name", ["maskclip_concept4_voc_c64_b2", "maskclip_concept3_city_c64_b2"])
@pytest.mark.parametrize("precise", [True, False])
def test_forward_maskclip_concept_tables_match_reference_golden(golden_dir, name, precise):
    """MaskCLIP pseudo-labels with the reference's real concept tables (VOC concept4: 98 rows -> 21 classes, Cityscapes concept3: 54 -> 19;
    model/vlm.py:98-109 + aggregate_concept_predictions, model/text_embeddings.py:188-193): svl_group_max with real group offsets."""
    from oracle import semivl_oracle as O
    from semivl_b200.model import build_model
    g = dict(np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False))
    crop, nclass = int(g["crop"]), int(g["nclass"])
    cfg = _cfg(crop, nclass, precise, mcc=str(g["mcc_text"]))
    cfg["dataset"] = str(g["dataset"])
    m = build_model(cfg)
    assert m.loaded_mcc_text_feat.shape[0] > nclass                  # a concept table really is in use
    sd = O.fixture_state_dict(O.param_shapes(O.ModelCfg(img_size=crop, num_classes=nclass)), seed=0)
    m.load_state_dict(sd)
    m = m.cuda()
    img = torch.from_numpy(g["img"]).cuda()
    for key, th in (("maskclip", 0.9), ("maskclip_lo", float(g["maskclip_lo_thresh"])), ("maskclip_all", 0.0)):
        got = m.forward_maskclip(img, th).cpu().numpy().astype(np.uint8)
        agree = (got == g[key]).mean()
        print(f"{name} {key} precise={precise}: agreement {agree:.5f}")
        assert agree > (0.999 if precise else 0.97), (key, agree)


def _adamw_reference(p0, grads, lrs_per_step, n_bb, betas=(0.9, 0.999), eps=1e-8, wd=0.01):
    """torch.optim.AdamW on the same flat parameters with the reference's two learning-rate classes (experiments.py:246-255) and the
    per-step rate rewrite of semivl.py:338-345 (the rate used by step t is passed in)."""
    p = p0.clone().double().requires_grad_(True)
    a, b_ = p[:n_bb].detach().clone().requires_grad_(True), p[n_bb:].detach().clone().requires_grad_(True)
    opt = torch.optim.AdamW([dict(params=[a], lr=1.0), dict(params=[b_], lr=1.0)], lr=1.0, betas=betas, eps=eps, weight_decay=wd)
    for gstep, (lr_bb, lr_hd) in zip(grads, lrs_per_step):
        opt.param_groups[0]["lr"], opt.param_groups[1]["lr"] = lr_bb, lr_hd
        a.grad, b_.grad = gstep[:n_bb].double().clone(), gstep[n_bb:].double().clone()
        opt.step()
    return torch.cat((a.detach(), b_.detach()))


@pytest.mark.parametrize("path", ["svl_adamw", "svl_adamw_dev"])
def test_adamw_five_steps_match_torch_optim(path):
    """Five optimizer steps of the fused flat-buffer AdamW -- the eager entry point and the device-scalar one the CUDA graph replays --
    against torch.optim.AdamW (float64) with two LR classes and the poly schedule applied in the reference's order (semivl.py:326-345).
    Step 1 alone cannot see a wrong beta or bias correction (m_hat / sqrt(v_hat) = sign(g)); five steps with changing gradients can."""
    from semivl_b200 import lib as L
    from semivl_b200.train import OptimCfg, Trainer
    n_bb, n_hd = 70001, 30003
    n = n_bb + n_hd
    gen = torch.Generator().manual_seed(77)
    p0 = torch.randn(n, generator=gen)
    grads = [torch.randn(n, generator=gen) * (10.0 ** float(torch.randint(-3, 1, (1,), generator=gen))) for _ in range(5)]
    tr = Trainer.__new__(Trainer)
    tr.opt = OptimCfg(lr=1e-3, total_iters=7, backbone_lr_mult=0.01, head_lr_mult=10.0)
    lrs = [(tr.lr_at(t) * 0.01, tr.lr_at(t) * 10.0) for t in range(1, 6)]
    assert lrs[2][0] < lrs[1][0] == lrs[0][0]                              # the poly decay is really exercised
    want = _adamw_reference(p0, grads, lrs, n_bb)
    p, m, v = p0.clone().cuda(), torch.zeros(n).cuda(), torch.zeros(n).cuda()
    o = tr.opt
    world = 2.0                                                               # gscale = 1/world folds the data-parallel mean into the kernel
    for t, gstep in enumerate(grads, start=1):
        gd = (gstep * world).cuda()
        if path == "svl_adamw":
            L.call("svl_adamw", p, gd, m, v, n_bb, lrs[t - 1][0], o.betas[0], o.betas[1], o.eps, o.wd, t, 1.0 / world)
            L.call("svl_adamw", p[n_bb:], gd[n_bb:], m[n_bb:], v[n_bb:], n_hd, lrs[t - 1][1], o.betas[0], o.betas[1], o.eps, o.wd, t, 1.0 / world)
        else:
            tr.iters = t - 1
            hyper = torch.tensor(tr._step_scalars(t), dtype=torch.float32).cuda()
            for lo, k, idx in ((0, n_bb, 0), (n_bb, n_hd, 1)):
                L.call("svl_adamw_dev", p[lo:], gd[lo:], m[lo:], v[lo:], k, hyper, idx, o.betas[0], o.betas[1], o.eps, o.wd, 1.0 / world)
    got = p.double().cpu()
    err = (got - want).abs().max().item()
    moved = (want - p0.double()).abs().max().item()
    print(f"{path}: max |p - torch.optim.AdamW| after 5 steps {err:.3e} (parameters moved by up to {moved:.3e})")
    assert err < 2e-6 and err < 1e-3 * moved
    # and a wrong beta2 WOULD be seen: the same run with beta2 = 0.99 in the reference differs by far more than the bound
    other = _adamw_reference(p0, grads, lrs, n_bb, betas=(0.9, 0.99))
    assert (other - want).abs().max().item() > 50 * err


def test_converted_clip_checkpoint_loads_and_matches_oracle(tmp_path, text_dir):
    """Checkpoint compatibility on the GPU: a CLIP-named visual-tower state dict is converted with
    semivl_b200.convert_clip_weights (third_party/maskclip/convert_clip_weights.py:27-64), saved, loaded through the `pretrained` path of
    MaskClipVisionTransformer.init_weights (maskclip_vit.py:378-410, incl. the position-embedding resize on load: the checkpoint has a 14 x 14
    grid, the model 4 x 4) and must reproduce the oracle's logits with the same weights; the released `best.pth` layout (DDP `module.` prefix,
    clip_encoder.* entries dropped; third_party/unimatch/eval.py:130-139) must load too."""
    from oracle import semivl_oracle as O
    from semivl_b200.convert_clip_weights import clip_visual_to_mmseg
    from semivl_b200.model import build_model
    crop, nclass = 64, 21
    mc = O.ModelCfg(img_size=crop, num_classes=nclass)
    gen = torch.Generator().manual_seed(123)
    E, Lyr, grid = 768, 12, 14
    clip = {"visual.class_embedding": torch.randn(E, generator=gen) * 0.02,
            "visual.positional_embedding": torch.randn(grid * grid + 1, E, generator=gen) * 0.02,
            "visual.conv1.weight": torch.randn(E, 3, 16, 16, generator=gen) * 0.02,
            "visual.ln_pre.weight": 1 + 0.1 * torch.randn(E, generator=gen), "visual.ln_pre.bias": 0.1 * torch.randn(E, generator=gen),
            "visual.ln_post.weight": 1 + 0.1 * torch.randn(E, generator=gen), "visual.ln_post.bias": 0.1 * torch.randn(E, generator=gen),
            "visual.proj": torch.randn(E, 512, generator=gen) * 0.03}
    for i in range(Lyr):
        pre = f"visual.transformer.resblocks.{i}."
        clip.update({pre + "attn.in_proj_weight": torch.randn(3 * E, E, generator=gen) * 0.02, pre + "attn.in_proj_bias": torch.randn(3 * E, generator=gen) * 0.02,
                     pre + "attn.out_proj.weight": torch.randn(E, E, generator=gen) * 0.02, pre + "attn.out_proj.bias": torch.randn(E, generator=gen) * 0.02,
                     pre + "ln_1.weight": 1 + 0.1 * torch.randn(E, generator=gen), pre + "ln_1.bias": 0.1 * torch.randn(E, generator=gen),
                     pre + "ln_2.weight": 1 + 0.1 * torch.randn(E, generator=gen), pre + "ln_2.bias": 0.1 * torch.randn(E, generator=gen),
                     pre + "mlp.c_fc.weight": torch.randn(4 * E, E, generator=gen) * 0.02, pre + "mlp.c_fc.bias": torch.randn(4 * E, generator=gen) * 0.02,
                     pre + "mlp.c_proj.weight": torch.randn(E, 4 * E, generator=gen) * 0.02, pre + "mlp.c_proj.bias": torch.randn(E, generator=gen) * 0.02})
    conv = clip_visual_to_mmseg(clip, backbone_prefix=True)
    ck = os.path.join(str(tmp_path), "clip2mmseg_ViT16_clip_backbone.pth")
    torch.save(conv, ck)
    cfg = _cfg(crop, nclass, True)
    cfg["model_args"] = dict(pretrained=ck)
    cfg["clip_encoder_args"] = dict(pretrained=ck)
    torch.manual_seed(5)
    m = build_model(cfg).cuda()
    sd_model = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    # the backbone must hold the checkpoint's tensors (pos_embed resized 14x14 -> 4x4 on load), not its random init
    assert torch.allclose(sd_model["backbone.layers.3.attn.attn.in_proj_weight"], clip["visual.transformer.resblocks.3.attn.in_proj_weight"])
    assert torch.allclose(sd_model["backbone.proj.weight"].flatten(1), clip["visual.proj"].t())
    assert sd_model["backbone.pos_embed"].shape == (1, (crop // 16) ** 2 + 1, E)
    want_pos = O.resize_pos_embed(conv["state_dict"]["backbone.pos_embed"], (crop // 16, crop // 16), (grid, grid))
    assert torch.allclose(sd_model["backbone.pos_embed"], want_pos, atol=1e-6)
    img = torch.randn(2, 3, crop, crop, generator=gen)
    text = torch.from_numpy(np.load(os.path.join(text_dir, "voc12_wbg_single.npy")))
    ref = O.model_forward(img, sd_model, text, mc)
    y = m(img.cuda())
    r = _rel(y, ref)
    print("converted checkpoint: logits rel vs oracle with the loaded weights", r)
    assert r < 1e-3
    # released best.pth layout: {'model': DDP state dict}; eval.py strips `module.` and drops clip_encoder.*
    best = {"model": {"module." + k: v for k, v in sd_model.items()}}
    new = {k.replace("module.", "", 1): v for k, v in best["model"].items() if "clip_encoder" not in k}
    torch.manual_seed(6)
    cfg2 = _cfg(crop, nclass, True)
    m2 = build_model(cfg2)
    missing = m2.load_state_dict(new, strict=False)
    assert all(k.startswith("clip_encoder.") for k in missing.missing_keys) and not missing.unexpected_keys
    y2 = m2.cuda()(img.cuda())
    assert _rel(y2, ref) < 1e-3


def test_semivl_step_head_groups_are_exact(golden_dir):
    """The SemiVL step cuts the gradient-tracked head batch into image groups to bound activation memory (Trainer._head_chunk; BASELINE
    configs 4 / 5).  No head op mixes images, so one image per group must reproduce the unsplit step: same loss terms, gradients equal up to
    the fp32 reduction-order floor measured on two unsplit runs."""
    from semivl_b200.train import OptimCfg, Trainer
    g = dict(np.load(os.path.join(golden_dir, "step_c96_b2.npz"), allow_pickle=False))
    crop, b = int(g["crop"]), int(g["b"])
    lk = ("mask_x", "ignore_mask", "ignore_mask_other")
    batch = {k: torch.from_numpy(g[k].astype(np.int64) if k in lk else g[k]).cuda()
             for k in ("img_x", "img_w", "img_s1", "img_s2", "img_w_other", "img_s1_other", "img_s2_other",
                       "mask_x", "ignore_mask", "ignore_mask_other", "mix1", "mix2")}
    hp = dict(conf_thresh=float(g["hp_conf_thresh"]), conf_mode=str(g["hp_conf_mode"]), mcc_conf_thresh=float(g["hp_mcc_conf_thresh"]),
              mcc_loss_reduce=str(g["hp_mcc_loss_reduce"]), mcc_lambda=float(g["hp_mcc_lambda"]))
    masks = [torch.from_numpy(g[f"drop_mask{i}"])[b:, :, 0, 0].contiguous().cuda() for i in range(3)]
    m, mc, sd = _build(crop, True, int(g["nclass"]))
    m.train()
    runs = []
    for budget in (None, None, 1.0):
        tr = Trainer(m, OptimCfg(), hp=hp, head_chunk_bytes=budget)
        assert tr._head_chunk(21, 24, 24) == (1 if budget else 1 << 30)
        loss, terms = tr.semivl_step(batch, drop_masks=masks, update=False)
        runs.append((loss.item(), np.array([v.item() for v in terms.values()]), tr.g_flat.clone()))
    (l0, t0, g0), (l1, t1, g1), (l2, t2, g2) = runs
    floor = ((g0 - g1).norm() / g0.norm()).item()
    err = ((g0 - g2).norm() / g0.norm()).item()
    print(f"one image per head group vs unsplit: loss {l2:.7f} vs {l0:.7f}, gradient rel {err:.3e} (unsplit run-to-run floor {floor:.3e})")
    assert abs(l2 - l0) <= 2e-6 * abs(l0) and np.abs(t2 - t0).max() <= 2e-6 * np.abs(t0).max()
    assert err <= max(3 * floor, 2e-5)


def test_trainer_checkpoint_resume_and_weight_reload(text_dir):
    """Trainer.state_dict / load_state_dict (flat parameters, Adam moments, step counter) resume a run exactly, and a model.load_state_dict
    after the trainer has run is honoured once invalidate_weights() drops the cached operand copies (the parameters are .data views
    whose version counter never moves)."""
    from semivl_b200.train import OptimCfg, Trainer
    crop, b = 64, 2
    g = torch.Generator().manual_seed(21)
    imgs = [torch.randn(b, 3, crop, crop, generator=g).cuda() for _ in range(4)]
    masks = [torch.randint(0, 21, (b, crop, crop), generator=g).cuda() for _ in range(4)]
    m, mc, sd = _build(crop, True)
    tr = Trainer(m, OptimCfg(lr=1e-4, total_iters=10))
    for i in range(2):
        tr.supervised_step(imgs[i], masks[i])
    ck = tr.state_dict()
    ref = [tr.supervised_step(imgs[i], masks[i]).item() for i in (2, 3)]
    p_ref = tr.p_flat.clone()
    m2, _, _ = _build(crop, True)
    tr2 = Trainer(m2, OptimCfg(lr=1e-4, total_iters=10))
    tr2.supervised_step(imgs[0], masks[0])                       # warms the operand caches with OTHER weights
    tr2.load_state_dict(ck)
    assert tr2.iters == 2
    got = [tr2.supervised_step(imgs[i], masks[i]).item() for i in (2, 3)]
    assert np.allclose(got, ref, rtol=2e-5), (got, ref)
    # the two runs take the same two steps from the same state; they differ by the fp32 reduction order of the weight gradients, which Adam's
    # m / sqrt(v) turns into up to ~lr per element where a gradient is near zero: compare with the distance the two steps moved the weights
    moved = (p_ref - ck["p_flat"]).norm().item()
    assert (tr2.p_flat - p_ref).norm().item() <= 0.05 * moved, ((tr2.p_flat - p_ref).norm().item(), moved)
    # a weight reload behind the trainer's back: frozen tensors included
    sd_b = {k: (v * 1.5 if k == "backbone.layers.3.ffn.layers.0.0.weight" else v) for k, v in sd.items()}
    l_old = tr2.supervised_step(imgs[0], masks[0], update=False).item()
    m2.load_state_dict(sd_b)
    tr2.invalidate_weights()
    l_new = tr2.supervised_step(imgs[0], masks[0], update=False).item()
    m3, _, _ = _build(crop, True)
    m3.load_state_dict(sd_b)
    m3 = m3.cuda()
    l_want = Trainer(m3, OptimCfg()).supervised_step(imgs[0], masks[0], update=False).item()
    assert abs(l_new - l_want) < 1e-5 * abs(l_want) and abs(l_new - l_old) > 1e-6


def test_operand_refresh_and_gradient_scatter_jobs():
    """svl_param_jobs through WeightCache: after the parameter changes, ONE launch rewrites every registered operand copy (gather path, the
    tiled-transpose path for non-multiple-of-32 shapes, fp32 and zero-padded layouts), and the staged-gradient scatter adds into the
    parameter layout and re-zeroes the staging buffer."""
    from semivl_b200 import lib as L
    from semivl_b200.engine.head import HeadEngine
    from semivl_b200.engine.vit import WeightCache
    L.check_device()
    g = torch.Generator().manual_seed(5)
    cache = WeightCache()
    cache.batched = True
    params = {"a": torch.randn(70, 45, generator=g).cuda(), "b": torch.randn(64, 32, 3, 3, generator=g).cuda(), "c": torch.randn(8, 1, 5, 5, generator=g).cuda(),
              "d": torch.randn(1, 32, 3, 3, generator=g).cuda()}
    cache.volatile = set(params)
    lays = {("a", "lin_t"): HeadEngine._layout("lin_t"), ("a", "lin"): HeadEngine._layout("lin"), ("b", "conv"): HeadEngine._layout("conv"),
            ("b", "conv_t"): HeadEngine._layout("conv_t"), ("c", "conv1"): HeadEngine._layout("conv1"), ("d", "out1"): HeadEngine._layout("out1")}
    ops_ = {k: cache.get_layout(k, params[k[0]], fn, False, raw_f32=k[1] == "out1") for k, fn in lays.items()}
    assert cache._jobs[("a", "lin_t")][3] & 2 and not cache._jobs[("a", "lin")][3] & 2          # the transpose is recognised, the plain copy is not
    for p in params.values():
        p.data.mul_(-1.7).add_(0.3)                                # "optimizer step" behind the cache's back: `.data` edits keep the version counter
    cache.refresh()
    torch.cuda.synchronize()
    for k, fn in lays.items():
        want = fn(params[k[0]].float()).contiguous()
        want = want if k[1] == "out1" else want.to(torch.bfloat16)
        assert ops_[k].data_ptr() == cache.get_layout(k, params[k[0]], fn, False, raw_f32=k[1] == "out1").data_ptr()      # persistent tensors
        assert torch.equal(ops_[k], want), k
    # gradient scatter
    grad = torch.randn(64, 32, 3, 3, generator=g).cuda()
    g0 = grad.clone()
    st = cache.grad_staging(("b", "conv"), grad, HeadEngine._layout("conv"))
    dw = torch.randn(st.shape, generator=g).cuda()
    st.copy_(dw)
    cache.scatter_grads()
    torch.cuda.synchronize()
    assert torch.allclose(grad, g0 + dw.view(3, 3, 64, 32).permute(2, 3, 0, 1)) and float(st.abs().max()) == 0.0
