"""The CPU oracle (oracle/semivl_oracle.py) against the golden vectors recorded from the
unmodified reference (oracle/make_golden.py).  Tolerances: fp32-vs-fp32 on CPU, different
op order only -> 2e-5 absolute on logits (range ~1.5), 1e-4 relative on the loss, 2e-3 on grads."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import semivl_oracle as O
from oracle.make_golden import sample_idx


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False))


def _params(mc, grad=True):
    sd = O.fixture_state_dict(O.param_shapes(mc), seed=0)
    return {k: v.clone().requires_grad_(grad) for k, v in sd.items()}


def _text(text_dir, name="voc12_wbg_single"):
    return torch.from_numpy(np.load(os.path.join(text_dir, name + ".npy")))


def _check_grads(g, p, rtol=2e-3):
    for name, norm, samp in zip(g["grad_names"], g["grad_norms"], g["grad_samples"]):
        gr = p[str(name)].grad
        assert gr is not None, name
        if norm < 1e-7:      # e.g. decode_head.head.bias: analytically zero (softmax grads sum to 0)
            continue
        assert abs(gr.double().norm().item() - norm) <= rtol * norm, (name, gr.norm().item(), norm)
        s = gr.flatten()[sample_idx(gr.numel())].numpy()
        scale = np.abs(samp[: len(s)]).max() + norm / np.sqrt(gr.numel())
        assert np.abs(s - samp[: len(s)]).max() <= rtol * scale * 5, name


@pytest.mark.parametrize("name", ["fwd_c64_b2", "fwd_c72_b1", "fwd_c224_b1"])
def test_forward_backward_matches_reference(golden_dir, text_dir, name):
    g = _load(golden_dir, name)
    mc = O.ModelCfg(img_size=int(g["crop"]), num_classes=int(g["nclass"]))
    p = _params(mc)
    img = torch.from_numpy(g["img"])
    lab = torch.from_numpy(g["label"].astype(np.int64))
    text = _text(text_dir)
    feats, glob = O.vit_forward(img, p, mc)
    assert np.abs(feats[-1].detach().numpy() - g["emb"]).max() < 2e-5
    assert np.abs(glob.detach().numpy() - g["global_emb"]).max() < 2e-5
    for i in (0, 1):
        s = feats[i].detach().flatten()[sample_idx(feats[i].numel(), 4096)].numpy()
        assert np.abs(s - g[f"feat{i}_sample"]).max() < 5e-5 * max(1.0, np.abs(s).max())
    y, low = O.model_forward(img, p, text, mc, return_lowres=True)
    assert np.abs(low.detach().numpy() - g["logits_lowres"]).max() < 2e-5
    ys = y.detach().flatten()[sample_idx(y.numel(), 8192)].numpy()
    assert np.abs(ys - g["logits_sample"]).max() < 2e-5
    assert (y.argmax(1).numpy() == g["argmax"]).mean() > 0.9995
    loss = F.cross_entropy(y, lab, ignore_index=255)
    assert abs(loss.item() - float(g["loss"])) < 1e-4 * float(g["loss"])
    loss.backward()
    _check_grads(g, p)
    for key, th in (("maskclip", 0.9), ("maskclip_lo", float(g["maskclip_lo_thresh"]))):
        mcl = O.forward_maskclip(img, p, text, mc, th, pos_img_size=512)
        assert (mcl.numpy().astype(np.uint8) == g[key]).mean() > 0.999


@pytest.mark.parametrize("name", ["step_c64_b1", "step_c96_b2", "step_c64_b2_pixelavg_mv", "step_c64_b2_pixelratio_mean"])
def test_semivl_iteration_matches_reference(golden_dir, text_dir, name):
    g = _load(golden_dir, name)
    mc = O.ModelCfg(img_size=int(g["crop"]), num_classes=int(g["nclass"]))
    p = _params(mc)
    text = _text(text_dir)
    lk = ("mask_x", "ignore_mask", "ignore_mask_other")
    batch = {k: torch.from_numpy(g[k].astype(np.int64) if k in lk else g[k])
             for k in ("img_x", "img_w", "img_s1", "img_s2", "img_w_other", "img_s1_other", "img_s2_other",
                       "mask_x", "ignore_mask", "ignore_mask_other", "mix1", "mix2")}
    hp = dict(conf_thresh=float(g["hp_conf_thresh"]), conf_mode=str(g["hp_conf_mode"]),
              mcc_conf_thresh=float(g["hp_mcc_conf_thresh"]), mcc_loss_reduce=str(g["hp_mcc_loss_reduce"]),
              mcc_lambda=float(g["hp_mcc_lambda"]))
    masks = [torch.from_numpy(g[f"drop_mask{i}"]) for i in range(3)]
    loss, terms = O.semivl_step_losses(batch, p, text, text, mc, hp, masks, clip_pos_img_size=512)
    got = np.array([terms[k].item() for k in ("loss_x", "loss_s1", "loss_s2", "loss_fp", "loss_mc_s1", "loss_mc_s2", "loss_mc_fp")])
    # pseudo-labels are argmaxes of near-degenerate logits: a handful of ties may flip -> 2e-3 rel on the terms
    assert np.abs(got - g["terms"]).max() < 2e-3 * np.abs(g["terms"]).max()
    assert abs(loss.item() - float(g["loss"])) < 1e-3 * float(g["loss"])
    loss.backward()
    _check_grads(g, p, rtol=1e-2)


def test_optimizer_rules():
    ck = {"backbone": dict(lr_mult=0.01), "text_encoder": dict(lr_mult=0.0), "conv_encoder": dict(lr_mult=1.0),
          "norm": dict(decay_mult=0.0), "ln": dict(decay_mult=0.0), "head": dict(lr_mult=10.0)}
    assert O.param_group_hparams("backbone.layers.0.ln1.weight", 1e-4, 0.01, ck) == pytest.approx((1e-6, 0.01))
    assert O.param_group_hparams("decode_head.conv1.weight", 1e-4, 0.01, ck) == pytest.approx((1e-3, 0.01))
    assert O.param_group_hparams("clip_encoder.ln0.weight", 1e-4, 0.01, ck) == pytest.approx((1e-4, 0.0))
    # AdamW restatement against torch.optim.AdamW
    torch.manual_seed(0)
    w = torch.randn(50, requires_grad=True)
    w2 = w.detach().clone()
    opt = torch.optim.AdamW([w], lr=1e-3, weight_decay=0.01)
    m, v = torch.zeros(50), torch.zeros(50)
    for step in range(1, 4):
        gr = torch.randn(50)
        w.grad = gr.clone()
        opt.step()
        O.adamw_step(w2, gr, m, v, step, 1e-3, 0.01)
    assert torch.allclose(w.detach(), w2, atol=1e-7)
    assert abs(O.poly_lr(1e-4, 50, 100) - 1e-4 * 0.5 ** 0.9) < 1e-12


def test_evaluation_restatement_matches_reference(golden_dir):
    """oracle predict_reference / intersection_and_union against `predict` and `intersectionAndUnion` of the unmodified reference
    (tests/golden/eval_predict_iou.npz, oracle/make_golden.py::eval_fixture)."""
    g = dict(np.load(os.path.join(golden_dir, "eval_predict_iou.npz"), allow_pickle=False))
    nclass = int(g["nclass"])
    model = O.StubSegModel(nclass, weight=g["weight"])
    for i, case in enumerate(g["cases"]):
        mode, h, w, crop, stride = str(case).split("|")
        cfg = dict(nclass=nclass, crop_size=int(crop), stride=float(stride) if "." in stride else int(stride))
        img, mask = torch.from_numpy(g[f"img{i}"]), torch.from_numpy(g[f"mask{i}"].astype(np.int64))
        pred, final = O.predict_reference(model, img, mask, mode, cfg)
        assert np.abs(final.numpy() - g[f"final{i}"]).max() < 1e-5, mode
        assert np.array_equal(pred.numpy().astype(np.uint8), g[f"pred{i}"]), mode
        m2 = mask
        if mode == "center_crop":
            c = int(crop)
            sh, sw = (int(h) - c) // 2, (int(w) - c) // 2
            m2 = mask[:, sh:sh + c, sw:sw + c]
        ai, au, at = O.intersection_and_union(g[f"pred{i}"].astype(np.int64), m2.numpy(), nclass, 255)
        assert np.array_equal(np.stack((ai, au, at)), g[f"iou{i}"]), mode


def test_head_conv_feature_branch_matches_reference(golden_dir):
    """oracle vlg_head_forward(conv_feats=...) against the unmodified reference VLGHead built with skip_from_conv_feat=True
    (vlg_head.py:196-205; tests/golden/head_convfeat_b2.npz): logits, input gradients and parameter-gradient norms."""
    from oracle.make_golden import HEAD_CONV_KW
    g = dict(np.load(os.path.join(golden_dir, "head_convfeat_b2.npz"), allow_pickle=False))
    kw = HEAD_CONV_KW
    mc = O.ModelCfg(img_size=kw["img_size"], num_classes=kw["num_classes"], skip_channels=kw["skip_channels"])
    names = [str(n) for n in g["grad_names"]]
    shapes = _head_shapes_with_conv(mc, kw)
    assert set(names) <= set(shapes)
    p = {k: v.clone().requires_grad_(True) for k, v in O.fixture_state_dict(shapes, seed=3).items()}
    v4, emb, conv = (torch.from_numpy(g[k]).requires_grad_(True) for k in ("v4", "emb", "conv"))
    out = O.vlg_head_forward([v4, emb], torch.from_numpy(g["text"]), p, mc, conv_feats=[conv])
    assert np.abs(out.detach().numpy() - g["out"]).max() < 2e-5 * np.abs(g["out"]).max()
    (out * torch.from_numpy(g["wgt"])).sum().backward()
    for name, t in (("d_v4", v4), ("d_emb", emb), ("d_conv", conv)):
        assert np.abs(t.grad.numpy() - g[name]).max() < 1e-4 * np.abs(g[name]).max() + 1e-7, name
    for n, norm in zip(names, g["grad_norms"]):
        if norm > 1e-7:
            assert abs(p[n].grad.double().norm().item() - norm) <= 1e-3 * norm, n


def _head_shapes_with_conv(mc, kw):
    """parameter shapes of the head with its second skip taken from a 256-channel conv feature (skip_in_channels = (768, 256))"""
    shapes = {k: v for k, v in O.param_shapes(mc, with_clip_encoder=False).items() if k.startswith("decode_head.")}
    shapes["decode_head.skip_proj.1.0.weight"] = (kw["skip_channels"][1], kw["skip_in_channels"][1], 3, 3)
    return shapes
