"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the authoring container (needs /root/reference):
    python -m oracle.make_golden
The reference modules are imported through oracle/ref_shim.py, loaded with the deterministic
fixture weights of oracle/semivl_oracle.fixture_state_dict and run on seeded inputs; the script
records inputs + reference outputs.  Weights are NOT stored (regenerated from the seed).
The training-iteration fixture drives the reference's own model.forward / forward_maskclip /
utils.train_utils functions in the order of semivl.py:224-323 (that loop lives under
`if __name__ == '__main__'` and cannot be imported).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle import semivl_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
TEXT_DIR = os.path.join(ROOT, "semivl_b200", "configs", "_base_", "datasets", "text_embedding")


def sample_idx(numel, k=64):
    g = torch.Generator().manual_seed(numel % 65521 + 17)
    return torch.randint(0, numel, (min(k, numel),), generator=g)


def grad_summary(named_grads):
    names, norms, samples = [], [], []
    for n, g in named_grads:
        names.append(n)
        norms.append(g.double().norm().item())
        s = g.flatten()[sample_idx(g.numel())].float().numpy()
        samples.append(np.pad(s, (0, 64 - len(s))))
    return dict(grad_names=np.array(names), grad_norms=np.array(norms), grad_samples=np.stack(samples))


def synth_batch(b, crop, nclass, seed):
    """Synthetic SemiVL batch: randn images, labels with a 255 block, pad-strip ignore masks, cutmix boxes."""
    g = torch.Generator().manual_seed(seed)
    r = lambda: torch.randn(b, 3, crop, crop, generator=g)
    batch = dict(img_x=r(), img_w=r(), img_s1=r(), img_s2=r(), img_w_other=r(), img_s1_other=r(), img_s2_other=r())
    mask_x = torch.randint(0, nclass, (b, crop, crop), generator=g)
    mask_x[:, : crop // 8, : crop // 3] = 255
    ign = torch.zeros(b, crop, crop, dtype=torch.long)
    ign[:, -crop // 10:, :] = 255
    ign_o = torch.zeros(b, crop, crop, dtype=torch.long)
    ign_o[:, :, -crop // 12:] = 255
    def box(frac):
        m = torch.zeros(b, crop, crop)
        h = int(crop * frac)
        m[:, crop // 5: crop // 5 + h, crop // 4: crop // 4 + h] = 1
        return m
    batch.update(mask_x=mask_x, ignore_mask=ign, ignore_mask_other=ign_o, mix1=box(0.5), mix2=box(0.3))
    return batch


def build(crop, nclass=21, dataset="pascal", text="single", mcc_text=None):
    cfg = ref_shim.default_cfg(dataset=dataset, nclass=nclass, crop_size=crop, text=text, mcc_text=mcc_text)
    m = ref_shim.build_reference_model(cfg)
    mc = O.ModelCfg(img_size=crop, num_classes=nclass)
    sd = O.fixture_state_dict(O.param_shapes(mc), seed=0)
    m.load_state_dict(sd)
    return m, mc, sd


def forward_fixture(name, crop, b, seed, nclass=21):
    m, mc, sd = build(crop, nclass)
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(b, 3, crop, crop, generator=g)
    lab = torch.randint(0, nclass, (b, crop, crop), generator=g)
    lab[:, : crop // 7, : crop // 2] = 255
    with ref_shim.in_reference_cwd():
        m.train()
        feats_all = m.extract_feat(img)
        feats, glob = feats_all[0]
        low = m.decode_head.forward(feats_all)
        y = m(img)
        loss = F.cross_entropy(y, lab, ignore_index=255)
        loss.backward()
        mcl = m.forward_maskclip(img, 0.9)
        # maskclip confidences are degenerate (all 255) at fixture scale: also record at a low threshold
        mcl_lo = m.forward_maskclip(img, 1.0 / nclass + 1e-3)
    out = dict(img=img.numpy(), label=lab.numpy().astype(np.int16), crop=crop, nclass=nclass, seed=seed,
               logits_lowres=low.detach().numpy(), emb=feats[-1].detach().numpy(),
               feat0_sample=feats[0].detach().flatten()[sample_idx(feats[0].numel(), 4096)].numpy(),
               feat1_sample=feats[1].detach().flatten()[sample_idx(feats[1].numel(), 4096)].numpy(),
               global_emb=glob.detach().numpy(), loss=np.float64(loss.item()),
               argmax=y.argmax(1).numpy().astype(np.uint8),
               logits_sample=y.detach().flatten()[sample_idx(y.numel(), 8192)].numpy(),
               logits_absmax=np.float64(y.abs().max().item()),
               maskclip=mcl.numpy().astype(np.uint8), maskclip_lo=mcl_lo.numpy().astype(np.uint8),
               maskclip_lo_thresh=np.float64(1.0 / nclass + 1e-3))
    out.update(grad_summary([(n, q.grad) for n, q in m.named_parameters() if q.grad is not None]))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "loss", loss.item(), "ngrads", len(out["grad_names"]))


def step_fixture(name, crop, b, seed, nclass=21, hp_over=None):
    """semivl.py:224-323 driven on the reference model with injected dropout2d masks."""
    m, mc, sd = build(crop, nclass)
    import model.builder as ref_builder
    from utils.train_utils import confidence_weighted_loss, cutmix_img_, cutmix_mask
    batch = synth_batch(b, crop, nclass, seed)
    g = torch.Generator().manual_seed(seed + 1)
    drop_masks = [(torch.rand(2 * b, c, 1, 1, generator=g) >= 0.5).float() for c in (768, 768, 512)]
    hp = dict(conf_thresh=1.0 / nclass + 2e-3, conf_mode="pixelwise", mcc_conf_thresh=1.0 / nclass + 1e-3,
              mcc_loss_reduce="mean_all", mcc_lambda=0.07)
    hp.update(hp_over or {})
    cfgd = dict(conf_mode=hp["conf_mode"], conf_thresh=hp["conf_thresh"])

    calls = {"i": 0}
    real_dropout2d = F.dropout2d

    def fake_dropout2d(f, p=0.5, training=True, inplace=False):
        mk = drop_masks[calls["i"] % 3]
        calls["i"] += 1
        return f * mk / (1 - p)
    ref_builder.F.dropout2d = fake_dropout2d
    try:
        with ref_shim.in_reference_cwd():
            bt = {k: v.clone() for k, v in batch.items()}
            cutmix_img_(bt["img_s1"], bt["img_s1_other"], bt["mix1"])
            cutmix_img_(bt["img_s2"], bt["img_s2_other"], bt["mix2"])
            with torch.no_grad():
                m.eval()
                pred_w_other = m(bt["img_w_other"]).detach()
                conf_w_other, mask_w_other = pred_w_other.softmax(dim=1).max(dim=1)
                mclip = m.forward_maskclip(torch.cat((bt["img_w"], bt["img_w_other"])), conf_tresh=hp["mcc_conf_thresh"])
                mclip, mclip_other = mclip.split([b, b])
                mclip[bt["ignore_mask"] == 255] = 255
                mclip_other[bt["ignore_mask_other"] == 255] = 255
            m.train()
            preds, preds_fp = m(torch.cat((bt["img_x"], bt["img_w"])), need_fp=True)
            pred_x, pred_w = preds.chunk(2)
            _, pred_w_fp = preds_fp.chunk(2)
            pred_s1, pred_s2 = m(torch.cat((bt["img_s1"], bt["img_s2"]))).chunk(2)
            pred_w = pred_w.detach()
            conf_w, mask_w = pred_w.softmax(dim=1).max(dim=1)
            mix1, mix2 = bt["mix1"], bt["mix2"]
            mm1, mm2 = cutmix_mask(mask_w, mask_w_other, mix1), cutmix_mask(mask_w, mask_w_other, mix2)
            cm1, cm2 = cutmix_mask(conf_w, conf_w_other, mix1), cutmix_mask(conf_w, conf_w_other, mix2)
            im1 = cutmix_mask(bt["ignore_mask"], bt["ignore_mask_other"], mix1)
            im2 = cutmix_mask(bt["ignore_mask"], bt["ignore_mask_other"], mix2)
            mc1, mc2 = cutmix_mask(mclip, mclip_other, mix1), cutmix_mask(mclip, mclip_other, mix2)
            cu = torch.nn.CrossEntropyLoss(reduction="none")
            cmc = torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none")
            loss_x = F.cross_entropy(pred_x, bt["mask_x"], ignore_index=255)
            loss_s1 = confidence_weighted_loss(cu(pred_s1, mm1), cm1, im1, cfgd)
            loss_s2 = confidence_weighted_loss(cu(pred_s2, mm2), cm2, im2, cfgd)
            loss_fp = confidence_weighted_loss(cu(pred_w_fp, mask_w), conf_w, bt["ignore_mask"], cfgd)

            def compute_mc_loss(pred, mask, ign):            # semivl.py:52-58 (module-level there, bound to __main__ globals)
                if hp["mcc_loss_reduce"] == "mean":
                    return torch.nn.CrossEntropyLoss(ignore_index=255)(pred, mask)
                l_mc = cmc(pred, mask)
                if hp["mcc_loss_reduce"] == "mean_valid":
                    return l_mc.sum() / (ign != 255).sum()
                return l_mc.sum() / ign.numel()
            l_mc1 = compute_mc_loss(pred_s1, mc1, im1)
            l_mc2 = compute_mc_loss(pred_s2, mc2, im2)
            l_mcf = compute_mc_loss(pred_w_fp, mclip, bt["ignore_mask"])
            lam = hp["mcc_lambda"]
            loss = (loss_x + loss_s1 * 0.25 + loss_s2 * 0.25 + loss_fp * 0.5) / 2.0
            loss = loss + l_mc1 * 0.25 * lam + l_mc2 * 0.25 * lam + l_mcf * 0.5 * lam
            loss.backward()
    finally:
        ref_builder.F.dropout2d = real_dropout2d
    out = {k: (v.numpy().astype(np.int16) if v.dtype == torch.long else v.numpy()) for k, v in batch.items()}
    out.update({f"drop_mask{i}": d.numpy() for i, d in enumerate(drop_masks)})
    out.update({f"hp_{k}": np.array(v) for k, v in hp.items()})
    out.update(crop=crop, nclass=nclass, seed=seed, b=b, loss=np.float64(loss.item()),
               terms=np.array([t.item() for t in (loss_x, loss_s1, loss_s2, loss_fp, l_mc1, l_mc2, l_mcf)], dtype=np.float64),
               conf_frac=np.float64((conf_w >= hp["conf_thresh"]).float().mean().item()),
               mclip_valid_frac=np.float64((mclip != 255).float().mean().item()),
               mask_w=mask_w.numpy().astype(np.uint8), mclip=mclip.numpy().astype(np.uint8))
    out.update(grad_summary([(n, q.grad) for n, q in m.named_parameters() if q.grad is not None]))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "loss", loss.item(), "terms", out["terms"], "conf_frac", out["conf_frac"], "mclip_valid", out["mclip_valid_frac"])


EVAL_CASES = [("original", 33, 50, 24, 0), ("center_crop", 33, 50, 24, 0), ("padded_sliding_window", 33, 50, 24, 16),
              ("padded_sliding_window", 33, 50, 24, 0.5), ("zegclip_sliding_window", 33, 50, 24, 17), ("sliding_window", 33, 50, 21, 0)]


def eval_fixture(name="eval_predict_iou"):
    """`predict` (third_party/unimatch/supervised.py:40-130) and `intersectionAndUnion` (third_party/unimatch/util/utils.py:91-103)
    of the UNMODIFIED reference on a stand-in model.  supervised.py cannot be imported whole here (its tqdm / mmseg imports), so
    the `predict` function is compiled from the reference file as it lies under /root/reference (nothing is copied); `.cuda()` is
    the identity while it runs on this GPU-less box."""
    import ast
    import importlib
    import types
    ref_root = ref_shim.REF_ROOT if hasattr(ref_shim, "REF_ROOT") else "/root/reference"
    path = os.path.join(ref_root, "third_party", "unimatch", "supervised.py")
    tree = ast.parse(open(path).read(), filename=path)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "predict"]
    mmseg = types.SimpleNamespace(ops=types.SimpleNamespace(resize=lambda x, size, mode, align_corners, warning=False:
                                                          F.interpolate(x, size=size, mode=mode, align_corners=align_corners)))
    ns = dict(torch=torch, F=F, mmseg=mmseg)
    exec(compile(ast.Module(body=fn, type_ignores=[]), path, "exec"), ns)
    with ref_shim.in_reference_cwd():
        ref_iou = importlib.import_module("third_party.unimatch.util.utils").intersectionAndUnion
    nclass = 5
    model = O.StubSegModel(nclass, seed=4)
    out = dict(weight=model.w.detach().numpy(), nclass=nclass, cases=np.array([f"{m}|{h}|{w}|{c}|{s}" for m, h, w, c, s in EVAL_CASES]))
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for i, (mode, h, w, crop, stride) in enumerate(EVAL_CASES):
            g = torch.Generator().manual_seed(40 + i)
            img = torch.randn(1, 3, h, w, generator=g)
            mask = torch.randint(0, nclass, (1, h, w), generator=g)
            mask[:, :4] = 255
            pred, final = ns["predict"](model, img, mask, mode, dict(nclass=nclass, crop_size=crop, stride=stride), return_logits=True)
            m2 = mask
            if mode == "center_crop":
                sh, sw = (h - crop) // 2, (w - crop) // 2
                m2 = mask[:, sh:sh + crop, sw:sw + crop]
            ai, au, at = ref_iou(pred.numpy(), m2.numpy(), nclass, 255)
            out.update({f"img{i}": img.numpy(), f"mask{i}": mask.numpy().astype(np.int16), f"pred{i}": pred.numpy().astype(np.uint8),
                        f"final{i}": final.numpy(), f"iou{i}": np.stack((ai, au, at)).astype(np.int64)})
    finally:
        torch.Tensor.cuda = real_cuda
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "cases", len(EVAL_CASES))


INPUT_CASES = [(40, 56, 32, 255), (30, 50, 48, 254), (64, 20, 32, 254), (33, 33, 33, 255)]       # (h, w, crop size, pad value)


def input_fixture(name="input_stage"):
    """crop / hflip / normalize / obtain_cutmix_box of the UNMODIFIED reference (third_party/unimatch/dataset/transform.py), driven
    with seeded `random` / `numpy.random`; the ignore mask follows semi.py:99-103."""
    import importlib
    import random
    from PIL import Image
    sys.path.insert(0, "/root/reference")
    T = importlib.import_module("third_party.unimatch.dataset.transform")
    out = dict(cases=np.array([f"{h}|{w}|{s}|{pad}" for h, w, s, pad in INPUT_CASES]))
    for i, (h, w, size, pad) in enumerate(INPUT_CASES):
        rng = np.random.RandomState(70 + i)
        img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        mask = rng.randint(0, 21, (h, w)).astype(np.uint8)
        random.seed(500 + i)
        pi, pm = T.crop(Image.fromarray(img), Image.fromarray(mask), size, pad)
        pi, pm = T.hflip(pi, pm, p=0.5)
        ti, tm = T.normalize(pi, pm)
        ign = torch.zeros(size, size).long()                    # semi.py:99-103
        ign[tm == 254] = 255
        out.update({f"img{i}": img, f"mask{i}": mask, f"out_img{i}": ti.numpy(), f"out_mask{i}": tm.numpy().astype(np.int16),
                    f"out_ign{i}": ign.numpy().astype(np.int16)})
    boxes = []
    for j in range(12):
        random.seed(900 + j)
        np.random.seed(900 + j)
        boxes.append(T.obtain_cutmix_box(48, p=0.5).numpy().astype(np.uint8))
    out["boxes"] = np.stack(boxes)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "cases", len(INPUT_CASES), "non-empty boxes", int(sum(b.any() for b in boxes)))


HEAD_CONV_KW = dict(img_size=64, num_classes=5, text_in_channels=512, text_channels=128, up_channels=(64, 32), skip_in_channels=(768, 256),
                    skip_channels=(32, 32), skip_from_conv_feat=True, num_layers=2, num_heads=4, channels=128, pool_size=(4, 4), conv1_ksize=7,
                    loss_decode=None, align_corners=False)


def head_conv_fixture(name="head_convfeat_b2"):
    """The UNMODIFIED reference VLGHead with `skip_from_conv_feat=True` (the Cityscapes skr04 head, vlg_head.py:196-205): one ViT tap + the
    CLIP embedding at 4 x 4 tokens and a 256-channel conv-encoder feature at 16 x 16 (the ResNet stem's stride 4); forward and backward."""
    import importlib
    ref_shim.install()
    with ref_shim.in_reference_cwd():
        mod = importlib.import_module("model.decode_heads.vlg_head")
        head = mod.VLGHead(**HEAD_CONV_KW)
    shapes = {"decode_head." + k: tuple(v.shape) for k, v in head.state_dict().items()}
    sd = O.fixture_state_dict(shapes, seed=3)
    head.load_state_dict({k[len("decode_head."):]: v for k, v in sd.items()})
    g = torch.Generator().manual_seed(31)
    B, N = 2, HEAD_CONV_KW["num_classes"]
    v4 = torch.randn(B, 768, 4, 4, generator=g).requires_grad_(True)
    emb = torch.randn(B, 512, 4, 4, generator=g).requires_grad_(True)
    conv = torch.randn(B, 256, 16, 16, generator=g).requires_grad_(True)
    text = torch.randn(N, 512, generator=g)
    wgt = torch.randn(B, N, 16, 16, generator=g)
    head.train()
    out = head([[[v4, emb], None], text, [conv]])
    (out * wgt).sum().backward()
    res = dict(v4=v4.detach().numpy(), emb=emb.detach().numpy(), conv=conv.detach().numpy(), text=text.numpy(), wgt=wgt.numpy(),
               out=out.detach().numpy(), d_v4=v4.grad.numpy(), d_emb=emb.grad.numpy(), d_conv=conv.grad.numpy())
    res.update(grad_summary([("decode_head." + n, q.grad) for n, q in head.named_parameters() if q.grad is not None]))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **res)
    print(name, "out", tuple(out.shape), "absmax", out.abs().max().item(), "ngrads", len(res["grad_names"]))


def fp_fixture(name="fp_c64_b2", crop=64, b=2, seed=14, nclass=21):
    """The feature-perturbation branches of the reference's patched forward (model/builder.py:56-102) through the public call:
    `model(img, need_fp=True)` -> (pred, pred_fp) and `model(img, only_fp=True)`, with F.dropout2d replaced by recorded keep masks
    (SURVEY.md §8c caveat ii); forward + backward of a loss that weights the three outputs differently."""
    m, mc, sd = build(crop, nclass)
    import model.builder as ref_builder
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(b, 3, crop, crop, generator=g)
    drop_masks = [(torch.rand(b, c, 1, 1, generator=g) >= 0.5).float() for c in (768, 768, 512)]
    wgt = [torch.randn(b, nclass, crop, crop, generator=g) for _ in range(3)]
    calls = {"i": 0}
    real_dropout2d = F.dropout2d

    def fake_dropout2d(f, p=0.5, training=True, inplace=False):
        mk = drop_masks[calls["i"] % 3]
        calls["i"] += 1
        return f * mk / (1 - p)
    ref_builder.F.dropout2d = fake_dropout2d
    try:
        with ref_shim.in_reference_cwd():
            m.train()
            pred, pred_fp = m(img, need_fp=True)
            only = m(img, only_fp=True)
            loss = (pred * wgt[0]).mean() + (pred_fp * wgt[1]).mean() + (only * wgt[2]).mean()
            loss.backward()
    finally:
        ref_builder.F.dropout2d = real_dropout2d
    assert calls["i"] == 6
    out = dict(img=img.numpy(), crop=crop, nclass=nclass, b=b, seed=seed, loss=np.float64(loss.item()),
               pred=pred.detach().numpy().astype(np.float32), pred_fp=pred_fp.detach().numpy().astype(np.float32),
               only_fp=only.detach().numpy().astype(np.float32), wgt_seed_note="weights = 3 x randn after img and masks from the same generator")
    out.update({f"drop_mask{i}": d.numpy() for i, d in enumerate(drop_masks)})
    out.update(grad_summary([(n, q.grad) for n, q in m.named_parameters() if q.grad is not None]))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "loss", loss.item(), "fp differs from clean:", (pred - pred_fp).abs().max().item(), "ngrads", len(out["grad_names"]))


def concept_fixture(name, dataset, nclass, mcc_text, crop=64, b=2, seed=15):
    """`forward_maskclip` with a concept text table (model/vlm.py:98-109 + aggregate_concept_predictions, model/text_embeddings.py:188-193):
    the reference's real VOC (`concept4_single`, 98 rows -> 21) and Cityscapes (`concept3_single`, 54 -> 19) settings."""
    m, mc, sd = build(crop, nclass, dataset=dataset, mcc_text=mcc_text)
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(b, 3, crop, crop, generator=g)
    th_lo = 1.0 / nclass + 1e-3
    with ref_shim.in_reference_cwd():
        m.eval()
        with torch.no_grad():
            mcl = m.forward_maskclip(img, 0.9)
            mcl_lo = m.forward_maskclip(img, th_lo)
            mcl_all = m.forward_maskclip(img, 0.0)
    out = dict(img=img.numpy(), crop=crop, nclass=nclass, dataset=dataset, mcc_text=mcc_text, maskclip=mcl.numpy().astype(np.uint8),
               maskclip_lo=mcl_lo.numpy().astype(np.uint8), maskclip_all=mcl_all.numpy().astype(np.uint8), maskclip_lo_thresh=np.float64(th_lo))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "classes used", np.unique(out["maskclip_all"]).tolist(), "valid at lo", float((out["maskclip_lo"] != 255).mean()))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    if "--only-new" in sys.argv:                            # round-2 additions only (the older files regenerate bit-identically)
        torch.manual_seed(0)
        torch.set_num_threads(8)
        fp_fixture()
        concept_fixture("maskclip_concept4_voc_c64_b2", "pascal", 21, "concept4_single")
        concept_fixture("maskclip_concept3_city_c64_b2", "cityscapes", 19, "concept3_single")
        return
    torch.manual_seed(0)
    torch.set_num_threads(8)
    forward_fixture("fwd_c64_b2", 64, 2, seed=11)          # tiny: every op, multiple of 16
    forward_fixture("fwd_c72_b1", 72, 1, seed=12)          # corner pad + bicubic pos-embed resize (maskclip_vit.py:448-459)
    forward_fixture("fwd_c224_b1", 224, 1, seed=13)        # BASELINE config 1
    step_fixture("step_c64_b1", 64, 1, seed=21)            # full SemiVL iteration (semivl.py:224-323)
    step_fixture("step_c96_b2", 96, 2, seed=22)
    # Cityscapes-style loss modes (experiments.py:451 conf_mode='pixelavg'; train_utils.py:40-46; semivl.py:52-58)
    step_fixture("step_c64_b2_pixelavg_mv", 64, 2, seed=23, hp_over=dict(conf_mode="pixelavg", mcc_loss_reduce="mean_valid"))
    step_fixture("step_c64_b2_pixelratio_mean", 64, 2, seed=24, hp_over=dict(conf_mode="pixelratio", mcc_loss_reduce="mean"))
    eval_fixture()
    input_fixture()
    head_conv_fixture()
    fp_fixture()
    concept_fixture("maskclip_concept4_voc_c64_b2", "pascal", 21, "concept4_single")
    concept_fixture("maskclip_concept3_city_c64_b2", "cityscapes", 19, "concept3_single")


if __name__ == "__main__":
    main()
