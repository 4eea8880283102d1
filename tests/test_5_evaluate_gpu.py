"""Evaluation path (third_party/unimatch/supervised.py:40-164): the five `predict` modes and the mIoU histograms against the
CPU restatement in oracle/semivl_oracle.py.  The model is a small deterministic stand-in (same weights on both sides) so that the
test isolates the window stitching; integer results (labels, histograms) are compared bit-exactly."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _Stub(nclass):
    from oracle import semivl_oracle as O
    return O.StubSegModel(nclass)


@pytest.mark.parametrize("mode,h,w,crop,stride", [
    ("original", 40, 52, 32, 0), ("center_crop", 40, 52, 32, 0), ("padded_sliding_window", 45, 70, 32, 20),
    ("padded_sliding_window", 45, 70, 32, 0.5), ("zegclip_sliding_window", 45, 70, 32, 21), ("zegclip_sliding_window", 32, 32, 32, 21),
    ("sliding_window", 50, 77, 30, 0)])
def test_predict_modes_match_reference(mode, h, w, crop, stride):
    from oracle import semivl_oracle as O
    from semivl_b200 import evaluate as E
    nclass = 7
    cfg = dict(nclass=nclass, crop_size=crop, stride=stride)
    g = torch.Generator().manual_seed(3)
    img = torch.randn(2, 3, h, w, generator=g)
    mask = torch.randint(0, nclass, (2, h, w), generator=g)
    stub = _Stub(nclass)
    ref_pred, ref_final = O.predict_reference(stub, img, mask, mode, cfg)
    pred, final = E.predict(stub.cuda(), img.cuda(), mask, mode, cfg, return_logits=True)
    assert pred.dtype == torch.int64 and tuple(pred.shape) == tuple(ref_pred.shape)
    err = (final.cpu() - ref_final).abs().max().item()
    assert err < 2e-5 * max(1.0, ref_final.abs().max().item()), err
    top2 = ref_final.topk(2, dim=1).values
    decidable = (top2[:, 0] - top2[:, 1]) > 4 * err
    assert decidable.float().mean() > 0.99
    assert torch.equal(pred.cpu()[decidable], ref_pred[decidable])
    # the arg-max kernel itself is exact on the GPU's own scores (first maximal index, like torch.argmax)
    assert torch.equal(pred, final.argmax(dim=1))


@pytest.mark.parametrize("K,n", [(21, 100000), (150, 512 * 512 * 2), (2, 17), (19, 0)])
def test_intersection_union_bit_exact(K, n):
    from oracle import semivl_oracle as O
    from semivl_b200 import evaluate as E
    g = torch.Generator().manual_seed(K)
    pred = torch.randint(0, K, (n,), generator=g)
    target = torch.randint(0, K, (n,), generator=g)
    if n:
        target[torch.rand(n, generator=g) < 0.1] = 255
        hit = torch.rand(n, generator=g) < 0.3
        pred[hit] = target[hit]                          # guaranteed intersections (and predictions of 255 on ignored pixels)
    counts = E.intersection_union_counts(pred.cuda(), target.cuda(), K, 255)
    counts = E.intersection_union_counts(pred.cuda(), target.cuda(), K, 255, counts)          # accumulates
    ai, au, at = O.intersection_and_union(pred.numpy(), target.numpy(), K, 255)
    c = counts.cpu().numpy()
    assert np.array_equal(c[0], 2 * ai) and np.array_equal(c[1] + c[2] - c[0], 2 * au) and np.array_equal(c[2], 2 * at)


def test_evaluate_loop_matches_reference():
    """evaluate(): mIoU / per-class IoU over a two-batch loader with the stand-in model, against the reference arithmetic"""
    from oracle import semivl_oracle as O
    from semivl_b200 import evaluate as E
    nclass = 5
    cfg = dict(nclass=nclass, crop_size=24, stride=16)
    g = torch.Generator().manual_seed(9)
    data = []
    for _ in range(2):
        mask = torch.randint(0, nclass, (2, 40, 44), generator=g)
        mask[:, :5] = 255
        data.append((torch.randn(2, 3, 40, 44, generator=g), mask, ["a", "b"]))
    stub = _Stub(nclass)
    inter = np.zeros(nclass)
    union = np.zeros(nclass)
    for img, mask, _ in data:
        p, _ = O.predict_reference(stub, img, mask, "zegclip_sliding_window", cfg)
        ai, au, _ = O.intersection_and_union(p.numpy(), mask.numpy(), nclass, 255)
        inter += ai
        union += au
    want = inter / (union + 1e-10) * 100.0
    miou, iou = E.evaluate(stub.cuda(), data, "zegclip_sliding_window", cfg)
    assert np.allclose(iou, want, atol=0.05) and abs(miou - want.mean()) < 0.05          # a label may flip on a near tie


def test_predict_and_iou_match_reference_golden(golden_dir):
    """CUDA evaluation path against outputs of the UNMODIFIED reference's `predict` + `intersectionAndUnion` (eval_predict_iou.npz):
    stitched scores within fp32 rounding, labels equal wherever the reference's top-2 margin is decidable, histograms bit-exact
    when fed the reference's own labels."""
    import os
    from oracle import semivl_oracle as O
    from semivl_b200 import evaluate as E
    g = dict(np.load(os.path.join(golden_dir, "eval_predict_iou.npz"), allow_pickle=False))
    nclass = int(g["nclass"])
    model = O.StubSegModel(nclass, weight=g["weight"]).cuda()
    for i, case in enumerate(g["cases"]):
        mode, h, w, crop, stride = str(case).split("|")
        cfg = dict(nclass=nclass, crop_size=int(crop), stride=float(stride) if "." in stride else int(stride))
        img, mask = torch.from_numpy(g[f"img{i}"]), torch.from_numpy(g[f"mask{i}"].astype(np.int64))
        pred, final = E.predict(model, img.cuda(), mask, mode, cfg, return_logits=True)
        ref_final = torch.from_numpy(g[f"final{i}"])
        err = (final.cpu() - ref_final).abs().max().item()
        assert err < 2e-5 * max(1.0, ref_final.abs().max().item()), (mode, err)
        top2 = ref_final.topk(2, dim=1).values
        decidable = (top2[:, 0] - top2[:, 1]) > 4 * err
        ref_pred = torch.from_numpy(g[f"pred{i}"].astype(np.int64))
        assert torch.equal(pred.cpu()[decidable], ref_pred[decidable]), mode
        target = E.crop_mask_for(mode, mask, cfg)
        counts = E.intersection_union_counts(ref_pred.cuda(), target.cuda(), nclass, 255).cpu().numpy()
        want = g[f"iou{i}"]
        assert np.array_equal(counts[0], want[0]) and np.array_equal(counts[1] + counts[2] - counts[0], want[1]) and np.array_equal(counts[2], want[2])


def test_zegclip_sliding_window_resizes_to_the_label_size():
    """supervised.py:95-100: when the label is not the image's size the averaged window logits are resized with align_corners=True before the
    arg-max (svl_resize_bilinear_ac); against the oracle's restatement of `predict` and against ATen for the resize itself."""
    from oracle import semivl_oracle as O
    from semivl_b200 import evaluate as E
    from semivl_b200 import lib as L
    nclass = 5
    cfg = dict(nclass=nclass, crop_size=24, stride=16)
    g = torch.Generator().manual_seed(11)
    img = torch.randn(2, 3, 40, 44, generator=g)
    mask = torch.randint(0, nclass, (2, 61, 53), generator=g)
    stub = _Stub(nclass)
    want, logits = O.predict_reference(stub, img, mask, "zegclip_sliding_window", cfg)
    pred, final = E.predict(stub.cuda(), img.cuda(), mask, "zegclip_sliding_window", cfg, return_logits=True)
    assert tuple(final.shape[-2:]) == (61, 53)
    assert (final.cpu() - logits).abs().max() < 1e-4 * logits.abs().max()
    top2 = logits.topk(2, dim=1).values
    decidable = (top2[:, 0] - top2[:, 1]) > 1e-3 * logits.abs().max()
    assert bool((pred.cpu()[decidable] == want[decidable]).all())
    # the kernel alone, up- and down-scaling, degenerate sizes
    for (hs, ws, H, W) in ((7, 9, 20, 31), (20, 31, 7, 9), (5, 1, 9, 4), (1, 6, 1, 11)):
        src = torch.randn(3, hs, ws, generator=g).cuda()
        out = torch.empty(3, H, W, device="cuda")
        L.call("svl_resize_bilinear_ac", src, out, 3, hs, ws, H, W)
        ref = F.interpolate(src[None], size=(H, W), mode="bilinear", align_corners=True)[0]
        assert (out - ref).abs().max() < 1e-5
