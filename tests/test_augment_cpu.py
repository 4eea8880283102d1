"""Pins oracle/pil_aug_oracle.py (the integer-exact restatement of the PIL / torchvision operations of the reference's unlabelled-stream
augmentations) bit for bit against the Pillow / torchvision installed in the image, and the host-side parameter draws of
semivl_b200.input_pipeline against the reference's call order (third_party/unimatch/dataset/semi.py:63-93, transform.py:43-64)."""
import random

import numpy as np
import pytest
import torch

Image = pytest.importorskip("PIL.Image")
from PIL import ImageFilter  # noqa: E402

TF = pytest.importorskip("torchvision.transforms.functional")
from oracle import pil_aug_oracle as A  # noqa: E402


def _img(h, w, seed):
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    img[: h // 8, : w // 8] = 0                     # flat regions: gray pixels (hue 0), black, white
    img[h // 8: h // 4, : w // 8] = 255
    img[h // 4: h // 3, : w // 8] = (50, 50, 50)
    return img


@pytest.mark.parametrize("h,w,oh,ow", [(40, 56, 61, 85), (375, 500, 240, 320), (64, 64, 33, 97), (100, 80, 250, 200), (500, 375, 187, 140), (64, 64, 64, 31)])
def test_resize_matches_pil(h, w, oh, ow):
    img = _img(h, w, h + w)
    assert np.array_equal(A.resize_bilinear(img, ow, oh), np.array(Image.fromarray(img).resize((ow, oh), Image.BILINEAR)))
    mask = np.random.RandomState(w).randint(0, 255, (h, w)).astype(np.uint8)
    assert np.array_equal(A.resize_nearest(mask, ow, oh), np.array(Image.fromarray(mask).resize((ow, oh), Image.NEAREST)))


@pytest.mark.parametrize("f", [0.5, 0.77, 1.0, 1.31, 1.5])
def test_color_ops_match_torchvision(f):
    img = _img(97, 131, 3)
    pil = Image.fromarray(img)
    assert np.array_equal(A.adjust_brightness(img, f), np.array(TF.adjust_brightness(pil, f)))
    assert np.array_equal(A.adjust_contrast(img, f), np.array(TF.adjust_contrast(pil, f)))
    assert np.array_equal(A.adjust_saturation(img, f), np.array(TF.adjust_saturation(pil, f)))
    h = (f - 1.0) * 0.5                              # [-0.25, 0.25]
    assert np.array_equal(A.adjust_hue(img, h), np.array(TF.adjust_hue(pil, h)))
    assert np.array_equal(A.rgb_to_grayscale3(img), np.array(TF.rgb_to_grayscale(pil, 3)))
    assert np.array_equal(A.rgb2hsv(img), np.array(pil.convert("HSV")))


@pytest.mark.parametrize("sigma", [0.1, 0.37, 0.9, 1.3, 1.77, 2.0, 3.4])
def test_gaussian_blur_matches_pil(sigma):
    img = _img(37, 53, 5)
    assert np.array_equal(A.gaussian_blur(img, sigma), np.array(Image.fromarray(img).filter(ImageFilter.GaussianBlur(radius=sigma))))


def test_strong_view_draws_follow_the_reference_order():
    """One strong view of semi.py:84-88 driven on PIL with seeded `random` / `numpy.random` / torch RNGs against the host-side draws of
    semivl_b200.input_pipeline.sample_strong_view + the oracle's operations."""
    from torchvision import transforms
    from semivl_b200 import input_pipeline as ip
    img = _img(64, 80, 7)
    for seed in range(12):
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        pil = Image.fromarray(img)
        if random.random() < 0.8:
            pil = transforms.ColorJitter(0.5, 0.5, 0.5, 0.25)(pil)
        pil = transforms.RandomGrayscale(p=0.2)(pil)
        if random.random() < 0.5:                    # transform.blur
            pil = pil.filter(ImageFilter.GaussianBlur(radius=np.random.uniform(0.1, 2.0)))
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        params = ip.sample_strong_view()
        out = img
        if params["jitter"] is not None:
            out = A.color_jitter(out, *params["jitter"])
        if params["gray"]:
            out = A.rgb_to_grayscale3(out)
        if params["blur_sigma"] is not None:
            out = A.gaussian_blur(out, params["blur_sigma"])
        assert np.array_equal(out, np.array(pil)), seed
