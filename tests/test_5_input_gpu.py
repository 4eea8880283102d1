"""Input-stage kernels (csrc/input.cu) against outputs of the UNMODIFIED reference's transform.py (tests/golden/input_stage.npz):
bit-exact floats (same IEEE operation order as torchvision's ToTensor + Normalize) and exact integer labels / masks / boxes."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_crop_flip_normalize_bit_exact(golden_dir):
    from semivl_b200 import input_pipeline as P
    g = dict(np.load(os.path.join(golden_dir, "input_stage.npz"), allow_pickle=False))
    for i, case in enumerate(g["cases"]):
        h, w, size, pad = (int(v) for v in str(case).split("|"))
        random.seed(500 + i)
        x0, y0 = P.sample_crop(w, h, size)
        flip = P.sample_hflip(0.5)
        img = torch.from_numpy(g[f"img{i}"]).cuda()
        mask = torch.from_numpy(g[f"mask{i}"]).cuda()
        t = P.crop_flip_normalize(img, size, x0, y0, flip)
        lab, ign = P.crop_flip_mask(mask, size, x0, y0, flip, ignore_value=pad, want_ignore_mask=True)
        assert torch.equal(t.cpu(), torch.from_numpy(g[f"out_img{i}"])), case            # bit-exact fp32
        assert torch.equal(lab.cpu(), torch.from_numpy(g[f"out_mask{i}"].astype(np.int64))), case
        assert torch.equal(ign.cpu(), torch.from_numpy(g[f"out_ign{i}"].astype(np.int64))), case
        _, ign_only = P.crop_flip_mask(mask, size, x0, y0, flip, ignore_value=pad, want_labels=False, want_ignore_mask=True)
        assert torch.equal(ign_only, ign)


def test_cutmix_boxes_match_reference(golden_dir):
    from semivl_b200 import input_pipeline as P
    g = dict(np.load(os.path.join(golden_dir, "input_stage.npz"), allow_pickle=False))
    for j, want in enumerate(g["boxes"]):
        random.seed(900 + j)
        np.random.seed(900 + j)
        box = P.cutmix_box(48, P.sample_cutmix_box(48, p=0.5))
        assert torch.equal(box.cpu(), torch.from_numpy(want).float()), j


def test_full_size_crop_properties():
    """512 x 512 crop of a larger image: flip twice = identity crop; a crop fully inside the source equals the torch slice"""
    from semivl_b200 import input_pipeline as P
    gen = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (700, 900, 3), generator=gen, dtype=torch.uint8).cuda()
    a = P.crop_flip_normalize(img, 512, 123, 77, False)
    b = P.crop_flip_normalize(img, 512, 123, 77, True)
    assert torch.equal(a, b.flip(-1))
    # the reference normalises on the CPU (DataLoader workers); torch's CUDA `div(255)` multiplies by a reciprocal and is not the oracle
    mean = torch.tensor(P.MEAN).view(3, 1, 1)
    std = torch.tensor(P.STD).view(3, 1, 1)
    ref = img.cpu()[77:77 + 512, 123:123 + 512].permute(2, 0, 1).float().div(255).sub(mean).div(std)
    assert torch.equal(a.cpu(), ref)
