"""GPU parity of the scale / strong augmentations of the unlabelled stream (csrc/augment.cu + semivl_b200.input_pipeline) -- bit-exact against
the Pillow / torchvision calls the reference's SemiDataset makes (third_party/unimatch/dataset/semi.py:63-107, transform.py:9-84), run live
with the same seeds, and against the numpy oracle (oracle/pil_aug_oracle.py, itself pinned to Pillow by tests/test_augment_cpu.py)."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
Image = pytest.importorskip("PIL.Image")
from PIL import ImageFilter, ImageOps  # noqa: E402

TF = pytest.importorskip("torchvision.transforms.functional")


def _img(h, w, seed):
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    img[: h // 8, : w // 8] = 0
    img[h // 8: h // 4, : w // 8] = 255
    img[h // 4: h // 3, : w // 8] = (50, 50, 50)
    return img


@pytest.mark.parametrize("h,w,oh,ow", [(40, 56, 61, 85), (375, 500, 240, 320), (64, 64, 33, 97), (100, 80, 250, 200), (500, 375, 187, 140),
                                       (366, 500, 750, 1024)])
def test_resize_kernels_match_pil(h, w, oh, ow):
    from semivl_b200 import input_pipeline as ip
    img = _img(h, w, h * w)
    got = ip.resize_bilinear(torch.from_numpy(img).cuda(), ow, oh).cpu().numpy()
    assert np.array_equal(got, np.array(Image.fromarray(img).resize((ow, oh), Image.BILINEAR)))
    mask = np.random.RandomState(w).randint(0, 255, (h, w)).astype(np.uint8)
    gotm = ip.resize_nearest(torch.from_numpy(mask).cuda(), ow, oh).cpu().numpy()
    assert np.array_equal(gotm, np.array(Image.fromarray(mask).resize((ow, oh), Image.NEAREST)))


@pytest.mark.parametrize("f", [0.5, 0.63, 0.77, 1.0, 1.31, 1.5])
def test_color_kernels_match_torchvision(f):
    from semivl_b200 import input_pipeline as ip
    img = _img(97, 131, 3)
    pil = Image.fromarray(img)
    dev = lambda: torch.from_numpy(img.copy()).cuda()
    h = (f - 1.0) * 0.5
    for fn_id, ref in ((0, TF.adjust_brightness(pil, f)), (1, TF.adjust_contrast(pil, f)), (2, TF.adjust_saturation(pil, f)), (3, TF.adjust_hue(pil, h))):
        factors = [f, f, f, h]
        got = ip.color_jitter_(dev(), [fn_id], factors).cpu().numpy()
        assert np.array_equal(got, np.array(ref)), fn_id
    assert np.array_equal(ip.grayscale_(dev()).cpu().numpy(), np.array(TF.rgb_to_grayscale(pil, 3)))


@pytest.mark.parametrize("sigma", [0.1, 0.37, 0.9, 1.3, 1.77, 2.0, 3.4])
def test_gaussian_blur_kernel_matches_pil(sigma):
    from semivl_b200 import input_pipeline as ip
    img = _img(67, 93, 5)
    got = ip.gaussian_blur(torch.from_numpy(img).cuda(), sigma).cpu().numpy()
    assert np.array_equal(got, np.array(Image.fromarray(img).filter(ImageFilter.GaussianBlur(radius=sigma))))


def _reference_train_u(img, mask, size, ratio_range):
    """SemiDataset.__getitem__ (mode 'train_u', img_scale None; semi.py:63-107) restated call for call on PIL / torchvision."""
    from torchvision import transforms
    pil, pm = Image.fromarray(img), Image.fromarray(mask)
    w, h = pil.size                                                   # transform.resize (transform.py:43-57)
    long_side = random.randint(int(max(h, w) * ratio_range[0]), int(max(h, w) * ratio_range[1]))
    if h > w:
        oh, ow = long_side, int(1.0 * w * long_side / h + 0.5)
    else:
        ow, oh = long_side, int(1.0 * h * long_side / w + 0.5)
    pil, pm = pil.resize((ow, oh), Image.BILINEAR), pm.resize((ow, oh), Image.NEAREST)
    w, h = pil.size                                                   # transform.crop (transform.py:9-24), ignore_value 254
    padw, padh = (size - w if w < size else 0), (size - h if h < size else 0)
    pil, pm = ImageOps.expand(pil, border=(0, 0, padw, padh), fill=0), ImageOps.expand(pm, border=(0, 0, padw, padh), fill=254)
    w, h = pil.size
    x, y = random.randint(0, w - size), random.randint(0, h - size)
    pil, pm = pil.crop((x, y, x + size, y + size)), pm.crop((x, y, x + size, y + size))
    if random.random() < 0.5:                                         # transform.hflip
        pil, pm = pil.transpose(Image.FLIP_LEFT_RIGHT), pm.transpose(Image.FLIP_LEFT_RIGHT)
    views, boxes = [], []
    from semivl_b200 import input_pipeline as ip
    for _ in range(2):                                                # semi.py:84-93
        v = pil.copy()
        if random.random() < 0.8:
            v = transforms.ColorJitter(0.5, 0.5, 0.5, 0.25)(v)
        v = transforms.RandomGrayscale(p=0.2)(v)
        if random.random() < 0.5:
            v = v.filter(ImageFilter.GaussianBlur(radius=np.random.uniform(0.1, 2.0)))
        views.append(v)
        boxes.append(ip.sample_cutmix_box(size))                      # obtain_cutmix_box's draws (pinned against the reference in test_host_cpu)
    norm = lambda p_: TF.normalize(TF.to_tensor(p_), [0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
    ign = torch.zeros(size, size, dtype=torch.long)
    ign[torch.from_numpy(np.array(pm)).long() == 254] = 255
    return norm(pil), norm(views[0]), norm(views[1]), ign, boxes


@pytest.mark.parametrize("h,w,size", [(120, 160, 96), (90, 70, 128), (200, 150, 64)])
def test_unlabeled_sample_matches_the_reference_pipeline(h, w, size):
    """The whole 'train_u' sample of the reference (rescale, pad / crop with 254, flip, two strong views, CutMix boxes, normalisation) for
    several seeds: every float of the three views and every entry of the ignore mask / boxes identical."""
    from semivl_b200 import input_pipeline as ip
    img = _img(h, w, h + w + size)
    mask = np.random.RandomState(h).randint(0, 21, (h, w)).astype(np.uint8)
    for seed in range(6):
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        rw, rs1, rs2, rign, rboxes = _reference_train_u(img, mask, size, (0.5, 2.0))
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        gw, gs1, gs2, gign, gb1, gb2 = ip.unlabeled_sample(torch.from_numpy(img).cuda(), torch.from_numpy(mask).cuda(), size, (0.5, 2.0))
        for name, a, b in (("img_w", gw, rw), ("img_s1", gs1, rs1), ("img_s2", gs2, rs2)):
            assert torch.equal(a.cpu(), b), (seed, name, (a.cpu() - b).abs().max().item())
        assert torch.equal(gign.cpu(), rign), seed
        for g, bx in ((gb1, rboxes[0]), (gb2, rboxes[1])):
            want = torch.zeros(size, size)
            if bx is not None:
                x, y, cw, ch = bx
                want[y:y + ch, x:x + cw] = 1
            assert torch.equal(g.cpu(), want), seed
