"""Input stage of the reference's SemiDataset on the GPU (third_party/unimatch/dataset/transform.py:9-41,66-84;
third_party/unimatch/dataset/semi.py:76-107): pad + random crop + horizontal flip + ToTensor/Normalize of the image, the label
/ ignore-mask conversion, and the CutMix box.  The random draws stay on the host and follow the reference's call order
(python `random` for crop / flip / box probability, `numpy.random` for the box geometry), so a seeded run picks the same
crops; the per-pixel work runs in svl_* kernels on uint8 sources already resident on the device.

Not covered here (PIL-specific filters, kept on the host): the random rescale (transform.py:44-57), ColorJitter / RandomGrayscale /
GaussianBlur of the strong views (semi.py:84-93)."""
import ctypes as C
import random

import numpy as np
import torch

from . import lib as L

MEAN = (0.485, 0.456, 0.406)          # transform.py:37
STD = (0.229, 0.224, 0.225)
_MEAN3 = (C.c_float * 3)(*MEAN)
_STD3 = (C.c_float * 3)(*STD)


def sample_crop(w, h, size):
    """the two draws of transform.crop (transform.py:16-18) on the padded size; returns (x0, y0)"""
    pw, ph = max(w, size), max(h, size)
    x = random.randint(0, pw - size)
    y = random.randint(0, ph - size)
    return x, y


def sample_hflip(p=0.5):
    """transform.hflip's draw (transform.py:27)"""
    return random.random() < p


def sample_cutmix_box(img_size, p=0.5, size_min=0.02, size_max=0.4, ratio_1=0.3, ratio_2=1 / 0.3):
    """transform.obtain_cutmix_box's draws (transform.py:66-84); returns (x, y, w, h) or None for the empty box"""
    if random.random() > p:
        return None
    size = np.random.uniform(size_min, size_max) * img_size * img_size
    while True:
        ratio = np.random.uniform(ratio_1, ratio_2)
        cw = int(np.sqrt(size / ratio))
        ch = int(np.sqrt(size * ratio))
        x = np.random.randint(0, img_size)
        y = np.random.randint(0, img_size)
        if x + cw <= img_size and y + ch <= img_size:
            return x, y, cw, ch


def crop_flip_normalize(img_u8, size, x0, y0, flip, out=None):
    """img_u8: uint8 [h, w, 3] on the GPU -> f32 [3, size, size] = normalize(hflip(crop(pad(img))))   (transform.py:9-41)"""
    assert img_u8.dtype == torch.uint8 and img_u8.dim() == 3 and img_u8.shape[2] == 3 and img_u8.is_contiguous()
    h, w = img_u8.shape[:2]
    if out is None:
        out = torch.empty(3, size, size, device=img_u8.device, dtype=torch.float32)
    L.call("svl_crop_flip_normalize", img_u8, h, w, out, size, x0, y0, 1 if flip else 0, _MEAN3, _STD3)
    return out


def crop_flip_mask(mask_u8, size, x0, y0, flip, ignore_value=255, want_labels=True, want_ignore_mask=False):
    """mask_u8: uint8 [h, w] -> (int64 labels [size, size] padded with `ignore_value`, int64 ignore mask: 255 where the label is 254)
    (transform.py:9-24, semi.py:74,99-103: unlabelled samples pad with 254 and only keep the ignore mask)"""
    assert mask_u8.dtype == torch.uint8 and mask_u8.dim() == 2 and mask_u8.is_contiguous()
    h, w = mask_u8.shape
    lab = torch.empty(size, size, device=mask_u8.device, dtype=torch.int64) if want_labels else None
    ign = torch.empty(size, size, device=mask_u8.device, dtype=torch.int64) if want_ignore_mask else None
    L.call("svl_crop_flip_mask", mask_u8, h, w, lab, ign, size, x0, y0, 1 if flip else 0, ignore_value)
    return lab, ign


def cutmix_box(img_size, params, device="cuda", out=None):
    """f32 [img_size, img_size] box mask from sample_cutmix_box's result (None -> all zeros)"""
    if out is None:
        out = torch.empty(img_size, img_size, device=device, dtype=torch.float32)
    x, y, w, h = params if params is not None else (0, 0, 0, 0)
    L.call("svl_cutmix_box", out, img_size, x, y, w, h)
    return out
