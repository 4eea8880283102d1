"""Input stage of the reference's SemiDataset on the GPU (third_party/unimatch/dataset/transform.py:9-41,66-84;
third_party/unimatch/dataset/semi.py:76-107): pad + random crop + horizontal flip + ToTensor/Normalize of the image, the label
/ ignore-mask conversion, and the CutMix box.  The random draws stay on the host and follow the reference's call order
(python `random` for crop / flip / box probability, `numpy.random` for the box geometry), so a seeded run picks the same
crops; the per-pixel work runs in svl_* kernels on uint8 sources already resident on the device.

The scale and strong augmentations of the unlabelled stream -- random rescale (transform.py:43-57), ColorJitter(0.5, 0.5, 0.5, 0.25) with
p = 0.8, RandomGrayscale(p = 0.2), GaussianBlur with p = 0.5 (semi.py:84-93, transform.py:59-64) -- run on the GPU as integer-exact
counterparts of the Pillow / torchvision code paths (csrc/augment.cu); their random parameters are drawn on the host from the same three
generators in the same order as the reference (python `random`, `numpy.random`, torch's global generator), and the small coefficient /
index tables of the resize are built on the host exactly as Pillow builds them.  `unlabeled_sample` chains the whole
`SemiDataset.__getitem__` of mode 'train_u' (semi.py:63-107) on a decoded uint8 image."""
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


# ---------------------------------------------------------------------------------------------- scale / strong augmentations (csrc/augment.cu)
PRECISION_BITS = 22          # Pillow Resample.c: 32 - 8 - 2


def sample_resize(w, h, ratio_range=(0.5, 2.0)):
    """transform.resize's draw (transform.py:44-46) and the resulting size (long side drawn, short side rounded half up); returns (ow, oh)"""
    long_side = random.randint(int(max(h, w) * ratio_range[0]), int(max(h, w) * ratio_range[1]))
    if h > w:
        return int(1.0 * w * long_side / h + 0.5), long_side
    return long_side, int(1.0 * h * long_side / w + 0.5)


def sample_strong_view():
    """The draws of one strong view (semi.py:84-88 / 90-94) in the reference's order: `random.random() < 0.8` -> ColorJitter.get_params
    (torch.randperm(4), then brightness / contrast / saturation / hue factors from torch's global generator), RandomGrayscale's
    `torch.rand(1) < 0.2`, transform.blur's `random.random() < 0.5` and `np.random.uniform(0.1, 2.0)`."""
    jitter = None
    if random.random() < 0.8:
        order = torch.randperm(4).tolist()
        factors = [float(torch.empty(1).uniform_(0.5, 1.5)), float(torch.empty(1).uniform_(0.5, 1.5)), float(torch.empty(1).uniform_(0.5, 1.5)),
                   float(torch.empty(1).uniform_(-0.25, 0.25))]
        jitter = (order, factors)
    gray = bool(torch.rand(1) < 0.2)
    sigma = None
    if random.random() < 0.5:
        sigma = float(np.random.uniform(0.1, 2.0))
    return dict(jitter=jitter, gray=gray, blur_sigma=sigma)


def bilinear_tables(in_size, out_size):
    """Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc for the triangle filter (support 1, widened by the scale when shrinking):
    (kk int32 [out, ksize] 22-bit fixed point, bounds int32 [out, 2] = (first source index, taps)).  Double arithmetic, as in C."""
    import math
    scale = in_size / out_size
    fs = max(scale, 1.0)
    support = 1.0 * fs
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), np.int32)
    bounds = np.zeros((out_size, 2), np.int32)
    ss = 1.0 / fs
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w, ww = [], 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            v = 1.0 - a if a < 1.0 else 0.0
            w.append(v)
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return kk, bounds


def nearest_table(in_size, out_size):
    """Pillow ImagingScaleAffine: source index = (int) of a double that starts at scale / 2 and is incremented by the scale per output pixel"""
    a = in_size / out_size
    idx = np.zeros(out_size, np.int32)
    o = a * 0.5
    for x in range(out_size):
        idx[x] = min(max(int(o), 0), in_size - 1)
        o += a
    return idx


def resize_bilinear(img_u8, ow, oh):
    """uint8 [h, w, 3] on the GPU -> [oh, ow, 3] = PIL Image.resize((ow, oh), BILINEAR): horizontal pass, then vertical pass (transform.py:55)"""
    assert img_u8.dtype == torch.uint8 and img_u8.dim() == 3 and img_u8.is_contiguous()
    h, w, c = img_u8.shape
    dev = img_u8.device
    cur = img_u8
    if ow != w:
        kk, b = bilinear_tables(w, ow)
        tmp = torch.empty(h, ow, c, device=dev, dtype=torch.uint8)
        L.call("svl_resample_pass_u8", cur, tmp, torch.from_numpy(kk).to(dev), torch.from_numpy(b).to(dev), kk.shape[1], h, w, c, ow, 1)
        cur, w = tmp, ow
    if oh != h:
        kk, b = bilinear_tables(h, oh)
        out = torch.empty(oh, w, c, device=dev, dtype=torch.uint8)
        L.call("svl_resample_pass_u8", cur, out, torch.from_numpy(kk).to(dev), torch.from_numpy(b).to(dev), kk.shape[1], h, w, c, oh, 0)
        cur = out
    return cur


def resize_nearest(mask_u8, ow, oh):
    """uint8 [h, w] -> [oh, ow] = PIL Image.resize((ow, oh), NEAREST) (transform.py:56)"""
    assert mask_u8.dtype == torch.uint8 and mask_u8.dim() == 2 and mask_u8.is_contiguous()
    h, w = mask_u8.shape
    dev = mask_u8.device
    out = torch.empty(oh, ow, device=dev, dtype=torch.uint8)
    L.call("svl_gather_nearest_u8", mask_u8, out, torch.from_numpy(nearest_table(h, oh)).to(dev), torch.from_numpy(nearest_table(w, ow)).to(dev), w, oh, ow)
    return out


def color_jitter_(img_u8, order, factors):
    """in place: torchvision ColorJitter's four operations in the drawn order (fn_id 0 brightness, 1 contrast, 2 saturation, 3 hue)"""
    npix = img_u8.shape[0] * img_u8.shape[1]
    scratch = torch.empty(1, device=img_u8.device, dtype=torch.int64)
    for fn_id in order:
        f = float(factors[fn_id])
        if fn_id == 3:
            L.call("svl_color_op_u8", img_u8, npix, 4, 0.0, int(np.array(f * 255).astype(np.uint8)), scratch)
        else:
            L.call("svl_color_op_u8", img_u8, npix, int(fn_id), f, 0, scratch, n_launch=2 if fn_id == 1 else 1)
    return img_u8


def grayscale_(img_u8):
    L.call("svl_color_op_u8", img_u8, img_u8.shape[0] * img_u8.shape[1], 3, 0.0, 0, None)
    return img_u8


def gaussian_box_radius(sigma, passes=3):
    """Pillow BoxBlur.c _gaussian_blur_radius (single precision)"""
    import math
    f32 = np.float32
    sigma = f32(sigma)
    sigma2 = f32(sigma * sigma / f32(passes))
    big_l = f32(math.sqrt(12.0 * float(sigma2) + 1.0))
    l = f32(math.floor((float(big_l) - 1.0) / 2.0))
    a = f32((2 * l + 1) * (l * (l + 1) - 3 * sigma2))
    a = f32(a / f32(6 * (sigma2 - (l + 1) * (l + 1))))
    return f32(l + a)


def gaussian_blur(img_u8, sigma, passes=3):
    """PIL img.filter(ImageFilter.GaussianBlur(radius=sigma)) (transform.py:59-64): `passes` box blurs along x, then along y"""
    fr = gaussian_box_radius(sigma, passes)
    radius = int(fr)
    ww = int(np.float32(1 << 24) / np.float32(np.float32(fr) * np.float32(2) + np.float32(1)))
    fw = ((1 << 24) - (radius * 2 + 1) * ww) // 2
    h, w, c = img_u8.shape
    a, b = img_u8, torch.empty_like(img_u8)
    for axis in (1, 0):
        for _ in range(passes):
            L.call("svl_box_blur_pass_u8", a, b, h, w, c, axis, radius, ww, fw)
            a, b = b, (a if a is not img_u8 else torch.empty_like(img_u8))
    return a


def strong_view(img_u8, params):
    """one strong view of semi.py:84-88 from `sample_strong_view()`'s draws; the input is left untouched"""
    out = img_u8.clone()
    if params["jitter"] is not None:
        color_jitter_(out, *params["jitter"])
    if params["gray"]:
        grayscale_(out)
    if params["blur_sigma"] is not None:
        out = gaussian_blur(out, params["blur_sigma"])
    return out


def unlabeled_sample(img_u8, mask_u8, size, ratio_range=(0.5, 2.0)):
    """`SemiDataset.__getitem__` of mode 'train_u' (semi.py:63-107) on a decoded uint8 image [h, w, 3] / label map [h, w] resident on the
    GPU: random rescale, pad + random crop (padding value 254), horizontal flip, two strong views, CutMix boxes, normalisation.  Returns
    (img_w, img_s1, img_s2 f32 [3, size, size], ignore_mask int64 [size, size], cutmix_box1, cutmix_box2 f32 [size, size]) -- the reference's
    tuple -- with every random draw taken from the same generators in the same order."""
    h, w = mask_u8.shape
    ow, oh = sample_resize(w, h, ratio_range)
    img = resize_bilinear(img_u8, ow, oh)
    mask = resize_nearest(mask_u8, ow, oh)
    x0, y0 = sample_crop(ow, oh, size)
    flip = sample_hflip()
    # crop + flip once on the uint8 image (the strong views start from the cropped, flipped weak view)
    weak_u8 = crop_flip_u8(img, size, x0, y0, flip)
    p1 = sample_strong_view()
    box1 = sample_cutmix_box(size)
    p2 = sample_strong_view()
    box2 = sample_cutmix_box(size)
    s1, s2 = strong_view(weak_u8, p1), strong_view(weak_u8, p2)
    _, ign = crop_flip_mask(mask, size, x0, y0, flip, ignore_value=254, want_labels=False, want_ignore_mask=True)
    dev = img_u8.device
    return (crop_flip_normalize(weak_u8, size, 0, 0, False), crop_flip_normalize(s1, size, 0, 0, False), crop_flip_normalize(s2, size, 0, 0, False),
            ign, cutmix_box(size, box1, dev), cutmix_box(size, box2, dev))


def crop_flip_u8(img_u8, size, x0, y0, flip):
    """uint8 [h, w, 3] -> uint8 [size, size, 3]: zero pad to at least `size`, crop at (x0, y0), optional horizontal flip (transform.py:9-28);
    plain device indexing (a byte copy: no arithmetic to get wrong)"""
    h, w, c = img_u8.shape
    ph, pw = max(h, size), max(w, size)
    if (ph, pw) != (h, w):
        padded = torch.zeros(ph, pw, c, device=img_u8.device, dtype=torch.uint8)
        padded[:h, :w] = img_u8
        img_u8 = padded
    out = img_u8[y0:y0 + size, x0:x0 + size]
    if flip:
        out = out.flip(1)
    return out.contiguous()
