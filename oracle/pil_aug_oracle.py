"""TEST INFRASTRUCTURE ONLY.  Integer-exact numpy restatement of the PIL / torchvision image operations the reference's SemiDataset applies
to the unlabelled stream (third_party/unimatch/dataset/transform.py:43-64 `resize`, `blur`; third_party/unimatch/dataset/semi.py:84-93
ColorJitter(0.5, 0.5, 0.5, 0.25), RandomGrayscale, GaussianBlur) -- the checker of semivl_b200/csrc/augment.cu.

The arithmetic lives in third-party packages that are not vendored under /root/reference: Pillow (Resample.c two-pass fixed-point resampling
with 22-bit coefficients, ImagingScaleAffine for NEAREST, Blend.c, Convert.c rgb2l / rgb2hsv / hsv2rgb, BoxBlur.c extended box blur x 3 for
GaussianBlur) and torchvision.transforms (ColorJitter = ImageEnhance.Brightness / Contrast / Color + hue shift in HSV, in a random order).
PINNING: every function here is checked bit for bit against the Pillow / torchvision installed in the image (12.2.0 / 0.26.0) on random
images and parameter sweeps by tests/test_augment_cpu.py; the reference pins Pillow only loosely (requirements.txt), so the pin is to the
installed versions.
"""
import math

import numpy as np

PRECISION_BITS = 22          # Resample.c: 32 - 8 - 2


# ----------------------------------------------------------------------------- Image.resize(..., BILINEAR)
def bilinear_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the triangle filter (support 1, widened by the scale when shrinking):
    returns (kk int32 [out, ksize], bounds int32 [out, 2] = (first source index, number of taps))."""
    scale = in_size / out_size
    fs = max(scale, 1.0)
    support = 1.0 * fs
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), np.int32)
    bounds = np.zeros((out_size, 2), np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / fs
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = []
        ww = 0.0
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


def _resample_axis(img, out_size, axis):
    kk, b = bilinear_coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.zeros((out_size,) + src.shape[1:], np.int64)
    for xx in range(out_size):
        xmin, n = b[xx]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(n):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def resize_bilinear(img, ow, oh):
    """uint8 [h, w, c] -> [oh, ow, c]: horizontal pass, then vertical pass on the uint8 intermediate (ImagingResample)."""
    h, w = img.shape[:2]
    if ow != w:
        img = _resample_axis(img, ow, 1)
    if oh != h:
        img = _resample_axis(img, oh, 0)
    return img


# ----------------------------------------------------------------------------- Image.resize(..., NEAREST)
def nearest_indices(in_size, out_size):
    """ImagingScaleAffine: source index = (int) of a double that starts at scale / 2 and is INCREMENTED by the scale per output pixel."""
    a = in_size / out_size
    idx = np.zeros(out_size, np.int32)
    o = a * 0.5
    for x in range(out_size):
        idx[x] = min(max(int(o) if o >= 0 else 0, 0), in_size - 1)
        o += a
    return idx


def resize_nearest(mask, ow, oh):
    h, w = mask.shape
    return mask[nearest_indices(h, oh)][:, nearest_indices(w, ow)]


def resized_size(w, h, long_side):
    """transform.resize (transform.py:43-57): the long side becomes `long_side`, the short one is rounded half up."""
    if h > w:
        return int(1.0 * w * long_side / h + 0.5), long_side
    return long_side, int(1.0 * h * long_side / w + 0.5)


# ----------------------------------------------------------------------------- ImageEnhance / ColorJitter pieces
def to_gray(img):
    """Image.convert('L') (Convert.c rgb2l)."""
    r, g, b = (img[..., i].astype(np.int64) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend(a, b, alpha):
    """Image.blend (Blend.c): a + alpha * (b - a) in single precision, truncated; clipped to [0, 255] when alpha is outside [0, 1]."""
    t = a.astype(np.float32) + np.float32(alpha) * (b.astype(np.int32) - a.astype(np.int32)).astype(np.float32)
    if 0.0 <= alpha <= 1.0:
        return t.astype(np.int32).astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t.astype(np.int32))).astype(np.uint8)


def adjust_brightness(img, f):
    return blend(np.zeros_like(img), img, f)


def gray_mean(img):
    g = to_gray(img)
    return int(int(g.astype(np.int64).sum()) / g.size + 0.5)


def adjust_contrast(img, f):
    return blend(np.full_like(img, gray_mean(img)), img, f)


def adjust_saturation(img, f):
    return blend(np.stack([to_gray(img)] * 3, -1), img, f)


def rgb_to_grayscale3(img):
    """RandomGrayscale: F.rgb_to_grayscale(img, num_output_channels=3)."""
    return np.stack([to_gray(img)] * 3, -1)


def rgb2hsv(img):
    """Convert.c rgb2hsv_row: float variables, double literals -- the hue expression and the fmod are evaluated in double and ROUNDED to float
    when they are stored, h * 255.0 is a double product truncated to int."""
    f32, f64 = np.float32, np.float64
    r, g, b = (img[..., i].astype(np.int32) for i in range(3))
    maxc = np.maximum(r, np.maximum(g, b))
    minc = np.minimum(r, np.minimum(g, b))
    cr = (maxc - minc).astype(f32)
    safe = np.where(cr == 0, 1, cr).astype(f32)
    s = (cr / np.where(maxc == 0, 1, maxc).astype(f32)).astype(f32)
    rc, gc, bc = (((maxc - c).astype(f32) / safe).astype(f32) for c in (r, g, b))
    h = np.where(r == maxc, bc.astype(f64) - gc.astype(f64),
                 np.where(g == maxc, 2.0 + rc.astype(f64) - bc.astype(f64), 4.0 + gc.astype(f64) - rc.astype(f64))).astype(f32)
    h = np.fmod(h.astype(f64) / 6.0 + 1.0, 1.0).astype(f32)
    uh = np.clip((h.astype(f64) * 255.0).astype(np.int32), 0, 255)
    us = np.clip((s.astype(f64) * 255.0).astype(np.int32), 0, 255)
    gray = maxc == minc
    return np.stack([np.where(gray, 0, uh), np.where(gray, 0, us), maxc], -1).astype(np.uint8)


def hsv2rgb(hsv):
    """Convert.c hsv2rgb: single precision, p / q / t rounded half up."""
    f32 = np.float32
    h, s, v = (hsv[..., i].astype(np.int32) for i in range(3))
    fs = s.astype(f32) / f32(255.0)
    hf = h.astype(f32) * f32(6.0) / f32(255.0)
    i = np.floor(hf).astype(np.int32)
    f = hf - i.astype(f32)
    vf = v.astype(f32)
    rnd = lambda x: np.floor(x + f32(0.5)).astype(np.int32)
    p, q, t = rnd(vf * (f32(1.0) - fs)), rnd(vf * (f32(1.0) - fs * f)), rnd(vf * (f32(1.0) - fs * (f32(1.0) - f)))
    i6 = i % 6
    out = np.stack([np.choose(i6, [v, q, p, p, t, v]), np.choose(i6, [t, v, v, q, p, p]), np.choose(i6, [p, p, t, v, v, q])], -1)
    out = np.where((s == 0)[..., None], np.stack([v, v, v], -1), out)
    return np.clip(out, 0, 255).astype(np.uint8)


def hue_shift_u8(f):
    """torchvision F_pil.adjust_hue: the H channel is shifted by np.uint8(hue_factor * 255) with uint8 wrap-around."""
    return int(np.array(f * 255).astype(np.uint8))


def adjust_hue(img, f):
    hsv = rgb2hsv(img)
    hsv[..., 0] = (hsv[..., 0].astype(np.int32) + hue_shift_u8(f)) % 256
    return hsv2rgb(hsv)


COLOR_OPS = (adjust_brightness, adjust_contrast, adjust_saturation, adjust_hue)      # torchvision ColorJitter's fn_id 0..3


def color_jitter(img, order, factors):
    for fn_id in order:
        img = COLOR_OPS[int(fn_id)](img, float(factors[int(fn_id)]))
    return img


# ----------------------------------------------------------------------------- ImageFilter.GaussianBlur
def gaussian_box_radius(sigma, passes=3):
    """BoxBlur.c _gaussian_blur_radius (single precision): the fractional box radius whose `passes`-fold application has variance sigma^2."""
    f32 = np.float32
    sigma = f32(sigma)
    sigma2 = f32(sigma * sigma / f32(passes))
    L = f32(math.sqrt(12.0 * float(sigma2) + 1.0))
    l = f32(math.floor((float(L) - 1.0) / 2.0))
    a = f32((2 * l + 1) * (l * (l + 1) - 3 * sigma2))
    a = f32(a / f32(6 * (sigma2 - (l + 1) * (l + 1))))
    return f32(l + a)


def box_weights(fr):
    """(integer radius, ww, fw) of ImagingHorizontalBoxBlur: 8.24 fixed-point weights of the inner pixels and of the two far neighbours."""
    radius = int(fr)
    ww = int(np.float32(1 << 24) / np.float32(np.float32(fr) * np.float32(2) + np.float32(1)))
    fw = ((1 << 24) - (radius * 2 + 1) * ww) // 2
    return radius, ww, fw


def box_blur_pass(img, axis, radius, ww, fw):
    """One extended box blur along `axis` with edge replication: out = ((window sum) * ww + (two far neighbours) * fw + 2^23) >> 24, all in
    uint32 arithmetic (ImagingLineBoxBlur8; the C code keeps a running window sum, which equals the direct sum modulo 2^32)."""
    src = np.moveaxis(img, axis, 0).astype(np.uint64)
    n = src.shape[0]
    idx = np.arange(n)
    M = np.uint64(0xFFFFFFFF)
    acc = np.zeros_like(src)
    for d in range(-radius, radius + 1):
        acc = (acc + src[np.clip(idx + d, 0, n - 1)]) & M
    far = src[np.clip(idx - radius - 1, 0, n - 1)] + src[np.clip(idx + radius + 1, 0, n - 1)]
    bulk = (acc * np.uint64(ww) + far * np.uint64(fw)) & M
    out = ((bulk + np.uint64(1 << 23)) & M) >> np.uint64(24)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def gaussian_blur(img, sigma, passes=3):
    """img.filter(ImageFilter.GaussianBlur(radius=sigma)): `passes` horizontal box blurs, then `passes` vertical ones."""
    radius, ww, fw = box_weights(gaussian_box_radius(sigma, passes))
    for axis in (1, 0):
        for _ in range(passes):
            img = box_blur_pass(img, axis, radius, ww, fw)
    return img
