"""CPU oracle for image resizing as the reference does it: Pillow's Image.resize (ImageNet-S `pil-*` resize types,
RobustART/noise/utils/imagenet_s_gen.py:19-26,120-126,148-160, and the eval transform's Resize, imagenet_dataloader.py:74-80).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/ and nothing else.

The arithmetic lives in a third-party dependency that is not vendored in /root/reference: Pillow (requirements.txt pins
none; the reference environment is Pillow 7-8, this container has 12.2 -- libImaging/Resample.c is unchanged across them
for 8-bit images).  This module restates its published algorithm in numpy:
  * Resample.c  precompute_coeffs / normalize_coeffs_8bpc / ImagingResampleHorizontal_8bpc / Vertical_8bpc / ImagingResample:
    separable filter with support scaled by the shrink factor, coefficients rounded to 22-bit fixed point, horizontal pass
    then vertical pass with a uint8 image in between, only the source rows the vertical pass needs;
  * Geometry.c  ImagingScaleAffine for NEAREST: source index (int)(x0 + a*0.5 + a + a + ...) with the running double sum.
Pinned by tests/test_oracle_cpu.py::test_resize_restatement_equals_pil against Pillow itself (every filter, up / down / mixed
scaling, odd sizes)."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
FILTERS = ("nearest", "box", "bilinear", "hamming", "bicubic", "lanczos")
_F32_054, _F32_046 = float(np.float32(0.54)), float(np.float32(0.46))     # the C source writes 0.54f / 0.46f


def _box(x):
    return 1.0 if -0.5 < x <= 0.5 else 0.0


def _bilinear(x):
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def _hamming(x):
    x = abs(x)
    if x == 0.0:
        return 1.0
    if x >= 1.0:
        return 0.0
    x *= math.pi
    return math.sin(x) / x * (_F32_054 + _F32_046 * math.cos(x))


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _sinc(x):
    if x == 0.0:
        return 1.0
    x *= math.pi
    return math.sin(x) / x


def _lanczos(x):
    return _sinc(x) * _sinc(x / 3) if -3.0 <= x < 3.0 else 0.0


_SUPPORT = {"box": (_box, 0.5), "bilinear": (_bilinear, 1.0), "hamming": (_hamming, 1.0), "bicubic": (_bicubic, 2.0), "lanczos": (_lanczos, 3.0)}


def precompute_coeffs(in_size, out_size, name):
    """Resample.c precompute_coeffs (box = the whole axis) + normalize_coeffs_8bpc -> (xmin[out], xcnt[out], coef[out, ksize] int32)."""
    f, support0 = _SUPPORT[name]
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin, xcnt = np.zeros(out_size, np.int32), np.zeros(out_size, np.int32)
    coef = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        cnt = hi - lo
        k = [f((x + lo - center + 0.5) * ss) for x in range(cnt)]
        ww = 0.0
        for w in k:
            ww += w
        for x in range(cnt):
            v = k[x] / ww if ww != 0.0 else k[x]
            v *= (1 << PRECISION_BITS)
            coef[xx, x] = int(-0.5 + v) if v < 0 else int(0.5 + v)
        xmin[xx], xcnt[xx] = lo, cnt
    return xmin, xcnt, coef


def nearest_index(in_size, out_size):
    """Geometry.c ImagingScaleAffine: the running double sum xo = a*0.5; xo += a."""
    a = in_size / out_size
    xo = a * 0.5
    idx = np.zeros(out_size, np.int32)
    for x in range(out_size):
        xin = -1 if xo < 0.0 else int(xo)
        idx[x] = min(max(xin, 0), in_size - 1)          # always inside for a pure scale
        xo += a
    return idx


def _pass(img, xmin, xcnt, coef, axis):
    """One 8bpc pass along `axis` (0 = vertical, 1 = horizontal) of an [h, w, c] uint8 image."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((len(xmin),) + src.shape[1:], np.uint8)
    for xx in range(len(xmin)):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for j in range(int(xcnt[xx])):
            acc += src[xmin[xx] + j] * int(coef[xx, j])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize(img, out_h, out_w, name):
    """PIL.Image.fromarray(img).resize((out_w, out_h), FILTER) for an [h, w, 3] uint8 array."""
    h, w = img.shape[:2]
    if (h, w) == (out_h, out_w):
        return img.copy()
    if name == "nearest":
        return img[nearest_index(h, out_h)][:, nearest_index(w, out_w)]
    out = img
    if out_w != w:
        tab_h = precompute_coeffs(w, out_w, name)
        if out_h != h:                       # only the rows the vertical pass reads (ImagingResample ybox_first / ybox_last)
            vmin, vcnt, vco = precompute_coeffs(h, out_h, name)
            first, last = int(vmin[0]), int(vmin[-1] + vcnt[-1])
            tmp = _pass(img[first:last], *tab_h, axis=1)
            return _pass(tmp, vmin - first, vcnt, vco, axis=0)
        return _pass(img, *tab_h, axis=1)
    return _pass(out, *precompute_coeffs(h, out_h, name), axis=0)


def imagenet_s_val(img, resize_type="pil-bilinear", size=224):
    """ImageTransfer.image_resize, transform_type 'val', pil-* types (imagenet_s_gen.py:131-141): resize to
    (size*8/7, size*8/7) -- not aspect preserving -- then the centre crop with the reference's rounding."""
    name = {"pil-bilinear": "bilinear", "pil-nearest": "nearest", "pil-box": "box", "pil-hamming": "hamming",
            "pil-cubic": "bicubic", "pil-lanczos": "lanczos"}[resize_type]
    first = int(size * 8 / 7)
    out = resize(img, first, first, name)
    i = int(round((first - size) / 2.0))
    return out[i:i + size, i:i + size]
