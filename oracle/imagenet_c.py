"""CPU oracle for RobustART's ImageNet-C corruptions -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (robustart_b200 / RobustART.noise) never does.

This is a restatement of /root/reference/RobustART/noise/utils/imagenet_c/corruptions.py and
imagenet_c/__init__.py (cited per function as corruptions.py:LINE) that

  * runs on NumPy 2 / SciPy / OpenCV / Pillow as found in this image (the reference file itself
    cannot be imported: it needs scikit-image, Wand/ImageMagick and pkg_resources and uses
    np.float_ / binary np.fromstring that NumPy 2 removed),
  * replaces third-party calls that are absent by their published algorithms
      - skimage.filters.gaussian(multichannel=True)  -> scipy.ndimage.gaussian_filter(sigma=[s,s,0],
        mode='nearest', truncate=4.0)  (skimage 0.17.2 filters/_gaussian.py)
      - skimage.util.random_noise(mode='s&p')        -> two np.random.choice draws (util/noise.py)
      - skimage.color.rgb2hsv / hsv2rgb              -> color/colorconv.py formulas
      - Wand MagickMotionBlurImage                   -> ImageMagick 6 effect.c MotionBlurImage
  * takes its randomness from a `Draws` object so the very same draws can be handed to the CUDA
    kernels ("shared-draw parity mode").  NumpyDraws issues exactly the np.random calls the
    reference issues, in the same order, so oracle(seed) == reference(np.random.seed(seed)).

Parity status: the reference ships no tests or golden vectors.  C1-C3, C9-C12,
C16, C19 are pure NumPy and pinned by construction; C4, C7, C13, C14, C15, C17, C18 call the same
OpenCV / SciPy / Pillow functions the reference calls; C5 uses the SciPy filter skimage wraps;
C6 and C8 depend on ImageMagick, absent here: *parity unpinned* for the motion-blur step.
"""
from __future__ import annotations

import io
import math
from typing import Optional

import numpy as np

CORRUPTION_NAMES = (
    "gaussian_noise", "shot_noise", "impulse_noise", "defocus_blur", "glass_blur", "motion_blur",
    "zoom_blur", "snow", "frost", "fog", "brightness", "contrast", "elastic_transform", "pixelate",
    "jpeg_compression", "speckle_noise", "gaussian_blur", "spatter", "saturate",
)  # imagenet_c/__init__.py:5-8


# ------------------------------------------------------------------------------------------------
# randomness
# ------------------------------------------------------------------------------------------------
class NumpyDraws:
    """Issues the np.random calls of the reference, in order, from a private RandomState."""

    def __init__(self, seed: int = 0):
        self.rs = np.random.RandomState(seed)
        self.log = []  # (kind, ndarray) in draw order -- fed to the GPU as ext_noise

    def _rec(self, kind, a):
        self.log.append((kind, np.asarray(a)))
        return a

    def normal(self, size, loc=0.0, scale=1.0):
        # reference calls np.random.normal(size=, loc=, scale=); we record the *standard* normal so
        # the GPU can apply loc/scale itself: loc + scale*z is what numpy computes internally.
        z = self.rs.standard_normal(size)
        self._rec("normal", z)
        return loc + scale * z

    def poisson(self, lam):
        return self._rec("poisson", self.rs.poisson(lam))

    def uniform(self, low, high, size=None):
        # numpy: low + (high-low)*random_sample()
        u = self.rs.random_sample(size)
        self._rec("uniform01", u)
        return low + (high - low) * u

    def randint(self, low, high=None, size=None):
        return self._rec("randint", self.rs.randint(low, high, size=size))

    def choice_bool(self, p_true, size):
        # np.random.choice([True, False], size, p=[p, 1-p]) == cdf.searchsorted(u, 'right') == 0
        u = self.rs.random_sample(size)
        self._rec("uniform01", u)
        cdf = np.cumsum(np.array([p_true, 1.0 - p_true]))
        cdf /= cdf[-1]
        return cdf.searchsorted(u, side="right") == 0


class ReplayDraws(NumpyDraws):
    """Replays a recorded log (used to make the oracle consume GPU-generated draws)."""

    def __init__(self, log):
        self.log = []
        self._it = iter(log)

    def _next(self, kind):
        k, a = next(self._it)
        assert k == kind, (k, kind)
        self.log.append((k, a))
        return a

    def normal(self, size, loc=0.0, scale=1.0):
        return loc + scale * self._next("normal").reshape(size)

    def poisson(self, lam):
        return self._next("poisson").reshape(np.shape(lam))

    def uniform(self, low, high, size=None):
        u = self._next("uniform01")
        return low + (high - low) * (u.reshape(size) if size is not None else u)

    def randint(self, low, high=None, size=None):
        return self._next("randint")

    def choice_bool(self, p_true, size):
        u = self._next("uniform01").reshape(size)
        cdf = np.cumsum(np.array([p_true, 1.0 - p_true]))
        cdf /= cdf[-1]
        return cdf.searchsorted(u, side="right") == 0


# ------------------------------------------------------------------------------------------------
# third-party restatements
# ------------------------------------------------------------------------------------------------
def sk_gaussian(img, sigma, mode="nearest", truncate=4.0, multichannel=None):
    """skimage.filters.gaussian (0.17.2): float image -> scipy.ndimage.gaussian_filter."""
    from scipy import ndimage as ndi
    img = np.asarray(img)
    if img.dtype.kind != "f":  # img_as_float
        img = img.astype(np.float64) / np.iinfo(img.dtype).max
    if multichannel is None:
        multichannel = img.ndim == 3 and img.shape[-1] == 3  # skimage guesses RGB
    if multichannel:
        sig = [sigma] * (img.ndim - 1) + [0]
    else:
        sig = sigma
    out = np.empty_like(img, dtype=np.float64 if img.dtype != np.float32 else np.float32)
    ndi.gaussian_filter(img, sig, output=out, mode=mode, truncate=truncate)
    return out


def sk_rgb2hsv(rgb):
    """skimage.color.rgb2hsv (0.17.2 color/colorconv.py)."""
    arr = np.asarray(rgb, dtype=np.float64)
    out = np.empty_like(arr)
    out_v = arr.max(-1)
    delta = np.ptp(arr, -1)
    with np.errstate(invalid="ignore", divide="ignore"):
        out_s = delta / out_v
        out_s[delta == 0.0] = 0.0
        idx = arr[..., 0] == out_v
        out[idx, 0] = (arr[idx, 1] - arr[idx, 2]) / delta[idx]
        idx = arr[..., 1] == out_v
        out[idx, 0] = 2.0 + (arr[idx, 2] - arr[idx, 0]) / delta[idx]
        idx = arr[..., 2] == out_v
        out[idx, 0] = 4.0 + (arr[idx, 0] - arr[idx, 1]) / delta[idx]
        out_h = (out[..., 0] / 6.0) % 1.0
        out_h[delta == 0.0] = 0.0
    out[..., 0] = out_h
    out[..., 1] = out_s
    out[..., 2] = out_v
    out[np.isnan(out)] = 0
    return out


def sk_hsv2rgb(hsv):
    """skimage.color.hsv2rgb (0.17.2)."""
    arr = np.asarray(hsv, dtype=np.float64)
    hi = np.floor(arr[..., 0] * 6)
    f = arr[..., 0] * 6 - hi
    p = arr[..., 2] * (1 - arr[..., 1])
    q = arr[..., 2] * (1 - f * arr[..., 1])
    t = arr[..., 2] * (1 - (1 - f) * arr[..., 1])
    v = arr[..., 2]
    hi = np.stack([hi, hi, hi], axis=-1).astype(np.uint8) % 6
    out = np.choose(hi, np.stack([np.stack((v, t, p), axis=-1), np.stack((q, v, p), axis=-1),
                                  np.stack((p, v, t), axis=-1), np.stack((p, q, v), axis=-1),
                                  np.stack((t, p, v), axis=-1), np.stack((v, p, q), axis=-1)]))
    return out


def motion_blur_kernel(radius: float, sigma: float, angle_deg: float):
    """ImageMagick 6 effect.c: GetMotionBlurKernel + offsets of MotionBlurImage.
    Returns (weights[width], offs_x[width], offs_y[width])."""
    width = int(2.0 * math.ceil(radius) + 1.0) if radius > 1e-12 else 3  # GetOptimalKernelWidth1D
    i = np.arange(width, dtype=np.float64)
    k = np.exp(-(i * i) / (2.0 * sigma * sigma)) / (math.sqrt(2.0 * math.pi) * sigma)
    k /= k.sum()
    ang = math.radians(angle_deg)
    px, py = width * math.sin(ang), width * math.cos(ang)
    hyp = math.hypot(px, py)
    offx = np.ceil(i * py / hyp - 0.5).astype(np.int64)
    offy = np.ceil(i * px / hyp - 0.5).astype(np.int64)
    return k, offx, offy


def magick_motion_blur_u8(img_u8, radius, sigma, angle_deg):
    """MotionBlurImage on an 8-bit image through a Q16 build: u8*257 -> weighted sum over the kernel
    with edge virtual pixels -> ClampToQuantum (round half up) -> ScaleQuantumToChar."""
    k, offx, offy = motion_blur_kernel(radius, sigma, angle_deg)
    a = img_u8.astype(np.float64) * 257.0
    h, w = a.shape[:2]
    yy, xx = np.mgrid[0:h, 0:w]
    acc = np.zeros_like(a)
    for ki, ox, oy in zip(k, offx, offy):
        ys = np.clip(yy + oy, 0, h - 1)
        xs = np.clip(xx + ox, 0, w - 1)
        acc += ki * a[ys, xs]
    q = np.floor(np.clip(acc, 0, 65535.0) + 0.5)
    return np.floor((q + 128.0) / 257.0).astype(np.uint8)  # ScaleQuantumToChar


# ------------------------------------------------------------------------------------------------
# helpers (corruptions.py:26-114)
# ------------------------------------------------------------------------------------------------
def disk(radius, alias_blur=0.1, dtype=np.float32):  # corruptions.py:26-38
    import cv2
    if radius <= 8:
        L = np.arange(-8, 8 + 1)
        ksize = (3, 3)
    else:
        L = np.arange(-radius, radius + 1)
        ksize = (5, 5)
    X, Y = np.meshgrid(L, L)
    aliased_disk = np.array((X ** 2 + Y ** 2) <= radius ** 2, dtype=dtype)
    aliased_disk /= np.sum(aliased_disk)
    return cv2.GaussianBlur(aliased_disk, ksize=ksize, sigmaX=alias_blur)


def plasma_fractal(draws, mapsize=256, wibbledecay=3):  # corruptions.py:55-101
    assert mapsize & (mapsize - 1) == 0
    maparray = np.empty((mapsize, mapsize), dtype=np.float64)
    maparray[0, 0] = 0
    stepsize = mapsize
    wibble = 100.0

    def wibbledmean(array):
        return array / 4 + wibble * draws.uniform(-wibble, wibble, array.shape)

    while stepsize >= 2:
        # fillsquares
        cornerref = maparray[0:mapsize:stepsize, 0:mapsize:stepsize]
        squareaccum = cornerref + np.roll(cornerref, shift=-1, axis=0)
        squareaccum += np.roll(squareaccum, shift=-1, axis=1)
        maparray[stepsize // 2:mapsize:stepsize, stepsize // 2:mapsize:stepsize] = wibbledmean(squareaccum)
        # filldiamonds
        drgrid = maparray[stepsize // 2:mapsize:stepsize, stepsize // 2:mapsize:stepsize]
        ulgrid = maparray[0:mapsize:stepsize, 0:mapsize:stepsize]
        ldrsum = drgrid + np.roll(drgrid, 1, axis=0)
        lulsum = ulgrid + np.roll(ulgrid, -1, axis=1)
        ltsum = ldrsum + lulsum
        maparray[0:mapsize:stepsize, stepsize // 2:mapsize:stepsize] = wibbledmean(ltsum)
        tdrsum = drgrid + np.roll(drgrid, 1, axis=1)
        tulsum = ulgrid + np.roll(ulgrid, -1, axis=0)
        ttsum = tdrsum + tulsum
        maparray[stepsize // 2:mapsize:stepsize, 0:mapsize:stepsize] = wibbledmean(ttsum)
        stepsize //= 2
        wibble /= wibbledecay

    maparray -= maparray.min()
    return maparray / maparray.max()


def clipped_zoom(img, zoom_factor):  # corruptions.py:104-114
    from scipy.ndimage import zoom as scizoom
    h = img.shape[0]
    ch = int(np.ceil(h / float(zoom_factor)))
    top = (h - ch) // 2
    img = scizoom(img[top:top + ch, top:top + ch], (zoom_factor, zoom_factor, 1), order=1)
    trim_top = (img.shape[0] - h) // 2
    return img[trim_top:trim_top + h, trim_top:trim_top + h]


ZOOM_FACTORS = [np.arange(1, 1.11, 0.01), np.arange(1, 1.16, 0.01), np.arange(1, 1.21, 0.02),
                np.arange(1, 1.26, 0.02), np.arange(1, 1.31, 0.03)]  # corruptions.py:220-224


# ------------------------------------------------------------------------------------------------
# the 19 corruptions; x is uint8 [H,W,3]; return float [0,255] or uint8 like the reference
# ------------------------------------------------------------------------------------------------
def gaussian_noise(x, severity, draws):  # corruptions.py:122-126
    c = [.08, .12, 0.18, 0.26, 0.38][severity - 1]
    x = np.array(x) / 255.
    return np.clip(x + draws.normal(size=x.shape, scale=c), 0, 1) * 255


def shot_noise(x, severity, draws):  # corruptions.py:129-133
    c = [60, 25, 12, 5, 3][severity - 1]
    x = np.array(x) / 255.
    return np.clip(draws.poisson(x * c) / float(c), 0, 1) * 255


def impulse_noise(x, severity, draws):  # corruptions.py:136-140 + skimage util/noise.py 's&p'
    c = [.03, .06, .09, 0.17, 0.27][severity - 1]
    out = np.array(x) / 255.
    flipped = draws.choice_bool(c, out.shape)
    salted = draws.choice_bool(0.5, out.shape)
    peppered = ~salted
    out[flipped & salted] = 1
    out[flipped & peppered] = 0
    return np.clip(out, 0, 1) * 255


def speckle_noise(x, severity, draws):  # corruptions.py:143-147
    c = [.15, .2, 0.35, 0.45, 0.6][severity - 1]
    x = np.array(x) / 255.
    return np.clip(x + x * draws.normal(size=x.shape, scale=c), 0, 1) * 255


def gaussian_blur(x, severity, draws=None):  # corruptions.py:162-166
    c = [1, 2, 3, 4, 6][severity - 1]
    x = sk_gaussian(np.array(x) / 255., sigma=c, multichannel=True)
    return np.clip(x, 0, 1) * 255


GLASS_PARAMS = [(0.7, 1, 2), (0.9, 2, 1), (1, 2, 3), (1.1, 3, 2), (1.5, 4, 2)]  # corruptions.py:171


def glass_blur(x, severity, draws):  # corruptions.py:169-184
    c = GLASS_PARAMS[severity - 1]
    x = np.uint8(sk_gaussian(np.array(x) / 255., sigma=c[0], multichannel=True) * 255)
    H = x.shape[0]
    # the reference draws randint(-d, d, size=(2,)) per visited pixel; one vectorised draw of the
    # same total size from the same stream yields the same numbers in the same order.
    n_pix = (H - 2 * c[1]) * (H - 2 * c[1])
    for _ in range(c[2]):
        d = draws.randint(-c[1], c[1], size=(n_pix, 2))
        j = 0
        for h in range(H - c[1], c[1], -1):
            for w in range(H - c[1], c[1], -1):
                dx, dy = d[j]
                j += 1
                hp, wp = h + dy, w + dx
                # `x[h, w], x[hp, wp] = x[hp, wp], x[h, w]` (corruptions.py:182) on a 3-channel array
                # swaps VIEWS: the first assignment already overwrote x[h, w] when the second one
                # reads it, so the net effect is a copy x[h, w] <- x[hp, wp]; x[hp, wp] keeps its value.
                # (Verified against the reference's own code by tests/golden/make_golden.py.)
                x[h, w] = x[hp, wp]
    return np.clip(sk_gaussian(x / 255., sigma=c[0], multichannel=True), 0, 1) * 255


def defocus_blur(x, severity, draws=None):  # corruptions.py:187-198
    import cv2
    c = [(3, 0.1), (4, 0.5), (6, 0.5), (8, 0.5), (10, 0.5)][severity - 1]
    x = np.array(x) / 255.
    kernel = disk(radius=c[0], alias_blur=c[1])
    channels = []
    for d in range(3):
        channels.append(cv2.filter2D(x[:, :, d], -1, kernel))
    channels = np.array(channels).transpose((1, 2, 0))
    return np.clip(channels, 0, 1) * 255


MOTION_PARAMS = [(10, 3), (15, 5), (15, 8), (15, 12), (20, 15)]  # corruptions.py:202


def motion_blur(x, severity, draws):  # corruptions.py:201-216 (PNG round trips are lossless)
    c = MOTION_PARAMS[severity - 1]
    angle = draws.uniform(-45, 45)
    out = magick_motion_blur_u8(np.array(x), c[0], c[1], float(angle))
    return np.clip(out, 0, 255)


def zoom_blur(x, severity, draws=None):  # corruptions.py:219-232
    c = ZOOM_FACTORS[severity - 1]
    x = (np.array(x) / 255.).astype(np.float32)
    out = np.zeros_like(x)
    for zoom_factor in c:
        out += clipped_zoom(x, zoom_factor)
    x = (x + out) / (len(c) + 1)
    return np.clip(x, 0, 1) * 255


FOG_PARAMS = [(1.5, 2), (2., 2), (2.5, 1.7), (2.5, 1.5), (3., 1.4)]  # corruptions.py:236


def fog(x, severity, draws):  # corruptions.py:235-241
    c = FOG_PARAMS[severity - 1]
    x = np.array(x) / 255.
    H = x.shape[0]
    max_val = x.max()
    x = x + c[0] * plasma_fractal(draws, wibbledecay=c[1])[:H, :H][..., np.newaxis]
    return np.clip(x * max_val / (max_val + c[0]), 0, 1) * 255


FROST_PARAMS = [(1, 0.4), (0.8, 0.6), (0.7, 0.7), (0.65, 0.7), (0.6, 0.75)]  # corruptions.py:245-249


def frost(x, severity, draws, textures=None):  # corruptions.py:244-262
    """textures: list of 6 uint8 RGB arrays (the reference's frost1..6 files, absent from its repo;
    cv2.imread returns BGR and the reference flips to RGB, so RGB arrays are equivalent)."""
    c = FROST_PARAMS[severity - 1]
    H = np.array(x).shape[0]
    idx = int(draws.randint(5))
    tex = textures[idx]
    x_start = int(draws.randint(0, tex.shape[0] - H))
    y_start = int(draws.randint(0, tex.shape[1] - H))
    crop = tex[x_start:x_start + H, y_start:y_start + H]
    return np.clip(c[0] * np.array(x) + c[1] * crop, 0, 255)


SNOW_PARAMS = [(0.1, 0.3, 3, 0.5, 10, 4, 0.8), (0.2, 0.3, 2, 0.5, 12, 4, 0.7),
               (0.55, 0.3, 4, 0.9, 12, 8, 0.7), (0.55, 0.3, 4.5, 0.85, 12, 8, 0.65),
               (0.55, 0.3, 2.5, 0.85, 12, 12, 0.55)]  # corruptions.py:266-270


def snow(x, severity, draws):  # corruptions.py:265-290
    import cv2
    c = SNOW_PARAMS[severity - 1]
    x = np.array(x, dtype=np.float32) / 255.
    H = x.shape[0]
    snow_layer = draws.normal(size=x.shape[:2], loc=c[0], scale=c[1])
    snow_layer = clipped_zoom(snow_layer[..., np.newaxis], c[2])
    snow_layer[snow_layer < c[3]] = 0
    layer_u8 = (np.clip(snow_layer.squeeze(), 0, 1) * 255).astype(np.uint8)
    angle = draws.uniform(-135, -45)
    layer_u8 = magick_motion_blur_u8(layer_u8, c[4], c[5], float(angle))
    snow_layer = (layer_u8 / 255.)[..., np.newaxis]
    x = c[6] * x + (1 - c[6]) * np.maximum(
        x, cv2.cvtColor(x, cv2.COLOR_RGB2GRAY).reshape(H, H, 1) * 1.5 + 0.5)
    return np.clip(x + snow_layer + np.rot90(snow_layer, k=2), 0, 1) * 255


SPATTER_PARAMS = [(0.65, 0.3, 4, 0.69, 0.6, 0), (0.65, 0.3, 3, 0.68, 0.6, 0),
                  (0.65, 0.3, 2, 0.68, 0.5, 0), (0.65, 0.3, 1, 0.65, 1.5, 1),
                  (0.67, 0.4, 1, 0.65, 1.5, 1)]  # corruptions.py:294-298


def spatter(x, severity, draws):  # corruptions.py:293-342
    import cv2
    c = SPATTER_PARAMS[severity - 1]
    x = np.array(x, dtype=np.float32) / 255.
    liquid_layer = draws.normal(size=x.shape[:2], loc=c[0], scale=c[1])
    liquid_layer = sk_gaussian(liquid_layer, sigma=c[2], multichannel=False)
    liquid_layer[liquid_layer < c[3]] = 0
    if c[5] == 0:
        liquid_layer = (liquid_layer * 255).astype(np.uint8)
        dist = 255 - cv2.Canny(liquid_layer, 50, 150)
        dist = cv2.distanceTransform(dist, cv2.DIST_L2, 5)
        _, dist = cv2.threshold(dist, 20, 20, cv2.THRESH_TRUNC)
        dist = cv2.blur(dist, (3, 3)).astype(np.uint8)
        dist = cv2.equalizeHist(dist)
        ker = np.array([[-2, -1, 0], [-1, 1, 1], [0, 1, 2]])
        dist = cv2.filter2D(dist, cv2.CV_8U, ker)
        dist = cv2.blur(dist, (3, 3)).astype(np.float32)
        m = cv2.cvtColor(liquid_layer * dist, cv2.COLOR_GRAY2BGRA)
        m /= np.max(m, axis=(0, 1))
        m *= c[4]
        color = np.concatenate((175 / 255. * np.ones_like(m[..., :1]),
                                238 / 255. * np.ones_like(m[..., :1]),
                                238 / 255. * np.ones_like(m[..., :1])), axis=2)
        color = cv2.cvtColor(color, cv2.COLOR_BGR2BGRA)
        x = cv2.cvtColor(x, cv2.COLOR_BGR2BGRA)
        return cv2.cvtColor(np.clip(x + m * color, 0, 1), cv2.COLOR_BGRA2BGR) * 255
    else:
        m = np.where(liquid_layer > c[3], 1, 0)
        m = sk_gaussian(m.astype(np.float32), sigma=c[4], multichannel=False)
        m[m < 0.8] = 0
        color = np.concatenate((63 / 255. * np.ones_like(x[..., :1]),
                                42 / 255. * np.ones_like(x[..., :1]),
                                20 / 255. * np.ones_like(x[..., :1])), axis=2)
        color *= m[..., np.newaxis]
        x *= (1 - m[..., np.newaxis])
        return np.clip(x + color, 0, 1) * 255


def contrast(x, severity, draws=None):  # corruptions.py:345-350
    c = [0.4, .3, .2, .1, .05][severity - 1]
    x = np.array(x) / 255.
    means = np.mean(x, axis=(0, 1), keepdims=True)
    return np.clip((x - means) * c + means, 0, 1) * 255


def brightness(x, severity, draws=None):  # corruptions.py:353-361
    c = [.1, .2, .3, .4, .5][severity - 1]
    x = np.array(x) / 255.
    x = sk_rgb2hsv(x)
    x[:, :, 2] = np.clip(x[:, :, 2] + c, 0, 1)
    x = sk_hsv2rgb(x)
    return np.clip(x, 0, 1) * 255


SATURATE_PARAMS = [(0.3, 0), (0.1, 0), (2, 0), (5, 0.1), (20, 0.2)]  # corruptions.py:365


def saturate(x, severity, draws=None):  # corruptions.py:364-372
    c = SATURATE_PARAMS[severity - 1]
    x = np.array(x) / 255.
    x = sk_rgb2hsv(x)
    x[:, :, 1] = np.clip(x[:, :, 1] * c[0] + c[1], 0, 1)
    x = sk_hsv2rgb(x)
    return np.clip(x, 0, 1) * 255


JPEG_QUALITY = [25, 18, 15, 10, 7]  # corruptions.py:376


def jpeg_compression(x, severity, draws=None):  # corruptions.py:375-382
    from PIL import Image
    c = JPEG_QUALITY[severity - 1]
    output = io.BytesIO()
    Image.fromarray(np.asarray(x)).save(output, 'JPEG', quality=c)
    return np.array(Image.open(output))


PIXELATE_FRAC = [0.6, 0.5, 0.4, 0.3, 0.25]  # corruptions.py:386


def pixelate(x, severity, draws=None):  # corruptions.py:385-391
    from PIL import Image
    c = PIXELATE_FRAC[severity - 1]
    im = Image.fromarray(np.asarray(x))
    H = im.size[0]
    im = im.resize((int(H * c), int(H * c)), Image.BOX)
    im = im.resize((H, H), Image.BOX)
    return np.array(im)


ELASTIC_PARAMS = [(244 * 2, 244 * 0.7, 244 * 0.1), (244 * 2, 244 * 0.08, 244 * 0.2),
                  (244 * 0.05, 244 * 0.01, 244 * 0.02), (244 * 0.07, 244 * 0.01, 244 * 0.02),
                  (244 * 0.12, 244 * 0.01, 244 * 0.02)]  # corruptions.py:396-400


def elastic_transform(image, severity, draws):  # corruptions.py:395-424
    import cv2
    from scipy.ndimage import map_coordinates
    c = ELASTIC_PARAMS[severity - 1]
    image = np.array(image, dtype=np.float32) / 255.
    shape = image.shape
    shape_size = shape[:2]
    center_square = np.float32(shape_size) // 2
    square_size = min(shape_size) // 3
    pts1 = np.float32([center_square + square_size,
                       [center_square[0] + square_size, center_square[1] - square_size],
                       center_square - square_size])
    pts2 = pts1 + draws.uniform(-c[2], c[2], size=pts1.shape).astype(np.float32)
    M = cv2.getAffineTransform(pts1, pts2)
    image = cv2.warpAffine(image, M, shape_size[::-1], borderMode=cv2.BORDER_REFLECT_101)
    dx = (sk_gaussian(draws.uniform(-1, 1, size=shape[:2]), c[1], mode='reflect', truncate=3,
                      multichannel=False) * c[0]).astype(np.float32)
    dy = (sk_gaussian(draws.uniform(-1, 1, size=shape[:2]), c[1], mode='reflect', truncate=3,
                      multichannel=False) * c[0]).astype(np.float32)
    dx, dy = dx[..., np.newaxis], dy[..., np.newaxis]
    x, y, z = np.meshgrid(np.arange(shape[1]), np.arange(shape[0]), np.arange(shape[2]))
    indices = np.reshape(y + dy, (-1, 1)), np.reshape(x + dx, (-1, 1)), np.reshape(z, (-1, 1))
    return np.clip(map_coordinates(image, indices, order=1, mode='reflect').reshape(shape), 0, 1) * 255


_FUNCS = (gaussian_noise, shot_noise, impulse_noise, defocus_blur, glass_blur, motion_blur, zoom_blur,
          snow, frost, fog, brightness, contrast, elastic_transform, pixelate, jpeg_compression,
          speckle_noise, gaussian_blur, spatter, saturate)
CORRUPTION_DICT = {f.__name__: f for f in _FUNCS}
assert tuple(f.__name__ for f in _FUNCS) == CORRUPTION_NAMES


def corrupt(x, severity=1, corruption_name=None, corruption_number=-1, draws: Optional[NumpyDraws] = None,
            **kw):
    """imagenet_c/__init__.py:13-35 -- returns np.uint8(x_corrupted) (truncation)."""
    if draws is None:
        draws = NumpyDraws(0)
    if corruption_name:
        fn = CORRUPTION_DICT[corruption_name]
    elif corruption_number != -1:
        fn = _FUNCS[corruption_number]
    else:
        raise ValueError("Either corruption_name or corruption_number must be passed")
    return np.uint8(fn(np.asarray(x), severity, draws, **kw))


def add_noise_for_imagenet_c(image, severity=1, corruption_name=None, corruption_number=-1, draws=None,
                             **kw):
    """add_noise_utils.py:22-31: per-image Python loop over a uint8 NHWC batch, in place."""
    if draws is None:
        draws = NumpyDraws(0)
    for i in range(image.shape[0]):
        image[i] = corrupt(image[i], severity, corruption_name, corruption_number, draws, **kw)
    return image
