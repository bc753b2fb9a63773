"""cv2.resize for uint8 images, INTER_NEAREST, INTER_LINEAR, INTER_AREA, INTER_LANCZOS4 and INTER_CUBIC, restated -- TEST INFRASTRUCTURE ONLY (see oracle/imagenet_c.py).

The reference's ImageNet-S generator calls cv2.resize for the `opencv-*` resize types (RobustART/noise/utils/imagenet_s_gen.py:
28-34,120-148).  OpenCV is a third-party dependency of the reference; it is installed in this container (4.13.0), so this
restatement of imgproc/resize.cpp (resizeNN; resizeGeneric_ with HResizeLinear / VResizeLinear on 11-bit fixed-point
coefficients; INTER_AREA's three regimes: integer factors -> resizeAreaFast_, both axes shrinking -> resizeArea_ on
computeResizeAreaTab weights, otherwise the linear kernel on "area mode" coefficients) is pinned against cv2.resize itself: bit-exact for up- and down-scaling, degenerate sizes included
(tests/test_oracle_cpu.py::test_cv_resize_restatement).  INTER_LANCZOS4 is OpenCV's own 8-tap fixed-point path (bit-exact here).
INTER_CUBIC is different: the opencv-python wheels route it to Intel IPP (cv2.ipp.setUseIPP(False) switches back to the 4-tap
fixed-point path), whose result is a plain float32 Keys cubic (A = -0.75, unquantised weights); that float form agrees with cv2 on
all but <= 3 pixels in 200 000 (never by more than 1) and is what is restated here -- IPP is closed source, so "within 1 LSB on
<= 1e-4 of the pixels" is the parity bar for this one type.  csrc/resize_cv.cu follows this file.
"""
import math
import numpy as np

COEF_BITS = 11
ONE = 1 << COEF_BITS


def _scale(nin, nout):
    return 1.0 / (float(nout) / float(nin))        # resize.cpp: scale = 1. / inv_scale, inv_scale = dsize / ssize (doubles)


def linear_coeffs(nin, nout, clamp):
    """(source index, [w0, w1] in 1/2048) per output coordinate.  Horizontal (clamp=True): the interpolation weight is zeroed where
    the tap pair leaves the row; vertical (clamp=False): weights are kept and the ROW indices are clipped instead."""
    scale = _scale(nin, nout)
    idx = np.zeros(nout, np.int64)
    w = np.zeros((nout, 2), np.int64)
    for d in range(nout):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp:
            if s < 0:
                s, f = 0, np.float32(0)
            if s >= nin - 1:
                s, f = nin - 1, np.float32(0)
        idx[d] = s
        w[d, 0] = int(np.rint(np.float32((np.float32(1) - f) * np.float32(ONE))))
        w[d, 1] = int(np.rint(np.float32(f * np.float32(ONE))))
    return idx, w


def resize_linear(img, wout, hout):
    hin, win, _ = img.shape
    xi, xa = linear_coeffs(win, wout, True)
    yi, ya = linear_coeffs(hin, hout, False)
    I = img.astype(np.int64)
    x1 = np.minimum(xi + 1, win - 1)
    rows = I[:, xi, :] * xa[:, 0][None, :, None] + I[:, x1, :] * xa[:, 1][None, :, None]
    S0, S1 = rows[np.clip(yi, 0, hin - 1)], rows[np.clip(yi + 1, 0, hin - 1)]
    b0, b1 = ya[:, 0][:, None, None], ya[:, 1][:, None, None]
    out = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def resize_nearest(img, wout, hout):
    hin, win, _ = img.shape
    xs = np.minimum(np.floor(np.arange(wout) * _scale(win, wout)).astype(np.int64), win - 1)
    ys = np.minimum(np.floor(np.arange(hout) * _scale(hin, hout)).astype(np.int64), hin - 1)
    return img[ys][:, xs]


def area_linear_coeffs(nin, nout, clamp):
    """resize.cpp, area_mode branch of the linear coefficient loop (INTER_AREA when an axis grows)."""
    inv = float(nout) / float(nin)
    scale = 1.0 / inv
    idx = np.zeros(nout, np.int64)
    w = np.zeros((nout, 2), np.int64)
    for d in range(nout):
        s = int(np.floor(d * scale))
        f = np.float32((d + 1) - (s + 1) * inv)
        f = np.float32(0) if f <= 0 else np.float32(f - np.floor(f))
        if clamp:
            if s < 0:
                s, f = 0, np.float32(0)
            if s >= nin - 1:
                s, f = nin - 1, np.float32(0)
        idx[d] = s
        w[d, 0] = int(np.rint(np.float32((np.float32(1) - f) * np.float32(ONE))))
        w[d, 1] = int(np.rint(np.float32(f * np.float32(ONE))))
    return idx, w


def _two_tap(img, wout, hout, coeffs):
    hin, win, _ = img.shape
    xi, xa = coeffs(win, wout, True)
    yi, ya = coeffs(hin, hout, False)
    I = img.astype(np.int64)
    x1 = np.minimum(xi + 1, win - 1)
    rows = I[:, xi, :] * xa[:, 0][None, :, None] + I[:, x1, :] * xa[:, 1][None, :, None]
    S0, S1 = rows[np.clip(yi, 0, hin - 1)], rows[np.clip(yi + 1, 0, hin - 1)]
    b0, b1 = ya[:, 0][:, None, None], ya[:, 1][:, None, None]
    return np.clip((((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2, 0, 255).astype(np.uint8)


def area_tab(ssize, dsize, scale):
    """computeResizeAreaTab: (dst index, src index, float32 weight) in OpenCV's order."""
    tab = []
    for d in range(dsize):
        f1 = d * scale
        f2 = f1 + scale
        cell = min(scale, ssize - f1)
        s1, s2 = int(np.ceil(f1)), int(np.floor(f2))
        s2 = min(s2, ssize - 1)
        s1 = min(s1, s2)
        if s1 - f1 > 1e-3:
            tab.append((d, s1 - 1, np.float32((s1 - f1) / cell)))
        for sx in range(s1, s2):
            tab.append((d, sx, np.float32(1.0 / cell)))
        if f2 - s2 > 1e-3:
            tab.append((d, s2, np.float32(min(min(f2 - s2, 1.), cell) / cell)))
    return tab


def resize_area(img, wout, hout):
    """cv2.resize(..., interpolation=INTER_AREA)."""
    hin, win, C = img.shape
    sx, sy = _scale(win, wout), _scale(hin, hout)
    if not (sx >= 1 and sy >= 1):
        return _two_tap(img, wout, hout, area_linear_coeffs)
    kx, ky = int(sx), int(sy)
    if abs(sx - kx) < np.finfo(np.float64).eps and abs(sy - ky) < np.finfo(np.float64).eps:      # resizeAreaFast_
        I = img.astype(np.int64)[:hout * ky, :wout * kx].reshape(hout, ky, wout, kx, C).sum((1, 3))
        if kx == 2 and ky == 2:
            return ((I + 2) >> 2).astype(np.uint8)
        return np.clip(np.rint(I.astype(np.float32) * np.float32(1.0 / (kx * ky))), 0, 255).astype(np.uint8)
    xtab, ytab = area_tab(win, wout, sx), area_tab(hin, hout, sy)
    I = img.astype(np.float32)
    rows = {}
    sums = np.zeros((hout, wout, C), np.float32)
    started = set()
    for (dy, sy_, beta) in ytab:
        if sy_ not in rows:
            b = np.zeros((wout, C), np.float32)
            for (dx, sx_, a) in xtab:
                b[dx] = b[dx] + I[sy_, sx_] * a
            rows[sy_] = b
        sums[dy] = beta * rows[sy_] if dy not in started else sums[dy] + beta * rows[sy_]
        started.add(dy)
    return np.clip(np.rint(sums), 0, 255).astype(np.uint8)


def lanczos4_weights(x):
    """interpolateLanczos4 (imgproc/precomp / resize.cpp): 8 float32 weights for the fractional position x (float32)."""
    s45 = 0.70710678118654752440084436210485
    cs = ((1, 0), (-s45, -s45), (0, 1), (s45, -s45), (-1, 0), (s45, s45), (0, -1), (-s45, s45))
    x = np.float32(x)
    y0 = -(float(x) + 3) * math.pi * 0.25
    s0, c0 = math.sin(y0), math.cos(y0)
    co = np.zeros(8, np.float32)
    sm = np.float32(0)
    for i in range(8):
        d = np.float32(np.float32(x + np.float32(3)) - np.float32(i))
        if abs(d) >= 1e-6:
            y = -float(d) * math.pi * 0.25
            co[i] = np.float32((cs[i][0] * s0 + cs[i][1] * c0) / (y * y))
        else:
            co[i] = np.float32(1e30)
        sm = np.float32(sm + co[i])
    sm = np.float32(np.float32(1) / sm)
    return np.array([np.float32(c * sm) for c in co], np.float32)


def lanczos4_table(nin, nout):
    """(first-tap anchor sx, 8 weights in 1/2048) per output coordinate; taps are sx - 3 .. sx + 4, indices clipped."""
    scale = _scale(nin, nout)
    idx = np.zeros(nout, np.int64)
    w = np.zeros((nout, 8), np.int64)
    for d in range(nout):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        idx[d] = s
        w[d] = [int(np.rint(np.float32(v * np.float32(ONE)))) for v in lanczos4_weights(np.float32(f - np.float32(s)))]
    return idx, w


def resize_lanczos4(img, wout, hout):
    hin, win, _ = img.shape
    xi, xq = lanczos4_table(win, wout)
    yi, yq = lanczos4_table(hin, hout)
    I = img.astype(np.int64)
    rows = sum(I[:, np.clip(xi - 3 + k, 0, win - 1), :] * xq[:, k][None, :, None] for k in range(8))
    v = (sum(rows[np.clip(yi - 3 + k, 0, hin - 1)] * yq[:, k][:, None, None] for k in range(8)) + (1 << 21)) >> 22
    return np.clip(v, 0, 255).astype(np.uint8)


def cubic_table(nin, nout):
    """(anchor sx, 4 float32 Keys weights, A = -0.75) per output coordinate, evaluated in double; taps sx - 1 .. sx + 2, clipped."""
    scale = _scale(nin, nout)
    A = -0.75
    idx = np.zeros(nout, np.int64)
    w = np.zeros((nout, 4), np.float32)
    for d in range(nout):
        f = (d + 0.5) * scale - 0.5
        s = int(np.floor(f))
        x = f - s
        c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
        c1 = ((A + 2) * x - (A + 3)) * x * x + 1
        c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
        idx[d] = s
        w[d] = [c0, c1, c2, 1 - c0 - c1 - c2]
    return idx, w


def resize_cubic(img, wout, hout):
    """float32 separable cubic, horizontal pass first, taps accumulated left to right / top to bottom, round-half-even."""
    hin, win, C = img.shape
    xi, xa = cubic_table(win, wout)
    yi, ya = cubic_table(hin, hout)
    I = img.astype(np.float32)
    rows = np.zeros((hin, wout, C), np.float32)
    for k in range(4):
        rows = rows + I[:, np.clip(xi - 1 + k, 0, win - 1), :] * xa[:, k][None, :, None]
    v = np.zeros((hout, wout, C), np.float32)
    for k in range(4):
        v = v + rows[np.clip(yi - 1 + k, 0, hin - 1)] * ya[:, k][:, None, None]
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def resize(img, wout, hout, interpolation):
    return {"nearest": resize_nearest, "bilinear": resize_linear, "area": resize_area, "lanczos": resize_lanczos4,
            "cubic": resize_cubic}[interpolation](img, wout, hout)


def imagenet_s_val(img, resize_type, size=224):
    """ImageTransfer.image_resize, transform 'val', opencv-* types (imagenet_s_gen.py:138-148): resize to int(size*8/7) squared,
    then the centre crop."""
    first = int(size * 8 / 7)
    full = resize(img, first, first, {"opencv-nearest": "nearest", "opencv-bilinear": "bilinear", "opencv-area": "area", "opencv-cubic": "cubic",
                              "opencv-lanczos": "lanczos"}[resize_type])
    d = int(round((first - size) / 2.))
    return full[d:d + size, d:d + size]
