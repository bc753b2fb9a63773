"""cv2.resize for uint8 images, INTER_NEAREST and INTER_LINEAR, restated -- TEST INFRASTRUCTURE ONLY (see oracle/imagenet_c.py).

The reference's ImageNet-S generator calls cv2.resize for the `opencv-*` resize types (RobustART/noise/utils/imagenet_s_gen.py:
28-34,120-148).  OpenCV is a third-party dependency of the reference; it is installed in this container (4.13.0), so this
restatement of imgproc/resize.cpp (resizeNN; resizeGeneric_ with HResizeLinear / VResizeLinear on 11-bit fixed-point
coefficients) is pinned against cv2.resize itself: bit-exact for up- and down-scaling, degenerate sizes included
(tests/test_oracle_cpu.py::test_cv_resize_restatement).  csrc/resize_cv.cu follows this file.
"""
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


def resize(img, wout, hout, interpolation):
    return {"nearest": resize_nearest, "bilinear": resize_linear}[interpolation](img, wout, hout)


def imagenet_s_val(img, resize_type, size=224):
    """ImageTransfer.image_resize, transform 'val', opencv-* types (imagenet_s_gen.py:138-148): resize to int(size*8/7) squared,
    then the centre crop."""
    first = int(size * 8 / 7)
    full = resize(img, first, first, {"opencv-nearest": "nearest", "opencv-bilinear": "bilinear"}[resize_type])
    d = int(round((first - size) / 2.))
    return full[d:d + size, d:d + size]
