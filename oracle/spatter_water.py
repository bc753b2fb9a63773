"""Plain-numpy restatement of the OpenCV chain in spatter's "water" branch -- TEST INFRASTRUCTURE ONLY (see oracle/imagenet_c.py).

RobustART/noise/utils/imagenet_c/corruptions.py:305-328 calls cv2.Canny, cv2.distanceTransform, cv2.threshold, cv2.blur,
cv2.equalizeHist, cv2.filter2D, cv2.blur, cv2.cvtColor.  OpenCV is a third-party dependency of the reference (opencv-python,
requirements.txt); it IS installed in this container (4.13.0, IPP build), so this restatement is pinned against the very calls
the reference makes (tests/test_oracle_cpu.py::test_spatter_water_restatement: byte-exact final images, exact Canny / blur /
equalizeHist / filter2D stages; the float32 chamfer distances agree with IPP's raster scan to 1 ulp on < 1 % of the pixels, which
never survives the uint8 truncation that follows).  It exists because the CUDA kernel (csrc/corrupt_spatter_water.cu) needs an
algorithm, not a library call, to follow; oracle/imagenet_c.py:spatter keeps calling cv2 itself.
"""
import numpy as np

R = 20          # cv2.threshold(dist, 20, 20, THRESH_TRUNC): nothing further than 20 pixels can matter


def canny(img, lo=50, hi=150):
    """cv2.Canny(img, lo, hi) (aperture 3, L1 gradient) -> boolean edge map."""
    I = img.astype(np.int32)
    h, w = img.shape
    P = np.pad(I, 1, mode="edge")                                   # Sobel with BORDER_REPLICATE
    gx = (P[:-2, 2:] + 2 * P[1:-1, 2:] + P[2:, 2:]) - (P[:-2, :-2] + 2 * P[1:-1, :-2] + P[2:, :-2])
    gy = (P[2:, :-2] + 2 * P[2:, 1:-1] + P[2:, 2:]) - (P[:-2, :-2] + 2 * P[:-2, 1:-1] + P[:-2, 2:])
    mag = np.abs(gx) + np.abs(gy)
    M = np.pad(mag, 1)                                              # magnitude 0 outside the image
    ax, ay = np.abs(gx), np.abs(gy) << 15
    tg22x = ax * 13573                                              # tan(22.5 deg) in 15-bit fixed point
    tg67x = tg22x + (ax << 16)
    c = M[1:-1, 1:-1]
    horiz, vert = ay < tg22x, ay > tg67x
    s = np.where((gx ^ gy) < 0, -1, 1)
    yy, xx = np.mgrid[0:h, 0:w]
    keep = (horiz & (c > M[1:-1, :-2]) & (c >= M[1:-1, 2:])) | (vert & (c > M[:-2, 1:-1]) & (c >= M[2:, 1:-1])) | \
           (~horiz & ~vert & (c > M[yy, xx + 1 - s]) & (c > M[yy + 2, xx + 1 + s]))
    keep &= c > lo
    strong = keep & (c > hi)
    weak = keep & ~strong
    out = strong.copy()
    while True:                                                     # hysteresis: 8-connected growth of strong into weak
        Pd = np.pad(out, 1)
        nb = np.zeros_like(out)
        for dy in range(3):
            for dx in range(3):
                nb |= Pd[dy:dy + h, dx:dx + w]
        new = out | (weak & nb)
        if (new == out).all():
            return out
        out = new


def chamfer_table():
    """5x5 chamfer metric of DIST_L2 (1, 1.4, 2.1969) for |dx|, |dy| <= R, moves added in float32: axial, diagonal, knight."""
    a, b, c = np.float32(1.0), np.float32(1.4), np.float32(2.1969)
    T = np.zeros((R + 1, R + 1), np.float32)
    for dy in range(R + 1):
        for dx in range(R + 1):
            mx, mn = max(dx, dy), min(dx, dy)
            knight, axial, diag = (mn, mx - 2 * mn, 0) if mx >= 2 * mn else (mx - mn, 0, 2 * mn - mx)
            s = np.float32(0)
            for step, cnt in ((a, axial), (b, diag), (c, knight)):
                for _ in range(cnt):
                    s = np.float32(s + step)
            T[dy, dx] = s
    return T


def truncated_distance(edge):
    """min(cv2.distanceTransform(255 - 255*edge, DIST_L2, 5), 20) as float32."""
    h, w = edge.shape
    T = chamfer_table()
    P = np.pad(edge, R)
    best = np.full((h, w), np.float32(R), np.float32)
    for dy in range(-R, R + 1):
        for dx in range(-R, R + 1):
            t = T[abs(dy), abs(dx)]
            if t < R:
                z = P[R + dy:R + dy + h, R + dx:R + dx + w]
                best = np.where(z & (t < best), t, best)
    return best


def _win3(P, h, w):
    return sum(P[i:i + h, j:j + w] for i in range(3) for j in range(3))


def blur3_f32_to_u8(dist):
    """np.uint8(cv2.blur(float32 plane, (3,3))): double sums times (1/9), BORDER_REFLECT_101."""
    h, w = dist.shape
    return (_win3(np.pad(dist.astype(np.float64), 1, mode="reflect"), h, w) * (1.0 / 9)).astype(np.float32).astype(np.uint8)


def equalize_hist(u):
    hist = np.bincount(u.ravel(), minlength=256)
    i0 = int(np.nonzero(hist)[0][0])
    if hist[i0] == u.size:
        return np.full_like(u, i0)
    scale = np.float32(255.0) / np.float32(u.size - hist[i0])
    lut = np.zeros(256, np.uint8)
    s = 0
    for k in range(i0 + 1, 256):
        s += int(hist[k])
        lut[k] = min(255, int(np.rint(np.float32(s) * scale)))
    return lut[u]


def filter2d_emboss(u):
    """cv2.filter2D(u, CV_8U, [[-2,-1,0],[-1,1,1],[0,1,2]])."""
    ker = ((-2, -1, 0), (-1, 1, 1), (0, 1, 2))
    h, w = u.shape
    P = np.pad(u.astype(np.int32), 1, mode="reflect")
    return np.clip(sum(ker[i][j] * P[i:i + h, j:j + w] for i in range(3) for j in range(3)), 0, 255).astype(np.uint8)


def blur3_u8(u):
    h, w = u.shape
    return ((_win3(np.pad(u.astype(np.int32), 1, mode="reflect"), h, w) * 2 + 9) // 18).astype(np.uint8)


def water(l8, x_u8, c4):
    """corruptions.py:305-328 from the uint8 liquid layer on; returns uint8 [h, w, 3] (the np.uint8() of corrupt() included)."""
    dist = truncated_distance(canny(l8, 50, 150))
    b = blur3_u8(filter2d_emboss(equalize_hist(blur3_f32_to_u8(dist)))).astype(np.float32)
    m = l8.astype(np.float32) * b
    with np.errstate(invalid="ignore", divide="ignore"):
        m = m / m.max() * np.float32(c4)
        col = np.array([175 / 255., 238 / 255., 238 / 255.], np.float32)
        x = x_u8.astype(np.float32) / np.float32(255.)
        out = np.clip(x + m[..., None] * col, 0, 1) * 255
        return np.nan_to_num(out, nan=0.0).astype(np.uint8)
