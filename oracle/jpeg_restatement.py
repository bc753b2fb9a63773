"""Integer restatement of the libjpeg(-turbo) baseline round trip that PIL performs for
jpeg_compression (corruptions.py:375-382: Image.save(format='JPEG', quality=q) then Image.open) --
TEST INFRASTRUCTURE ONLY.  The CUDA kernel in robustart_b200/csrc/corrupt_codec.cu follows this file step
by step; tests/test_oracle_cpu.py checks this file against PIL itself (so the algorithm is pinned by the
very library call the reference makes).

Pipeline (libjpeg-turbo defaults as used by Pillow: 4:2:0, JDCT_ISLOW, fancy upsampling, baseline tables):
  jccolor.c rgb_ycc_convert -> jcsample.c h2v2_downsample -> jfdctint.c jpeg_fdct_islow (samples - 128)
  -> jcdctmgr.c quantize (divisor = q*8, round half away) -> [entropy coding is lossless: skipped]
  -> jidctint.c jpeg_idct_islow (dequantise inside) -> jdsample.c h2v2_fancy_upsample
  -> jdcolor.c ycc_rgb_convert.
"""
import numpy as np

STD_LUMA = np.array([16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
                     14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                     49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99], np.int64).reshape(8, 8)
STD_CHROMA = np.array([17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
                       47, 66, 99, 99, 99, 99, 99, 99] + [99] * 32, np.int64).reshape(8, 8)


def quant_tables(quality):
    """jcparam.c jpeg_quality_scaling + jpeg_add_quant_table(force_baseline=TRUE)."""
    quality = max(1, min(100, quality))
    scale = 5000 // quality if quality < 50 else 200 - quality * 2
    out = []
    for t in (STD_LUMA, STD_CHROMA):
        q = (t * scale + 50) // 100
        out.append(np.clip(q, 1, 255))
    return out


def _fix(x):
    return int(x * 65536 + 0.5)


def rgb_to_ycc(rgb):
    r, g, b = [rgb[..., i].astype(np.int64) for i in range(3)]
    half = 1 << 15
    off = 128 << 16
    y = (_fix(0.29900) * r + _fix(0.58700) * g + _fix(0.11400) * b + half) >> 16
    cb = (-_fix(0.16874) * r - _fix(0.33126) * g + _fix(0.50000) * b + off + half - 1) >> 16
    cr = (_fix(0.50000) * r - _fix(0.41869) * g - _fix(0.08131) * b + off + half - 1) >> 16
    return y, cb, cr


def h2v2_downsample(c):
    h, w = c.shape
    s = c[0::2, 0::2] + c[0::2, 1::2] + c[1::2, 0::2] + c[1::2, 1::2]
    bias = np.tile(np.array([1, 2]), w // 4 + 1)[: w // 2]      # 1,2,1,2,... along the row
    return (s + bias[None, :]) >> 2


CONST_BITS, PASS1_BITS = 13, 2
F = {k: int(v * (1 << CONST_BITS) + 0.5) for k, v in dict(
    f0_298631336=0.298631336, f0_390180644=0.390180644, f0_541196100=0.541196100, f0_765366865=0.765366865,
    f0_899976223=0.899976223, f1_175875602=1.175875602, f1_501321110=1.501321110, f1_847759065=1.847759065,
    f1_961570560=1.961570560, f2_053119869=2.053119869, f2_562915447=2.562915447, f3_072711026=3.072711026).items()}


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _fdct_1d(d, first):
    """one pass of jpeg_fdct_islow over the last axis; d: [..., 8] int64."""
    t0, t7 = d[..., 0] + d[..., 7], d[..., 0] - d[..., 7]
    t1, t6 = d[..., 1] + d[..., 6], d[..., 1] - d[..., 6]
    t2, t5 = d[..., 2] + d[..., 5], d[..., 2] - d[..., 5]
    t3, t4 = d[..., 3] + d[..., 4], d[..., 3] - d[..., 4]
    t10, t13 = t0 + t3, t0 - t3
    t11, t12 = t1 + t2, t1 - t2
    o = np.empty_like(d)
    if first:
        o[..., 0] = (t10 + t11) << PASS1_BITS
        o[..., 4] = (t10 - t11) << PASS1_BITS
        n = CONST_BITS - PASS1_BITS
    else:
        o[..., 0] = _descale(t10 + t11, PASS1_BITS)
        o[..., 4] = _descale(t10 - t11, PASS1_BITS)
        n = CONST_BITS + PASS1_BITS
    z1 = (t12 + t13) * F["f0_541196100"]
    o[..., 2] = _descale(z1 + t13 * F["f0_765366865"], n)
    o[..., 6] = _descale(z1 + t12 * (-F["f1_847759065"]), n)
    z1, z2, z3, z4 = t4 + t7, t5 + t6, t4 + t6, t5 + t7
    z5 = (z3 + z4) * F["f1_175875602"]
    t4 = t4 * F["f0_298631336"]; t5 = t5 * F["f2_053119869"]; t6 = t6 * F["f3_072711026"]; t7 = t7 * F["f1_501321110"]
    z1 = z1 * (-F["f0_899976223"]); z2 = z2 * (-F["f2_562915447"])
    z3 = z3 * (-F["f1_961570560"]) + z5; z4 = z4 * (-F["f0_390180644"]) + z5
    o[..., 7] = _descale(t4 + z1 + z3, n)
    o[..., 5] = _descale(t5 + z2 + z4, n)
    o[..., 3] = _descale(t6 + z2 + z3, n)
    o[..., 1] = _descale(t7 + z1 + z4, n)
    return o


def fdct_islow(blocks):
    """blocks [..., 8, 8] of (sample - 128); returns coefficients scaled by 8."""
    d = _fdct_1d(blocks.astype(np.int64), True)                  # rows
    d = _fdct_1d(d.swapaxes(-1, -2), False).swapaxes(-1, -2)     # columns
    return d


def quantize(coef, q):
    q8 = q.astype(np.int64) << 3
    a = np.abs(coef)
    return np.sign(coef) * ((a + (q8 >> 1)) // q8)


def _idct_1d(w, first):
    """jpeg_idct_islow pass over the last axis.  first: column pass (results scaled by 2^PASS1_BITS);
    second: row pass with final descale by CONST_BITS+PASS1_BITS+3."""
    z2, z3 = w[..., 2], w[..., 6]
    z1 = (z2 + z3) * F["f0_541196100"]
    t2 = z1 + z3 * (-F["f1_847759065"])
    t3 = z1 + z2 * F["f0_765366865"]
    z2, z3 = w[..., 0], w[..., 4]
    t0 = (z2 + z3) << CONST_BITS
    t1 = (z2 - z3) << CONST_BITS
    t10, t13, t11, t12 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    t0, t1, t2, t3 = w[..., 7], w[..., 5], w[..., 3], w[..., 1]
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * F["f1_175875602"]
    t0 = t0 * F["f0_298631336"]; t1 = t1 * F["f2_053119869"]; t2 = t2 * F["f3_072711026"]; t3 = t3 * F["f1_501321110"]
    z1 = z1 * (-F["f0_899976223"]); z2 = z2 * (-F["f2_562915447"])
    z3 = z3 * (-F["f1_961570560"]) + z5; z4 = z4 * (-F["f0_390180644"]) + z5
    t0 = t0 + z1 + z3; t1 = t1 + z2 + z4; t2 = t2 + z2 + z3; t3 = t3 + z1 + z4
    n = CONST_BITS - PASS1_BITS if first else CONST_BITS + PASS1_BITS + 3
    o = np.empty_like(w)
    o[..., 0] = _descale(t10 + t3, n); o[..., 7] = _descale(t10 - t3, n)
    o[..., 1] = _descale(t11 + t2, n); o[..., 6] = _descale(t11 - t2, n)
    o[..., 2] = _descale(t12 + t1, n); o[..., 5] = _descale(t12 - t1, n)
    o[..., 3] = _descale(t13 + t0, n); o[..., 4] = _descale(t13 - t0, n)
    return o


def idct_islow(qcoef, q):
    w = (qcoef * q.astype(np.int64))                               # dequantise
    w = _idct_1d(w.swapaxes(-1, -2), True).swapaxes(-1, -2)        # pass 1: columns
    w = _idct_1d(w, False)                                         # pass 2: rows
    return np.clip(w + 128, 0, 255)                                # range_limit (centered at 128)


def h2v2_fancy_upsample(c):
    """jdsample.c h2v2_fancy_upsample: triangle filter, 3/4 nearer + 1/4 further in each axis."""
    h, w = c.shape
    c = c.astype(np.int64)
    up = np.concatenate([c[:1], c[:-1]], 0)       # row above (replicated at the top)
    dn = np.concatenate([c[1:], c[-1:]], 0)       # row below (replicated at the bottom)
    out = np.empty((2 * h, 2 * w), np.int64)
    for v, other in ((0, up), (1, dn)):
        colsum = c * 3 + other                    # thiscolsum
        last = np.concatenate([colsum[:, :1], colsum[:, :-1]], 1)
        nxt = np.concatenate([colsum[:, 1:], colsum[:, -1:]], 1)
        even = (colsum * 3 + last + 8) >> 4
        odd = (colsum * 3 + nxt + 7) >> 4
        # first and last columns: special cases in libjpeg ((colsum*4 + 8)>>4 and (colsum*4 + 7)>>4)
        even[:, 0] = (colsum[:, 0] * 4 + 8) >> 4
        odd[:, -1] = (colsum[:, -1] * 4 + 7) >> 4
        out[v::2, 0::2] = even
        out[v::2, 1::2] = odd
    return out


def ycc_to_rgb(y, cb, cr):
    half = 1 << 15
    x_cb, x_cr = cb - 128, cr - 128
    r = y + ((_fix(1.40200) * x_cr + half) >> 16)
    g = y + ((-_fix(0.34414) * x_cb + half - _fix(0.71414) * x_cr) >> 16)
    b = y + ((_fix(1.77200) * x_cb + half) >> 16)
    return np.stack([np.clip(r, 0, 255), np.clip(g, 0, 255), np.clip(b, 0, 255)], -1).astype(np.uint8)


def _blocks(p):
    h, w = p.shape
    return p.reshape(h // 8, 8, w // 8, 8).swapaxes(1, 2)


def _unblocks(b):
    nh, nw = b.shape[:2]
    return b.swapaxes(1, 2).reshape(nh * 8, nw * 8)


def jpeg_roundtrip(rgb, quality):
    """rgb uint8 [H,W,3] with H, W multiples of 16."""
    ql, qc = quant_tables(quality)
    y, cb, cr = rgb_to_ycc(rgb)
    planes = [(y, ql), (h2v2_downsample(cb), qc), (h2v2_downsample(cr), qc)]
    rec = []
    for p, q in planes:
        coef = fdct_islow(_blocks(p) - 128)
        rec.append(_unblocks(idct_islow(quantize(coef, q), q)))
    return ycc_to_rgb(rec[0], h2v2_fancy_upsample(rec[1]), h2v2_fancy_upsample(rec[2]))
