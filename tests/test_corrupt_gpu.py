"""GPU parity of the ImageNet-C kernels against the CPU oracle (shared-draw mode) -- through the C-ABI.

Tolerance (stated, per BASELINE north_star "stated fp tolerance for corrupted pixels"): the kernels
compute in fp32, the reference in fp64, then both truncate to uint8, so a value landing within 1e-5 of
an integer may fall on the other side: |gpu - oracle| <= 1 LSB everywhere, and on at most MISMATCH_FRAC
of the pixels.  Integer pipelines (shot with shared draws, impulse, pixelate) are bit-exact.
"""
import numpy as np
import pytest
import torch

from util import synth_images, oracle_batch

pytestmark = pytest.mark.gpu

EXACT = {"shot_noise", "impulse_noise"}
# fraction of pixels allowed to differ by exactly 1 LSB
# brightness / saturate: the HSV round trip often lands EXACTLY on an integer in real arithmetic
# (e.g. saturate c=2: p*255 = 2*min - max), so fp64 rounding noise alone decides k vs k-1 in the
# reference; fp32 decides differently on a large share of such pixels.  Still <= 1 LSB.
# Limits = ~1.5x the largest fraction measured on a B200 over the five severities of these seeds (round 2: brightness 0.094,
# saturate 0.149, frost 0.028, everything else < 1e-4)
MISMATCH_FRAC = {"default": 0.002, "brightness": 0.15, "saturate": 0.22, "fog": 0.002, "frost": 0.05}
PIXEL_FAMILY = ["gaussian_noise", "shot_noise", "impulse_noise", "speckle_noise", "brightness", "saturate",
                "contrast", "frost", "fog"]


def _run(cuda, name, sev, images, ext):
    from robustart_b200 import ops
    d = torch.from_numpy(images).to(cuda)
    e = torch.from_numpy(ext).to(cuda) if ext.size else None
    out = ops.corrupt_u8(d, name, sev, ext_noise=e)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _compare(name, got, want):
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    if name in EXACT:
        assert diff.max() == 0, f"{name}: {np.count_nonzero(diff)} mismatching bytes, max {diff.max()}"
        return
    frac = np.count_nonzero(diff) / diff.size
    assert diff.max() <= 1, f"{name}: max abs diff {diff.max()} (frac>1: {(diff > 1).mean():.2e})"
    lim = MISMATCH_FRAC.get(name, MISMATCH_FRAC["default"])
    assert frac <= lim, f"{name}: {frac:.4f} of pixels differ by 1 LSB (limit {lim})"


@pytest.mark.parametrize("name", PIXEL_FAMILY)
@pytest.mark.parametrize("sev", [1, 2, 3, 4, 5])
def test_pixel_family_matches_oracle(cuda, name, sev):
    images = synth_images(4, seed=sev)
    kw = {}
    if name == "frost":
        from robustart_b200.assets import frost_textures
        kw["textures"] = frost_textures()
    want, ext = oracle_batch(images, name, sev, **kw)
    got = _run(cuda, name, sev, images, ext)
    _compare(name, got, want)


STENCIL_FAMILY = ["gaussian_blur", "glass_blur", "defocus_blur", "zoom_blur", "motion_blur", "snow", "elastic_transform"]
# kernels with a hard threshold / rounding step inside (snow: layer < c3 -> 0; motion: Q16 round half up;
# elastic: OpenCV's 1/32-pixel coordinate quantisation and a displacement field scaled by up to 488 px):
# an fp32-vs-fp64 tie flips a whole quantisation step at isolated pixels, bounded by OUTLIER_FRAC.
# measured (round 2, these seeds): no pixel of snow / motion_blur / glass_blur off by more than 1 LSB; elastic 1.5e-4 (max 5 LSB);
# pixels off by exactly 1 LSB: defocus 0.028 (severity 1), snow 1e-3, elastic 8e-4, the rest <= 5e-4
OUTLIER_FRAC = {"snow": 2e-4, "motion_blur": 2e-4, "glass_blur": 2e-4, "elastic_transform": 5e-4}
STENCIL_MISMATCH = {"elastic_transform": 0.003, "defocus_blur": 0.045, "snow": 0.003}


@pytest.mark.parametrize("name", STENCIL_FAMILY)
@pytest.mark.parametrize("sev", [1, 2, 3, 4, 5])
def test_stencil_family_matches_oracle(cuda, name, sev):
    images = synth_images(3, seed=10 + sev)
    want, ext = oracle_batch(images, name, sev)
    got = _run(cuda, name, sev, images, ext)
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    outl = (diff > 1).mean()
    assert outl <= OUTLIER_FRAC.get(name, 0.0), f"{name} s{sev}: {outl:.2e} of pixels differ by more than 1 LSB (max {diff.max()})"
    frac = np.count_nonzero(diff) / diff.size
    assert frac <= STENCIL_MISMATCH.get(name, 0.002), f"{name} s{sev}: {frac:.4f} of pixels differ"
    # in-place call gives the same bytes
    from robustart_b200 import ops
    d = torch.from_numpy(images).to(cuda)
    e = torch.from_numpy(ext).to(cuda) if ext.size else None
    ops.corrupt_u8(d, name, sev, ext_noise=e, out=d)
    assert np.array_equal(d.cpu().numpy(), got)


@pytest.mark.parametrize("sev", [1, 2, 3, 4, 5])
def test_pixelate_bit_exact(cuda, sev):
    """PIL's fixed-point BOX resample is restated integer for integer: no tolerance."""
    images = synth_images(4, seed=20 + sev)
    want, _ = oracle_batch(images, "pixelate", sev)
    got = _run(cuda, "pixelate", sev, images, np.zeros(0, np.float32))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("sev", [1, 2, 3, 4, 5])
def test_jpeg_bit_exact(cuda, sev):
    """libjpeg's integer pipeline (colour, 4:2:0, islow DCT, quantiser, fancy upsampling) restated exactly:
    the GPU bytes equal PIL's save(JPEG)+open."""
    images = synth_images(4, seed=30 + sev)
    want, _ = oracle_batch(images, "jpeg_compression", sev)
    got = _run(cuda, "jpeg_compression", sev, images, np.zeros(0, np.float32))
    assert np.array_equal(got, want)
    from robustart_b200 import ops
    d = torch.from_numpy(images).to(cuda)
    ops.corrupt_u8(d, "jpeg_compression", sev, out=d)           # in place
    assert np.array_equal(d.cpu().numpy(), want)


@pytest.mark.parametrize("sev", [4, 5])
def test_spatter_mud_branch(cuda, sev):
    images = synth_images(3, seed=40 + sev)
    want, ext = oracle_batch(images, "spatter", sev)
    got = _run(cuda, "spatter", sev, images, ext)
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert (diff > 1).mean() <= 2e-3, ((diff > 1).mean(), diff.max())     # threshold ties flip isolated mask pixels
    assert np.count_nonzero(diff) / diff.size <= 0.03


@pytest.mark.parametrize("name", ["gaussian_noise", "speckle_noise", "shot_noise", "impulse_noise"])
def test_device_rng_distribution(cuda, name):
    """Device Philox mode: the corruption's first two moments per input level match the oracle's."""
    from robustart_b200 import ops
    sev = 3
    n = 8
    images = np.empty((n, 224, 224, 3), np.uint8)
    levels = [0, 32, 64, 96, 128, 160, 200, 255]
    for i, lv in enumerate(levels):
        images[i] = lv
    want, _ = oracle_batch(images, name, sev, seed0=7)
    got = ops.corrupt_u8(torch.from_numpy(images).to(cuda), name, sev, seed=1234).cpu().numpy()
    for i, lv in enumerate(levels):
        a, b = got[i].astype(np.float64), want[i].astype(np.float64)
        se = b.std() / np.sqrt(b.size) + 1e-9
        assert abs(a.mean() - b.mean()) < 6 * se + 0.35, (name, lv, a.mean(), b.mean())
        assert abs(a.std() - b.std()) < 0.02 * b.std() + 0.35, (name, lv, a.std(), b.std())


def test_device_rng_is_counter_based(cuda):
    """Same seed + image_offset => same bytes, independent of batch split."""
    from robustart_b200 import ops
    images = torch.from_numpy(synth_images(6, seed=3)).to(cuda)
    whole = ops.corrupt_u8(images, "gaussian_noise", 2, seed=99, image_offset=10)
    a = ops.corrupt_u8(images[:2].contiguous(), "gaussian_noise", 2, seed=99, image_offset=10)
    b = ops.corrupt_u8(images[2:].contiguous(), "gaussian_noise", 2, seed=99, image_offset=12)
    assert torch.equal(whole, torch.cat([a, b]))
    other = ops.corrupt_u8(images, "gaussian_noise", 2, seed=100, image_offset=10)
    assert not torch.equal(whole, other)


def test_full_batch_properties(cuda):
    """BASELINE config 2 size (N=256): deterministic kernels are idempotent under re-run, in-place ==
    out-of-place, and severity is monotone in distortion for gaussian noise."""
    from robustart_b200 import ops
    rs = np.random.RandomState(0)
    images = torch.from_numpy(rs.randint(0, 256, size=(256, 224, 224, 3), dtype=np.uint8)).to(cuda)
    for name in ["brightness", "contrast", "saturate"]:
        o1 = ops.corrupt_u8(images, name, 3)
        o2 = ops.corrupt_u8(images, name, 3)
        assert torch.equal(o1, o2)
        tmp = images.clone()
        ops.corrupt_u8(tmp, name, 3, out=tmp)
        assert torch.equal(tmp, o1)
    prev = 0.0
    for sev in range(1, 6):
        o = ops.corrupt_u8(images, "gaussian_noise", sev, seed=5)
        err = (o.float() - images.float()).abs().mean().item()
        assert err > prev
        prev = err


def test_argument_errors(cuda):
    from robustart_b200 import ops
    images = torch.zeros((1, 224, 224, 3), dtype=torch.uint8, device=cuda)
    with pytest.raises(ValueError):
        ops.corrupt_u8(images, "gaussian_noise", 0)     # reference silently treats 0 as severity 5
    with pytest.raises(ValueError):
        ops.corrupt_u8(images, "gaussian_noise", 6)
    with pytest.raises(KeyError):
        ops.corrupt_u8(images, "no_such_noise", 1)
    with pytest.raises(TypeError):
        ops.corrupt_u8(images.cpu(), "gaussian_noise", 1)  # no CPU fallback
    empty = torch.zeros((0, 224, 224, 3), dtype=torch.uint8, device=cuda)
    assert ops.corrupt_u8(empty, "gaussian_noise", 1).shape[0] == 0


def test_noise_table_first_use_during_capture_fails_loudly(cuda):
    """The gaussian / speckle quantile table of a noise scale is uploaded at its first use; inside a stream capture that must be a clear
    error (and must not poison the capture), and the same launch is capturable once the table exists."""
    from robustart_b200 import ops
    img = torch.randint(0, 256, (2, 32, 32, 3), dtype=torch.uint8, device=cuda)
    out = torch.empty_like(img)
    sev = 2
    # a scale no other test uses: speckle severity 2 on a fresh process may already be cached, so accept either outcome of the first
    # capture but require the message when it fails
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    failed = False
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            try:
                ops.corrupt_u8(img, "speckle_noise", sev, seed=1, out=out)
            except ValueError as e:                       # B200R_EINVAL: a refused request, nothing was launched
                failed = True
                assert "stream capture" in str(e)
    torch.cuda.current_stream().wait_stream(s)
    ops.corrupt_u8(img, "speckle_noise", sev, seed=1, out=out)       # eager: builds the table
    want = out.clone()
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g2, stream=s):
            ops.corrupt_u8(img, "speckle_noise", sev, seed=1, out=out)
    out.zero_()
    g2.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want)


@pytest.mark.parametrize("sev", [1, 3, 5])
def test_motion_blur_kernel_against_independent_filter2d_statement(cuda, sev):
    """C6 (unpinned: no ImageMagick here) -- the GPU kernel against the SECOND statement of MotionBlurImage (cv2.filter2D of the documented
    one-sided Gaussian kernel, tests/test_oracle_cpu.py::_motion_blur_by_filter2d), with the angle handed in: random image, impulse,
    constant image."""
    from robustart_b200 import ops
    from oracle import imagenet_c as O
    from test_oracle_cpu import _motion_blur_by_filter2d
    rs = np.random.RandomState(sev)
    images = rs.randint(0, 256, (3, 224, 224, 3)).astype(np.uint8)
    images[1] = 0
    images[1, 100, 120] = 255
    images[2] = 91
    u = np.array([0.15, 0.63611, 0.98889], np.float32)            # the kernel takes the reference's uniform draw: angle = -45 + 90 u
    angles = (-45.0 + 90.0 * u.astype(np.float64))
    got = ops.corrupt_u8(torch.from_numpy(images).to(cuda), "motion_blur", sev, ext_noise=torch.from_numpy(u).to(cuda)).cpu().numpy()
    radius, sigma = O.MOTION_PARAMS[sev - 1]
    for i in range(3):
        want = _motion_blur_by_filter2d(images[i], radius, sigma, float(angles[i]))[0]
        d = np.abs(got[i].astype(int) - want.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3, (i, d.max(), (d > 0).mean())
    assert (got[2] == 91).all()
