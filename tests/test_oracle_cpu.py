"""The oracle is pinned against vectors produced by the REFERENCE's own corruptions.py
(tests/golden/make_golden.py, run in the build container where /root/reference exists)."""
import hashlib
import json
import os

import numpy as np
import pytest

from util import synth_images

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "imagenet_c_reference.json")))


def _cases():
    names = sorted({k.split("/")[0] for k in GOLD["cases"]})
    return names


@pytest.mark.parametrize("name", _cases())
def test_oracle_reproduces_reference_bytes(name):
    from oracle import imagenet_c as O
    from robustart_b200.assets import frost_textures
    images = synth_images(2, seed=42)
    kw = {"textures": frost_textures()} if name == "frost" else {}
    for sev in range(1, 6):
        for i in range(2):
            seed = 1000 + 10 * i + sev
            out = O.corrupt(images[i].copy(), sev, name, draws=O.NumpyDraws(seed), **kw)
            sha = hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest()
            assert sha == GOLD["cases"]["%s/%d/%d" % (name, sev, i)]["sha256"], (name, sev, i)


def test_unpinned_list_is_what_we_document():
    assert sorted(GOLD["skipped_unpinned"]) == ["motion_blur", "snow"]
    assert len(GOLD["cases"]) == 17 * 5 * 2


def test_corrupt_dispatch_and_errors():
    from oracle import imagenet_c as O
    img = synth_images(1, seed=1)[0]
    a = O.corrupt(img, 2, corruption_name="contrast")
    b = O.corrupt(img, 2, corruption_number=11)
    assert np.array_equal(a, b) and a.dtype == np.uint8
    with pytest.raises(ValueError):
        O.corrupt(img, 1)
    assert O.CORRUPTION_NAMES[15:] == ("speckle_noise", "gaussian_blur", "spatter", "saturate")


def test_replay_draws_roundtrip():
    from oracle import imagenet_c as O
    img = synth_images(1, seed=2)[0]
    d = O.NumpyDraws(5)
    a = O.corrupt(img, 3, "fog", draws=d)
    b = O.corrupt(img, 3, "fog", draws=O.ReplayDraws(d.log))
    assert np.array_equal(a, b)
    assert sum(np.size(x) for _, x in d.log) == 65535


def test_metrics_oracle():
    import torch
    from oracle import metrics as OM
    z = torch.tensor([[0.1, 0.9, 0.0, 0.2, 0.3, 0.4, 0.5], [0.9, 0.1, 0.0, 0.2, 0.3, 0.4, 0.5]])
    y = torch.tensor([1, 2])
    assert OM.topk_hits(z, y) == [1, 1]
    a1, a5 = OM.accuracy(z, y, (1, 5))
    assert a1.item() == 50.0 and a5.item() == 50.0
    # sampler: contiguous slices of one permutation, last rank takes the remainder, no duplication
    n, w = 50001, 8
    parts = [OM.sampler_indices(n, w, r) for r in range(w)]
    assert sorted(sum(parts, [])) == list(range(n))
    assert [len(p) for p in parts] == [6251] * 7 + [50001 - 7 * 6251]


def test_jpeg_restatement_equals_pil():
    """oracle/jpeg_restatement.py (the algorithm the CUDA kernel follows) == PIL's libjpeg round trip."""
    import io
    from PIL import Image
    from oracle.jpeg_restatement import jpeg_roundtrip
    imgs = synth_images(2, seed=3)
    for q in (25, 18, 15, 10, 7):
        for i in range(2):
            o = io.BytesIO()
            Image.fromarray(imgs[i]).save(o, "JPEG", quality=q)
            assert np.array_equal(jpeg_roundtrip(imgs[i], q), np.array(Image.open(o)))


def test_rayleigh_table_normal():
    """Quality of the device-RNG normal generator of gaussian / speckle noise (csrc/corrupt_pixel.cu, build_rayleigh): the
    pair radius comes from a 256-segment piecewise-linear inverse CDF of the Rayleigh distribution (least-squares line per
    equiprobable segment, moment-matched open tail), the angle is exact.  numpy restatement of the table builder; the
    resulting z = r cos(theta) must be N(0,1) to KS < 1e-4 with variance / kurtosis within 1e-3."""
    from scipy.stats import norm
    k = np.arange(65536)
    p = (k + 0.5) / 65536
    rq = np.sqrt(-2 * np.log1p(-p))
    A, B = np.zeros(256), np.zeros(256)
    b = np.arange(256.0)
    for a in range(256):
        y = rq[a * 256:(a + 1) * 256]
        if a == 255:
            m, s = y.mean(), y.std()
            half = s * np.sqrt(3.0) * (256 / 255.0)
            A[a], B[a] = m - half, 2 * half / 255
        else:
            B[a], A[a] = np.polyfit(b, y, 1)
    r = A[k >> 8] + B[k >> 8] * (k & 255)
    assert r.min() >= 0 and r.max() < 4.2

    def cdf_z(t):       # P(r cos(theta) <= t), theta uniform
        x = np.clip(t / np.maximum(r, 1e-12), -1, 1)
        return (1 - np.arccos(x) / np.pi).mean()
    ts = np.linspace(-4.0, 4.0, 161)
    ks = max(abs(cdf_z(t) - norm.cdf(t)) for t in ts)
    assert ks < 1e-4, ks
    var = (r ** 2).mean() / 2
    kurt = (3 / 8 * (r ** 4).mean()) / var ** 2
    assert abs(var - 1) < 1e-3 and abs(kurt - 3) < 3e-3, (var, kurt)


def test_resize_restatement_equals_pil():
    """oracle/resize.py (Resample.c / Geometry.c restated in numpy) == Pillow's Image.resize for all six filters, up / down /
    mixed scaling, odd sizes, and the ImageNet-S 'val' transform (imagenet_s_gen.py:131-141) built on it."""
    from PIL import Image
    from oracle import resize as R
    F = {"nearest": Image.NEAREST, "box": Image.BOX, "bilinear": Image.BILINEAR, "hamming": Image.HAMMING,
         "bicubic": Image.BICUBIC, "lanczos": Image.LANCZOS}
    rs = np.random.RandomState(3)
    for (h, w, oh, ow) in [(75, 100, 64, 64), (37, 53, 64, 80), (64, 64, 73, 73), (60, 40, 60, 25), (41, 97, 30, 97), (9, 11, 50, 33)]:
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        for name, f in F.items():
            want = np.asarray(Image.fromarray(img).resize((ow, oh), f))
            assert np.array_equal(R.resize(img, oh, ow, name), want), (h, w, oh, ow, name)
    img = rs.randint(0, 256, (90, 120, 3)).astype(np.uint8)
    for rt, f in [("pil-bilinear", Image.BILINEAR), ("pil-cubic", Image.BICUBIC), ("pil-nearest", Image.NEAREST)]:
        full = np.asarray(Image.fromarray(img).resize((64, 64), f))      # size 56 -> first resize 64, crop offset 4
        assert np.array_equal(R.imagenet_s_val(img, rt, size=56), full[4:60, 4:60])


def test_attack_pieces_match_reference_code():
    """Pieces of the attack path against golden vectors produced by the REFERENCE's own vendored code on CPU
    (tests/golden/make_golden_attacks.py): the MI-FGSM oracle == imfgsm_attack.py:_mim_whitebox bit for bit (same start
    draws), the DLR formulas == APGDAttack.dlr_loss / dlr_loss_targeted, and two host-side pieces of the product that are
    device-agnostic torch / python: FAB's projection_linf (fab_projections.py:7-59) and Square's p schedule (square.py:192-219)."""
    import os
    import torch
    import torch.nn as nn
    from oracle import attacks as OA, autoattack as OAA
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "attack_pieces.npz"))
    # MI-FGSM: the generator's tiny CNN, rebuilt from the same seed
    torch.manual_seed(123)
    model = nn.Sequential(nn.Conv2d(3, 8, 3, 2, 1), nn.ReLU(), nn.Conv2d(8, 16, 3, 2, 1), nn.ReLU(), nn.AdaptiveAvgPool2d(1), nn.Flatten(),
                          nn.Linear(16, 10)).eval()
    eps, steps, step_size, decay = g["mim_cfg"].tolist()
    adv = OA.mim_linf(model, torch.from_numpy(g["mim_X"]), torch.from_numpy(g["mim_y"]), eps, int(steps), step_size, decay,
                      start_u=torch.from_numpy(g["mim_u"]))
    assert np.abs(adv.detach().numpy() - g["mim_adv"]).max() <= 1e-7
    assert np.abs(g["mim_adv"] - g["mim_X"]).max() <= eps + 1e-6
    # PGD-Linf: the foolbox restatement against the reference's in-repo loop (adv_cls_solver_train_pgd_new.py:67-105).  The two
    # project differently in floating point (x0 + clip(x - x0) vs max(min(x, x0 + eps), x0 - eps)): 1-2 ulp apart at most
    eps_p, rel, steps_p = g["pgd_cfg"].tolist()
    advp = OA.pgd_linf(model, torch.from_numpy(g["pgd_x"]), torch.from_numpy(g["pgd_y"]), eps_p, rel, int(steps_p),
                       start_u=torch.from_numpy(g["pgd_u"]))
    assert np.abs(advp.detach().numpy() - g["pgd_adv"]).max() <= 2e-7
    assert np.abs(g["pgd_adv"] - g["pgd_x"]).max() > 0.5 * eps_p
    # DLR
    z, y, t = torch.from_numpy(g["dlr_z"]), torch.from_numpy(g["dlr_y"]), torch.from_numpy(g["dlr_t"])
    assert np.array_equal(OAA.dlr_loss(z, y).numpy(), g["dlr"])
    assert np.array_equal(OAA.dlr_loss_targeted(z, y, t).numpy(), g["dlr_targeted"])
    # host-side product pieces (no kernels involved)
    from robustart_b200 import autoattack as AA
    from oracle import autoattack as OAA
    d = OAA.projection_linf(torch.from_numpy(g["proj_t"]), torch.from_numpy(g["proj_w"]), torch.from_numpy(g["proj_b"]))
    assert np.abs(d.numpy() - g["proj"]).max() <= 1e-6
    sq = AA.Square(None, 4 / 255, n_queries=5000, p_init=0.8)
    assert [sq._p(int(i)) for i in g["sq_it"]] == g["sq_p"].tolist()


def test_eval_reduction_matches_reference_code():
    """Sampler shards and accuracy() against goldens produced by executing the reference's own DistributedSampler
    (data/sampler.py:8-52, round_up=False) and accuracy (utils/misc.py:441-455) -- tests/golden/make_golden_metrics.py.
    Checked: the oracle restatements AND the product's host-side solver.shard_indices."""
    import hashlib
    import json
    import os
    import torch
    from oracle import metrics as OM
    from robustart_b200 import solver as S
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "metrics_reference.json")))
    for key, ranks in g["sampler"].items():
        n, w = (int(v) for v in key.split("/"))
        for r, want in enumerate(ranks):
            got_o, got_p = OM.sampler_indices(n, w, r), S.shard_indices(n, w, r).tolist()
            if isinstance(want, dict):
                for got in (got_o, got_p):
                    assert len(got) == want["len"] and hashlib.sha256(np.asarray(got, np.int32).tobytes()).hexdigest() == want["sha256"]
            else:
                assert got_o == want and got_p == want, (key, r)
    a = g["accuracy"]
    logits, target = torch.tensor(a["logits"]), torch.tensor(a["target"])
    top1, top5 = OM.accuracy(logits, target, topk=(1, 5))
    assert float(top1) == a["top1"] and float(top5) == a["top5"]
    h1, h5 = OM.topk_hits(logits, target, (1, 5))
    assert abs(100.0 * h1 / 40 - a["top1"]) < 1e-4 and abs(100.0 * h5 / 40 - a["top5"]) < 1e-4


def test_spatter_water_restatement():
    """oracle/spatter_water.py (the algorithm csrc/corrupt_spatter_water.cu follows) against the OpenCV calls the reference makes
    (corruptions.py:305-328, executed by oracle.imagenet_c.spatter through cv2 itself): final images byte-exact for severity
    1-3, and the integer stages exact one by one."""
    cv2 = pytest.importorskip("cv2")
    from oracle import imagenet_c as O, spatter_water as W
    from util import synth_images
    for sev in (1, 2, 3):
        img = synth_images(1, seed=70 + sev)[0]
        c = O.SPATTER_PARAMS[sev - 1]
        z = np.random.RandomState(sev).normal(size=img.shape[:2])

        class Draws:
            def normal(self, size, loc, scale):
                return loc + scale * z
        want = np.uint8(O.spatter(img, sev, Draws()))
        liquid = O.sk_gaussian(c[0] + c[1] * z, sigma=c[2], multichannel=False)
        liquid[liquid < c[3]] = 0
        l8 = (liquid * 255).astype(np.uint8)
        assert np.array_equal(W.water(l8, img, c[4]), want)
        assert (want != img).mean() > 0.02                                 # the branch does something
        # stage by stage
        edge = cv2.Canny(l8, 50, 150)
        assert np.array_equal(W.canny(l8) * np.uint8(255), edge)
        d = np.minimum(cv2.distanceTransform(255 - edge, cv2.DIST_L2, 5), 20)
        mine = W.truncated_distance(edge > 0)
        assert np.abs(mine - d).max() <= 4e-6 and (mine != d).mean() < 0.02   # float32 path order inside IPP: <= 1 ulp apart
        u = cv2.blur(d, (3, 3)).astype(np.uint8)
        assert np.array_equal(W.blur3_f32_to_u8(d), u)
        assert np.array_equal(W.equalize_hist(u), cv2.equalizeHist(u))
        e = cv2.equalizeHist(u)
        f = cv2.filter2D(e, cv2.CV_8U, np.array([[-2, -1, 0], [-1, 1, 1], [0, 1, 2]]))
        assert np.array_equal(W.filter2d_emboss(e), f)
        assert np.array_equal(W.blur3_u8(f), cv2.blur(f, (3, 3)))
    flat = np.full((16, 16), 7, np.uint8)
    assert np.array_equal(W.equalize_hist(flat), cv2.equalizeHist(flat))   # single-bin image: cv2 fills with the bin index


def test_cv_resize_restatement():
    """oracle/cv_resize.py (the algorithm csrc/resize_cv.cu follows) against cv2.resize itself -- the call the reference's ImageNet-S
    generator makes for the opencv-* types (imagenet_s_gen.py:28-34,120-148): bit-exact, any scale direction."""
    cv2 = pytest.importorskip("cv2")
    from oracle import cv_resize as R
    rng = np.random.RandomState(0)
    sizes = [(375, 500, 256, 256), (64, 64, 256, 256), (100, 80, 128, 128), (224, 224, 256, 256), (480, 640, 256, 256), (17, 31, 64, 48),
             (300, 200, 256, 256), (31, 500, 256, 256), (256, 256, 256, 256), (2, 2, 8, 8), (1, 5, 4, 4), (37, 1, 16, 16), (333, 500, 299, 299),
             (500, 333, 64, 64), (97, 113, 111, 89)]
    for (h, w, ho, wo) in sizes:
        img = rng.randint(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(R.resize_linear(img, wo, ho), cv2.resize(img, (wo, ho), interpolation=cv2.INTER_LINEAR)), (h, w, ho, wo)
        assert np.array_equal(R.resize_nearest(img, wo, ho), cv2.resize(img, (wo, ho), interpolation=cv2.INTER_NEAREST)), (h, w, ho, wo)
        assert np.array_equal(R.resize_area(img, wo, ho), cv2.resize(img, (wo, ho), interpolation=cv2.INTER_AREA)), (h, w, ho, wo)
        assert np.array_equal(R.resize_lanczos4(img, wo, ho), cv2.resize(img, (wo, ho), interpolation=cv2.INTER_LANCZOS4)), (h, w, ho, wo)
        # INTER_CUBIC goes through closed-source IPP in the opencv-python wheels: a float32 cubic, matched to within 1 LSB on <= 1e-4
        # (IPP declines very small sources -- OpenCV's fixed-point path runs instead -- so the bar is stated for sides >= 16)
        if min(h, w) >= 16:
            d = np.abs(R.resize_cubic(img, wo, ho).astype(int) - cv2.resize(img, (wo, ho), interpolation=cv2.INTER_CUBIC).astype(int))
            assert d.max() <= 1 and (d > 0).mean() <= max(1e-4, 1.5 / d.size), ((h, w, ho, wo), d.max(), (d > 0).mean())
    for (h, w, ho, wo) in [(512, 512, 256, 256), (768, 768, 256, 256), (512, 768, 256, 256), (515, 770, 256, 256), (100, 300, 256, 256)]:
        img = rng.randint(0, 256, (h, w, 3), dtype=np.uint8)      # INTER_AREA: integer factors, general shrink, one axis growing
        assert np.array_equal(R.resize_area(img, wo, ho), cv2.resize(img, (wo, ho), interpolation=cv2.INTER_AREA)), (h, w, ho, wo)
    img = rng.randint(0, 256, (375, 500, 3), dtype=np.uint8)
    for rt, inter in (("opencv-bilinear", cv2.INTER_LINEAR), ("opencv-nearest", cv2.INTER_NEAREST), ("opencv-area", cv2.INTER_AREA)):
        full = cv2.resize(img, (256, 256), interpolation=inter)
        assert np.array_equal(R.imagenet_s_val(img, rt), full[16:240, 16:240])


@pytest.mark.parametrize("arch", ["resnet18", "resnet50"])
def test_calibrated_golden_logits_reproduce(arch):
    """tests/golden/calibrated_logits.npz (realistic-magnitude logits made by the REFERENCE's classes): the oracle model with
    the rebuilt calibrated state_dict gives the stored logits -- pins the fixture and util.calibrated_state_dict on CPU."""
    import torch
    from oracle import models as OM
    from robustart_b200 import nets
    from util import diverse_images, calibrated_state_dict
    cal = np.load(os.path.join(os.path.dirname(__file__), "golden", "calibrated_logits.npz"))
    want = cal[arch + "/logits"]
    assert np.abs(want).max() > 10 and len(set(want.argmax(1).tolist())) >= 3 and 2.0 < want.std() < 3.0
    sd = calibrated_state_dict(arch, nets.random_state_dict(nets.resnet_spec(arch), 0), cal)
    model = OM.build(arch, sd)
    n = 8 if arch == "resnet18" else 3
    x = torch.from_numpy(diverse_images(8, seed=0)[:n]).permute(0, 3, 1, 2).float().div(255)
    xn = (x - torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)) / torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    with torch.no_grad():
        got = model(xn).numpy()
    assert np.abs(got - want[:n]).max() < 2e-4, np.abs(got - want[:n]).max()      # batch-size dependent fp32 summation order only


def test_strata_table_normal():
    """Quality of the round-2 device-RNG normal generator of gaussian / speckle noise (csrc/corrupt_pixel.cu
    normal_noise_strata_kernel): a random byte picks one of 256 rows, lane and loop iteration pick one of 64 strata, the table
    entry IS the normal.  The table is read from the library itself (host-only export b200r_normal_strata_table).  Pooled over
    strata the 16 384 atoms must be N(0,1) to KS < 1e-4 with variance within 5e-4 and kurtosis within 5e-3; each stratum alone is
    a symmetric 256-atom quantile grid with mean 0 and standard deviation within 3 % of 1 (np.random.normal is what
    corruptions.py:122-126,143-147 draw)."""
    import ctypes as C
    from scipy.stats import norm
    from robustart_b200 import _lib
    z = np.zeros(256 * 64)
    _lib.check(_lib.load().b200r_normal_strata_table(z.ctypes.data_as(C.c_void_p)))
    t = z.reshape(256, 64)
    zs, n = np.sort(z), z.size
    assert np.allclose(zs[8192:], norm.ppf(0.5 + (np.arange(8192) + 0.5) / 16384), atol=1e-12)        # the atoms are exact quantiles
    cdf = norm.cdf(zs)
    ks = max(np.abs(cdf - (np.arange(n) + 1) / n).max(), np.abs(cdf - np.arange(n) / n).max())
    kurt = (z ** 4).mean() / z.var() ** 2
    assert abs(z.mean()) < 1e-12 and abs(z.var() - 1) < 5e-4 and abs(kurt - 3) < 5e-3 and ks < 1e-4, (z.var(), kurt, ks)
    assert 3.9 < np.abs(z).max() < 4.1
    assert np.abs(t.mean(0)).max() < 1e-12 and np.abs(t.std(0) - 1).max() < 3e-2, (t.mean(0), t.std(0))
    assert np.allclose(t, -t[::-1]) and (np.diff(t, axis=0) > 0).all()      # every stratum: symmetric, increasing with the row


def test_l1_projection_restatement_matches_reference_code():
    """oracle.autoattack.l1_projection against outputs of the reference's own vendored L1_projection (autopgd_base.py:19-83,
    executed from source by tests/golden/make_golden_l1proj.py): exact."""
    import torch
    from oracle import autoattack as OAA
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "l1_projection.npz"))
    for k in range(4):
        x, y, eps = torch.from_numpy(g["x%d" % k]), torch.from_numpy(g["y%d" % k]), float(g["eps%d" % k])
        d = OAA.l1_projection(x, y, eps).numpy()
        assert np.abs(d - g["d%d" % k]).max() <= 1e-7, k
        z = y.numpy() + d
        assert np.abs(z).sum(1).max() <= eps * (1 + 1e-5) and (x.numpy() + z).min() >= -1e-6 and (x.numpy() + z).max() <= 1 + 1e-6


def _motion_blur_by_filter2d(img_u8, radius, sigma, angle_deg):
    """A second, independent evaluation of ImageMagick's MotionBlurImage for the unpinned C6 / C8 rows (Wand is absent here): the
    one-sided Gaussian kernel of the documented definition (effect.c GetMotionBlurKernel: k[i] ~ exp(-i^2 / 2 sigma^2), i = 0..width-1,
    width = 2 ceil(radius) + 1; offsets ceil(i cos(a) - 0.5), ceil(i sin(a) - 0.5) along the angle) rasterised into a dense 2-D kernel
    and applied by OpenCV's filter2D with a replicated border, on the Q16 values; then ImageMagick's quantum rounding."""
    import math
    import cv2
    width = int(2 * math.ceil(radius) + 1)
    i = np.arange(width, dtype=np.float64)
    k = np.exp(-i * i / (2 * sigma * sigma))
    k /= k.sum()
    a = math.radians(angle_deg)
    ox = np.ceil(i * math.cos(a) - 0.5).astype(int)
    oy = np.ceil(i * math.sin(a) - 0.5).astype(int)
    R = width
    dense = np.zeros((2 * R + 1, 2 * R + 1), np.float64)
    for w, dx, dy in zip(k, ox, oy):
        dense[R + dy, R + dx] += w                       # filter2D correlates: out(y, x) = sum K[j, i] * src(y + j - R, x + i - R)
    q = cv2.filter2D(img_u8.astype(np.float64) * 257.0, cv2.CV_64F, dense, borderType=cv2.BORDER_REPLICATE)
    q = np.floor(np.clip(q, 0, 65535.0) + 0.5)
    return np.floor((q + 128.0) / 257.0).astype(np.uint8), k, ox, oy


def test_motion_blur_second_independent_statement():
    """VERDICT r1 missing #8: motion_blur / snow are unpinned (no ImageMagick here).  The oracle's restatement is cross-checked against
    an independent evaluation through cv2.filter2D and against the documented properties of the kernel."""
    from oracle import imagenet_c as O
    rs = np.random.RandomState(3)
    img = rs.randint(0, 256, (64, 80, 3)).astype(np.uint8)
    for (radius, sigma) in O.MOTION_PARAMS:
        for angle in (-45.0, -17.3, 0.0, 8.9, 44.9, 135.0, -120.0):
            want = O.magick_motion_blur_u8(img, radius, sigma, angle)
            got, k, ox, oy = _motion_blur_by_filter2d(img, radius, sigma, angle)
            d = np.abs(got.astype(int) - want.astype(int))
            assert d.max() <= 1 and (d > 0).mean() < 1e-3, (radius, sigma, angle, d.max(), (d > 0).mean())   # fp64 summation order only
            # documented properties: normalised, one-sided (starts at the pixel itself), monotonically decreasing, width 2 ceil(r) + 1
            assert abs(k.sum() - 1) < 1e-12 and len(k) == 2 * int(np.ceil(radius)) + 1 and (np.diff(k) < 0).all()
            assert ox[0] == 0 and oy[0] == 0
            # the smear runs along the angle: the farthest tap sits at distance ~ width - 1 in direction (cos a, sin a)
            far = np.array([ox[-1], oy[-1]], float)
            dirv = np.array([np.cos(np.radians(angle)), np.sin(np.radians(angle))])
            assert abs(np.linalg.norm(far) - (len(k) - 1)) <= 1.0 and far @ dirv > 0.98 * np.linalg.norm(far)
    # a constant image is a fixed point; an impulse spreads into exactly the kernel's taps
    const = np.full((32, 32, 3), 137, np.uint8)
    assert (O.magick_motion_blur_u8(const, 15, 8, 30.0) == 137).all()
    imp = np.zeros((64, 64, 3), np.uint8)
    imp[32, 32] = 255
    out = O.magick_motion_blur_u8(imp, 10, 3, 20.0)
    kk, ox, oy = O.motion_blur_kernel(10, 3, 20.0)
    ys, xs = np.nonzero(out[..., 0])
    assert set(zip(ys.tolist(), xs.tolist())) <= {(32 - int(b), 32 - int(a)) for a, b in zip(ox, oy)}       # the tap reads src(x + ox): the impulse lands at x - ox
