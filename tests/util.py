"""Shared test helpers: synthetic images and conversion of oracle draw logs to ext_noise buffers."""
import numpy as np

H = W = 224


def synth_images(n, seed=0, h=H, w=W):
    """Half i.i.d. uint8 noise, half smooth fields (SURVEY 8d synthetic inputs)."""
    from scipy import ndimage as ndi
    rs = np.random.RandomState(seed)
    out = np.empty((n, h, w, 3), np.uint8)
    for i in range(n):
        if i % 2 == 0:
            out[i] = rs.randint(0, 256, size=(h, w, 3))
        else:
            f = ndi.gaussian_filter(rs.rand(h, w, 3), [8, 8, 0])
            out[i] = np.uint8(255 * (f - f.min()) / np.ptp(f))
    # make sure extremes and flat regions occur
    out[0, :8, :8] = 0
    out[0, 8:16, :8] = 255
    if n > 1:
        out[1, :16, :16] = 128
    return out


def ext_from_log(name, log):
    """Flatten one image's oracle draw log into the ext_noise layout of b200r_corrupt_u8."""
    arrs = [np.asarray(a, dtype=np.float64).ravel() for _, a in log]
    if not arrs:
        return np.zeros(0, np.float32)
    return np.concatenate(arrs).astype(np.float32)


def oracle_batch(images, name, severity, seed0=100, **kw):
    """Run the oracle per image with its own RandomState; returns (outputs uint8, ext float32)."""
    from oracle import imagenet_c as O
    outs, exts = [], []
    for i in range(images.shape[0]):
        d = O.NumpyDraws(seed0 + i)
        outs.append(O.corrupt(images[i].copy(), severity, name, draws=d, **kw))
        exts.append(ext_from_log(name, d.log))
    return np.stack(outs), np.concatenate(exts) if exts else np.zeros(0, np.float32)


HEAD_KEYS = {"resnet18": "fc", "resnet50": "fc", "vit_b16_224": "head", "mixer_b16_224": "head",
             "mobilenet_v2": "classifier.1", "efficientnet_b0": "fc"}


def diverse_images(n, seed=0, h=H, w=W):
    """Structurally different synthetic images (tinted gradients, stripes, blobs, noise at several contrasts), so that a
    random-weight classifier sees different feature statistics per image and predicts DIFFERENT classes -- the
    realistic-magnitude goldens (tests/golden/calibrated_logits.npz) are made on these."""
    from scipy import ndimage as ndi
    rs = np.random.RandomState(1000 + seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    out = np.empty((n, h, w, 3), np.uint8)
    for i in range(n):
        tint = rs.rand(3) * 0.8 + 0.1
        kind = i % 4
        if kind == 0:      # tinted linear gradient + fine noise
            ang = rs.rand() * np.pi
            g = (np.cos(ang) * xx / w + np.sin(ang) * yy / h)
            f = tint[None, None, :] * (0.3 + 0.7 * g[..., None]) + 0.15 * rs.rand(h, w, 3)
        elif kind == 1:    # stripes of random period and phase per channel
            f = np.stack([0.5 + 0.5 * np.sin(xx * (0.05 + 0.4 * rs.rand()) + yy * (0.3 * rs.rand()) + 6 * rs.rand()) for _ in range(3)], -1)
            f = f * tint[None, None, :] + 0.1
        elif kind == 2:    # smooth blobs, strong colour cast
            f = ndi.gaussian_filter(rs.rand(h, w, 3), [12, 12, 0])
            f = (f - f.min()) / np.ptp(f) * tint[None, None, :] * 1.2
        else:              # i.i.d. noise at a random contrast / brightness
            f = 0.5 + (rs.rand(h, w, 3) - 0.5) * (0.2 + 0.8 * rs.rand()) + (tint[None, None, :] - 0.5) * 0.6
        out[i] = np.uint8(np.clip(f, 0, 1) * 255)
    return out


def calibrated_state_dict(arch, sd, cal):
    """Apply the head calibration stored in tests/golden/calibrated_logits.npz (made from the REFERENCE's classes by
    tests/golden/make_golden_calibrated.py): head.weight *= s (float32), head.bias = stored vector, and for the BatchNorm families the
    running_mean / running_var of every BN layer = the statistics of the calibration batch (what training leaves there; with
    unrelated random statistics a deep random ReLU net maps every image to nearly the same feature).  Everything else is the
    usual deterministic synthetic state_dict.  Gives logits with std ~2.5, |max| 10-20 and distinct top-1 per image."""
    import torch
    k = HEAD_KEYS[arch]
    sd = dict(sd)
    pre = arch + "/bn/"
    for name in cal.files if hasattr(cal, "files") else cal:          # BatchNorm running statistics of the calibration batch
        if name.startswith(pre):
            sd[name[len(pre):]] = torch.from_numpy(np.asarray(cal[name], dtype=np.float32).copy())
    sd[k + ".weight"] = sd[k + ".weight"].float() * float(np.float32(cal[arch + "/scale"]))
    sd[k + ".bias"] = torch.from_numpy(np.asarray(cal[arch + "/bias"], dtype=np.float32).copy())
    return sd
