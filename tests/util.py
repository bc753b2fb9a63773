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
