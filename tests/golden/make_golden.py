"""Generate golden vectors by running the REFERENCE's own Python (import shims, this container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.json / *.npz

/root/reference does not exist on the GPU box, so the vectors are committed.  What is real and what is
stubbed when importing /root/reference/RobustART/noise/utils/imagenet_c/corruptions.py:
  real     : numpy, scipy.ndimage.zoom/map_coordinates, cv2, PIL  (the same calls the reference makes)
  shimmed  : skimage (absent)  -> filters.gaussian / util.random_noise / color.* provided by the
             restatements in oracle/imagenet_c.py (skimage 0.17.2 algorithms);
             wand (absent)     -> motion_blur and snow cannot run: NO golden vector (parity unpinned);
             pkg_resources / frost files (absent) -> cv2.imread patched to return the procedural textures;
             np.float_ -> np.float64 (NumPy 2 removed the alias).
So C1,C2,C4,C7,C9,C10,C12,C13,C14,C15,C16 are pinned against the reference's own arithmetic and
C3,C5,C11,C17,C18,C19 against the reference's control flow around the restated skimage calls.
"""
import hashlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"


def install_shims():
    from oracle import imagenet_c as O
    np.float_ = np.float64
    sk = types.ModuleType("skimage")
    sk.filters = types.ModuleType("skimage.filters")
    sk.util = types.ModuleType("skimage.util")
    sk.color = types.ModuleType("skimage.color")

    def gaussian(image, sigma=1, output=None, mode="nearest", cval=0, multichannel=None, preserve_range=False,
                 truncate=4.0):
        return O.sk_gaussian(image, sigma, mode=mode, truncate=truncate, multichannel=multichannel)

    def random_noise(image, mode="s&p", amount=0.05, salt_vs_pepper=0.5, **kw):
        assert mode == "s&p"
        out = image.copy()
        flipped = np.random.choice([True, False], size=image.shape, p=[amount, 1 - amount])
        salted = np.random.choice([True, False], size=image.shape, p=[salt_vs_pepper, 1 - salt_vs_pepper])
        out[flipped & salted] = 1
        out[flipped & ~salted] = 0
        return out

    sk.filters.gaussian = gaussian
    sk.util.random_noise = random_noise
    sk.color.rgb2hsv = O.sk_rgb2hsv
    sk.color.hsv2rgb = O.sk_hsv2rgb
    sys.modules.update({"skimage": sk, "skimage.filters": sk.filters, "skimage.util": sk.util, "skimage.color": sk.color})
    wand = types.ModuleType("wand")
    for sub in ("image", "api", "color"):
        m = types.ModuleType("wand." + sub)
        setattr(wand, sub, m)
        sys.modules["wand." + sub] = m
    wand.image.Image = object

    class _Lib:
        class MagickMotionBlurImage:
            argtypes = None
    wand.api.library = _Lib
    sys.modules["wand"] = wand
    if "pkg_resources" not in sys.modules:
        try:
            import pkg_resources  # noqa: F401
        except Exception:
            pr = types.ModuleType("pkg_resources")
            pr.resource_filename = lambda name, path: path
            sys.modules["pkg_resources"] = pr
    try:
        import scipy.ndimage.interpolation  # noqa: F401
    except Exception:
        import scipy.ndimage as ndi
        m = types.ModuleType("scipy.ndimage.interpolation")
        m.map_coordinates = ndi.map_coordinates
        sys.modules["scipy.ndimage.interpolation"] = m


def load_reference_corruptions():
    import importlib.util
    path = os.path.join(REF, "RobustART/noise/utils/imagenet_c/corruptions.py")
    spec = importlib.util.spec_from_file_location("ref_corruptions", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_corruptions"] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    from PIL import Image
    import cv2
    from util import synth_images
    from robustart_b200.assets import frost_textures
    from oracle import imagenet_c as O
    install_shims()
    ref = load_reference_corruptions()
    tex = frost_textures()
    names = ["frost%d" % (i + 1) for i in range(6)]
    real_imread = cv2.imread

    def fake_imread(fn, *a):
        for i, nm in enumerate(names):
            if nm in os.path.basename(fn):
                return tex[i][..., ::-1].copy()   # cv2 returns BGR
        return real_imread(fn, *a)

    ref.cv2.imread = fake_imread
    images = synth_images(2, seed=42)
    golden = {"input": "tests/util.py synth_images(2, seed=42)", "seed_rule": "np.random.seed(1000 + 10*image + severity)",
              "cases": {}}
    skipped = {}
    for name in O.CORRUPTION_NAMES:
        fn = getattr(ref, name)
        for sev in range(1, 6):
            for i in range(2):
                key = "%s/%d/%d" % (name, sev, i)
                seed = 1000 + 10 * i + sev
                np.random.seed(seed)
                try:
                    out = np.uint8(fn(Image.fromarray(images[i]), sev))
                except Exception as e:  # wand-dependent functions
                    skipped[name] = repr(e)[:120]
                    continue
                kw = {"textures": tex} if name == "frost" else {}
                mine = O.corrupt(images[i].copy(), sev, name, draws=O.NumpyDraws(seed), **kw)
                d = np.abs(out.astype(int) - mine.astype(int))
                golden["cases"][key] = {"sha256": hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest(),
                                        "mean": float(out.mean()), "oracle_max_abs_diff_at_generation": int(d.max())}
                if d.max() != 0:
                    print("MISMATCH", key, int(d.max()), float((d > 0).mean()))
    golden["skipped_unpinned"] = skipped
    json.dump(golden, open(os.path.join(HERE, "imagenet_c_reference.json"), "w"), indent=1, sort_keys=True)
    print("cases:", len(golden["cases"]), "skipped:", skipped)


if __name__ == "__main__":
    main()
