"""Assets the reference expects next to its code but does not ship.

corruptions.py:251-257 reads frost/frost{1,2,3}.png and frost{4,5,6}.jpg; the files are absent from
the reference repository.  By default deterministic procedural textures are used; a user who owns the
original ImageNet-C frost files can point set_frost_dir() at them.
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np

_frost_dir: Optional[str] = os.environ.get("B200R_FROST_DIR")
_cache: Optional[List[np.ndarray]] = None
FROST_FILES = ("frost1.png", "frost2.png", "frost3.png", "frost4.jpg", "frost5.jpg", "frost6.jpg")


def set_frost_dir(path: Optional[str]):
    global _frost_dir, _cache
    _frost_dir, _cache = path, None


def _procedural(seed=1234, size=(480, 640)) -> List[np.ndarray]:
    """Blue-white crystalline noise: a smooth random field with bright ridges, six variants."""
    from scipy import ndimage as ndi
    rs = np.random.RandomState(seed)
    out = []
    for k in range(6):
        base = ndi.gaussian_filter(rs.rand(*size), 6 + k)
        fine = ndi.gaussian_filter(rs.rand(*size), 1.5)
        t = 0.6 * (base - base.min()) / np.ptp(base) + 0.4 * (fine - fine.min()) / np.ptp(fine)
        ridge = np.abs(ndi.sobel(base))
        t = np.clip(t + 2.0 * ridge / ridge.max(), 0, 1)
        rgb = np.stack([0.75 * t + 0.15, 0.85 * t + 0.12, 0.95 * t + 0.05], -1)
        out.append(np.uint8(np.clip(rgb, 0, 1) * 255))
    return out


def frost_textures() -> List[np.ndarray]:
    """Six uint8 RGB textures [th, tw, 3] (th, tw > 224)."""
    global _cache
    if _cache is None:
        if _frost_dir:
            from PIL import Image
            _cache = [np.array(Image.open(os.path.join(_frost_dir, f)).convert("RGB")) for f in FROST_FILES]
        else:
            _cache = _procedural()
    return _cache
