"""b200r_resize_u8 (csrc/resize.cu) == Pillow's Image.resize, bit for bit: the ImageNet-S `pil-*` resize types
(RobustART/noise/utils/imagenet_s_gen.py:19-26,120-141) and the eval transform's Resize + CenterCrop
(imagenet_dataloader.py:74-80).  Pillow itself is the reference here (same library call the reference makes)."""
import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu
F = {"nearest": Image.NEAREST, "box": Image.BOX, "bilinear": Image.BILINEAR, "hamming": Image.HAMMING,
     "bicubic": Image.BICUBIC, "lanczos": Image.LANCZOS}


def _pil(batch, oh, ow, f):
    return np.stack([np.asarray(Image.fromarray(im).resize((ow, oh), f)) for im in batch])


@pytest.mark.parametrize("shape", [(3, 500, 375, 256, 256), (2, 37, 53, 64, 80), (2, 224, 224, 256, 256), (1, 300, 200, 300, 256),
                                   (2, 123, 457, 256, 457), (2, 64, 64, 7, 9), (1, 5, 7, 50, 33), (2, 96, 96, 96, 96)])
def test_resize_equals_pil(cuda, shape):
    from robustart_b200 import ops
    n, h, w, oh, ow = shape
    rs = np.random.RandomState(sum(shape))
    batch = rs.randint(0, 256, (n, h, w, 3)).astype(np.uint8)
    batch[0, : h // 2] = np.linspace(0, 255, w).astype(np.uint8)[None, :, None]     # smooth ramp + hard edge: ringing filters clip
    d = torch.from_numpy(batch).to(cuda)
    for name, f in F.items():
        want = _pil(batch, oh, ow, f)
        got = ops.resize_u8(d, (oh, ow), name).cpu().numpy()
        assert np.array_equal(got, want), (shape, name, int(np.abs(got.astype(int) - want.astype(int)).max()))
        # a crop window of the resized image: only it is computed
        y0, x0, ch, cw = oh // 5, ow // 7, max(1, oh // 2), max(1, ow // 3)
        got = ops.resize_u8(d, (oh, ow), name, crop=(y0, x0, ch, cw)).cpu().numpy()
        assert np.array_equal(got, want[:, y0:y0 + ch, x0:x0 + cw]), (shape, name, "crop")
    assert torch.equal(d.cpu(), torch.from_numpy(batch))


def test_eval_transform_and_imagenet_s_plugin(cuda, tmp_path):
    """Resize(256) + CenterCrop(224) as torchvision does on PIL images, and AddNoise('imagenet-s') on a file path against the
    reference's ImageTransfer arithmetic (PIL resize to 256x256, crop at int(round(16.0)))."""
    from robustart_b200 import ops
    from RobustART.noise import AddNoise
    rs = np.random.RandomState(5)
    img = rs.randint(0, 256, (2, 375, 500, 3)).astype(np.uint8)
    got = ops.resize_center_crop_u8(torch.from_numpy(img).to(cuda), 256, 224).cpu().numpy()
    for i in range(2):
        pil = Image.fromarray(img[i]).resize((int(256 * 500 / 375), 256), Image.BILINEAR)      # torchvision F.resize, short side 256
        x0 = int(round((pil.size[0] - 224) / 2.0))
        assert np.array_equal(got[i], np.asarray(pil.crop((x0, 16, x0 + 224, 16 + 224))))
    path = str(tmp_path / "a.png")
    Image.fromarray(img[0]).save(path)
    for rt, f in [("pil-bilinear", Image.BILINEAR), ("pil-nearest", Image.NEAREST), ("pil-box", Image.BOX), ("pil-hamming", Image.HAMMING),
                  ("pil-cubic", Image.BICUBIC), ("pil-lanczos", Image.LANCZOS)]:
        gen = AddNoise("imagenet-s")
        gen.set_config(resize_type=rt)
        out = gen.add_noise(path)
        want = np.asarray(Image.open(path).convert("RGB").resize((256, 256), f).crop((16, 16, 240, 240)))
        assert out.shape == (224, 224, 3) and np.array_equal(out, want), rt
    gen = AddNoise("imagenet-s")
    gen.set_config(decoder_type="ffmpeg")            # the one decoder without an implementation fails loudly
    with pytest.raises(NotImplementedError):
        gen.add_noise(path)


def test_resize_table_cache_is_bounded(cuda):
    """More distinct (in, out) size pairs than the coefficient-table cache holds (a file-backed run resizes every image at its native
    size): old entries are evicted and rebuilt on demand, results unchanged."""
    from robustart_b200 import ops
    torch.manual_seed(0)
    first = torch.randint(0, 256, (1, 9, 8, 3), dtype=torch.uint8, device=cuda)
    want = ops.resize_u8(first, (5, 4), "bilinear").clone()
    for i in range(560):                                  # 560 new widths -> 560 new horizontal tables (+ the shared vertical one)
        img = torch.randint(0, 256, (1, 9, 10 + i, 3), dtype=torch.uint8, device=cuda)
        out = ops.resize_u8(img, (5, 4), "bilinear")
        assert out.shape == (1, 5, 4, 3)
    assert torch.equal(ops.resize_u8(first, (5, 4), "bilinear"), want)
