"""Split-precision one-launch stem (csrc/stem_pool_sm100.cu, SPLIT = true): conv1 7x7/s2 + bn1 + relu + maxpool from raw uint8
pixels against an fp64 statement of resnet_official.py:221-227,330-334 after ToTensor + Normalize
(imagenet_dataloader.py:78-79).  The A operand is exact (pixel / 256), the weights are an fp16 hi/lo pair: the bar is fp32-class."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda", 0)


def _reference(img, wt, s, b, sel):
    from robustart_b200 import ops
    mean = torch.tensor(ops.IMAGENET_MEAN, device=img.device, dtype=torch.float64).view(1, 3, 1, 1)
    std = torch.tensor(ops.IMAGENET_STD, device=img.device, dtype=torch.float64).view(1, 3, 1, 1)
    x = (img[sel].permute(0, 3, 1, 2).double() / 255.0 - mean) / std
    conv = F.conv2d(x, wt.double(), None, 2, 3) * s.double().view(1, -1, 1, 1) + b.double().view(1, -1, 1, 1)
    return F.max_pool2d(torch.relu(conv), 3, 2, 1).permute(0, 2, 3, 1)


@pytest.mark.parametrize("n,h,w", [(2, 224, 224), (48, 224, 224), (3, 64, 64), (1, 40, 248), (5, 8, 8), (2, 100, 72)])
def test_stem_pool_split_one_launch(cuda, n, h, w):
    from robustart_b200 import ops
    torch.manual_seed(n * 11 + h + w)
    img = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device=cuda)
    wt = torch.randn(64, 3, 7, 7, device=cuda) / 147 ** 0.5
    s, b = torch.rand(64, device=cuda) + 0.5, torch.randn(64, device=cuda) * 0.3
    s[::5] *= -1                                  # negative BN scales: the scale must act before the ReLU
    sel = list(range(n)) if n <= 8 else [0, 1, n // 3, n // 2, n - 2, n - 1]
    ref = _reference(img, wt, s, b, sel)
    wp, osc = ops.stem_pool_split_prepare(wt, s, b, cuda)
    y = ops.stem_pool_u8_split(img, wp, osc)
    assert y.shape == (2, n, h // 4, w // 4, 64)
    got = ops.from_planes(y).double()
    assert torch.isfinite(got).all()
    err = (got[sel] - ref).abs().max().item()
    # 22-bit weights, exact pixels, fp32 accumulation over K = 224 (the tensor core truncates: ~28 steps x 2^-24 of the partial
    # sums, whose terms reach ~4 x the result) and a 22-bit split on the way out
    assert err <= 3e-6 * max(1.0, ref.abs().max().item()), err
    # the two-launch path (three MMAs per product, normalised hi/lo pixels) agrees to fp32 noise
    if w % 16 == 0:
        w3 = ops.to_planes(ops.pack_stem_weight(wt).contiguous(), False)
        two = ops.from_planes(ops.maxpool3x3s2(ops.stem_conv7x7_u8(img, w3, s, b, act="relu"))).double()
        assert (got - two).abs().max().item() < 2e-5


def test_stem_pool_split_range(cuda):
    """Tiny and huge weights, large bias, constant images: the power-of-two range centring keeps the hi/lo pair exact."""
    from robustart_b200 import ops
    torch.manual_seed(5)
    n, h, w = 2, 32, 32
    for wscale, bscale in ((1e-4, 0.0), (30.0, 5.0), (0.05, 40.0)):
        img = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device=cuda)
        img[1] = 255
        wt = torch.randn(64, 3, 7, 7, device=cuda) * wscale
        s, b = torch.rand(64, device=cuda) + 0.5, torch.randn(64, device=cuda) * bscale
        ref = _reference(img, wt, s, b, [0, 1])
        wp, osc = ops.stem_pool_split_prepare(wt, s, b, cuda)
        got = ops.from_planes(ops.stem_pool_u8_split(img, wp, osc)).double()
        mag = max(ref.abs().max().item(), 1e-30)
        assert (got - ref).abs().max().item() <= 4e-6 * mag, (wscale, bscale, (got - ref).abs().max().item(), mag)


def test_resnet_forward_uses_split_stem(cuda):
    """nets.ResNet (split precision) takes the one-launch stem for uint8 input and agrees with the two-launch path."""
    from robustart_b200 import nets
    from util import synth_images
    images = torch.from_numpy(synth_images(4, seed=3)).to(cuda)
    model = nets.build_model("resnet18", device=cuda, seed=0)
    assert model.fused_stem_pool
    a = model(images)
    model.fused_stem_pool = False
    b = model(images)
    assert (a - b).abs().max().item() < 2e-5 * max(1.0, b.abs().max().item())


@pytest.mark.parametrize("n,h,w", [(2, 224, 224), (3, 64, 64), (1, 40, 248), (5, 8, 8), (2, 100, 72), (30, 224, 224)])
def test_stem_pool_f32_with_codes(cuda, n, h, w):
    """b200r_stem_pool_f32_split (MODE 2: float image in [0,1] as an fp16 hi/lo pair, three MMAs per product, arg-max codes out): pooled
    planes against fp64, codes against the window position of the fp64 maximum (first maximum in scan order, 0xF where not positive),
    and both against the two-launch path (b200r_stem_conv7x7_f32 + b200r_maxpool3x3s2_nhwc_codes) it replaces."""
    from robustart_b200 import ops
    torch.manual_seed(n * 13 + h + w)
    x01 = torch.rand(n, 3, h, w, device=cuda)
    wt = torch.randn(64, 3, 7, 7, device=cuda) / 147 ** 0.5
    s, b = torch.rand(64, device=cuda) + 0.5, torch.randn(64, device=cuda) * 0.3
    s[::5] *= -1
    mean = torch.tensor(ops.IMAGENET_MEAN, device=cuda, dtype=torch.float64).view(1, 3, 1, 1)
    std = torch.tensor(ops.IMAGENET_STD, device=cuda, dtype=torch.float64).view(1, 3, 1, 1)
    sel = list(range(n)) if n <= 8 else [0, 1, n // 2, n - 1]
    conv = F.conv2d((x01[sel].double() - mean) / std, wt.double(), None, 2, 3) * s.double().view(1, -1, 1, 1) + b.double().view(1, -1, 1, 1)
    act = torch.relu(conv)
    ref, idx = F.max_pool2d(act, 3, 2, 1, return_indices=True)
    wp, osc = ops.stem_pool_split_prepare(wt, s, b, cuda, f32_input=True)
    y, codes = ops.stem_pool_f32_split(x01, wp, osc)
    assert y.shape == (2, n, h // 4, w // 4, 64) and codes.shape == (n, h // 4, w // 4, 64)
    got = ops.from_planes(y).double()[sel]
    err = (got - ref.permute(0, 2, 3, 1)).abs().max().item()
    assert err <= 3e-6 * max(1.0, ref.abs().max().item()), err
    # codes: flat index of the fp64 arg-max -> (ky, kx) inside the window; positions where the maximum is 0 route nothing
    ho, wo = h // 2, w // 2
    iy, ix = idx // wo, idx % wo
    py = torch.arange(h // 4, device=cuda).view(1, 1, -1, 1)
    px = torch.arange(w // 4, device=cuda).view(1, 1, 1, -1)
    want = ((iy - (2 * py - 1)) * 3 + (ix - (2 * px - 1))).permute(0, 2, 3, 1)
    want = torch.where(ref.permute(0, 2, 3, 1) > 0, want, torch.full_like(want, 15))
    c = codes[sel].long()
    agree = (c == want).double().mean().item()
    assert agree > 0.9995, agree                     # near-ties between window entries may resolve differently in fp32
    assert ((c <= 8) | (c == 15)).all()
    # the two-launch path it replaces
    w3 = ops.to_planes(ops.pack_stem_weight(wt).contiguous(), False)
    if w % 16 == 0:
        y2, c2 = ops.maxpool3x3s2_codes(ops.stem_conv7x7_f32(x01, w3, s, b, act="relu"))
        assert (ops.from_planes(y2) - ops.from_planes(y)).abs().max().item() < 2e-5
        assert (c2 == codes).double().mean().item() > 0.9995
