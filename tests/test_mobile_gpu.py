"""MobileNetV2 / EfficientNet-B0 on the B200 kernels vs golden logits from the REFERENCE's own classes
(tests/golden/make_golden_models.py --mobile) + unit checks of the depthwise / SE / im2col kernels."""
import os

import numpy as np
import pytest
import torch

from util import synth_images

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "mobile_logits.npz"))


@pytest.mark.parametrize("k,stride,c", [(3, 1, 32), (3, 2, 96), (5, 1, 240), (5, 2, 144)])
def test_depthwise_conv(cuda, k, stride, c):
    from robustart_b200 import ops
    torch.manual_seed(k * 100 + c)
    x = torch.randn(3, c, 28, 28, device=cuda)
    w = torch.randn(c, 1, k, k, device=cuda) * 0.2
    s, b = torch.rand(c, device=cuda) + 0.5, torch.randn(c, device=cuda)
    xp = ops.split_f32(x.permute(0, 2, 3, 1).contiguous())
    xm = ops.merge_f32(xp).permute(0, 3, 1, 2)
    y = ops.merge_f32(ops.dwconv_nhwc(xp, w.reshape(c, -1).t().contiguous(), s, b, k=k, stride=stride, pad=k // 2, act="swish"))
    ref = torch.nn.functional.conv2d(xm.double(), w.double(), stride=stride, padding=k // 2, groups=c)
    ref = ref * s.double().view(1, -1, 1, 1) + b.double().view(1, -1, 1, 1)
    ref = ref * torch.sigmoid(ref)
    assert (y.permute(0, 3, 1, 2).double() - ref).abs().max().item() < 1e-4


@pytest.mark.parametrize("k,stride,c,h,w", [(3, 1, 32, 112, 112), (3, 2, 96, 112, 112), (3, 1, 144, 56, 56), (3, 2, 144, 56, 56), (3, 1, 384, 14, 14),
                                            (3, 1, 960, 7, 7), (5, 1, 240, 28, 28), (5, 2, 672, 14, 14), (5, 1, 1152, 7, 7), (3, 1, 8, 15, 17),
                                            (5, 2, 24, 33, 19), (3, 2, 16, 9, 9)])
def test_depthwise_tile_kernel_geometries(cuda, k, stride, c, h, w, monkeypatch):
    """dwconv_tile_kernel (shared-memory tile, one load + one unpack per input pixel) on the mobile families' layer shapes, ragged
    tiles, channel-group remainders and odd sizes: against fp64 and against the strip kernel it replaces
    (mobilenet_v2.py:31-47; efficientnet.py:322-336)."""
    from robustart_b200 import ops
    torch.manual_seed(k * 1000 + c + h)
    n = 2
    x = torch.randn(n, c, h, w, device=cuda)
    wt = torch.randn(c, 1, k, k, device=cuda) * 0.2
    s, b = torch.rand(c, device=cuda) + 0.5, torch.randn(c, device=cuda)
    xp = ops.split_f32(x.permute(0, 2, 3, 1).contiguous())
    xm = ops.merge_f32(xp).permute(0, 3, 1, 2)
    wk = wt.reshape(c, -1).t().contiguous()
    ref = torch.nn.functional.conv2d(xm.double(), wt.double(), stride=stride, padding=k // 2, groups=c)
    ref = torch.clamp(ref * s.double().view(1, -1, 1, 1) + b.double().view(1, -1, 1, 1), 0, 6).permute(0, 2, 3, 1)
    monkeypatch.setenv("B200R_DW_TILE", "1")
    got = ops.merge_f32(ops.dwconv_nhwc(xp, wk, s, b, k=k, stride=stride, pad=k // 2, act="relu6"))
    assert got.shape == ref.shape
    assert (got.double() - ref).abs().max().item() < 1e-4
    monkeypatch.setenv("B200R_DW_TILE", "0")
    old = ops.merge_f32(ops.dwconv_nhwc(xp, wk, s, b, k=k, stride=stride, pad=k // 2, act="relu6"))
    assert (got - old).abs().max().item() < 2e-5


def test_se_scale_and_small_channel_gemm(cuda):
    from robustart_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(4, 14, 14, 96, device=cuda)
    s = torch.rand(4, 96, device=cuda)
    y = ops.merge_f32(ops.channel_scale(ops.split_f32(x), ops.split_f32(s)))
    ref = ops.merge_f32(ops.split_f32(x)) * ops.merge_f32(ops.split_f32(s)).view(4, 1, 1, 96)
    assert (y - ref).abs().max().item() < 1e-4     # the product is re-split to 16 significant bits
    # pointwise conv with K tail (cin = 24 -> zero filled to 64) and narrow / ragged outputs
    for cin, cout in [(24, 144), (16, 96), (96, 24), (32, 16), (144, 40), (8, 8)]:
        xx = torch.randn(2, 28, 28, cin, device=cuda)
        w = torch.randn(cout, cin, device=cuda) / cin ** 0.5
        got = ops.merge_f32(ops.conv2d_nhwc(ops.split_f32(xx), ops.split_f32(w.view(cout, 1, 1, cin).contiguous()), act="relu6"))
        ref = torch.clamp(xx.double().view(-1, cin) @ w.double().t(), 0, 6).view(2, 28, 28, cout)
        assert (got.double() - ref).abs().max().item() < 1e-4, (cin, cout)


@pytest.mark.parametrize("arch", ["mobilenet_v2", "efficientnet_b0"])
def test_mobile_logits_match_reference(cuda, arch):
    from robustart_b200 import nets
    model = nets.build_model(arch, device=cuda, seed=0)
    images = torch.from_numpy(synth_images(2, seed=11)).to(cuda)
    got = model(images).cpu().numpy()
    want = GOLD[arch]
    err = np.abs(got - want).max()
    assert err < 1e-3 and err / np.abs(want).max() < 5e-3, (arch, err, err / np.abs(want).max())
    assert (got.argmax(1) == want.argmax(1)).all()
    x01 = images.permute(0, 3, 1, 2).float().div(255).contiguous()
    assert np.abs(model(x01).cpu().numpy() - got).max() < 1e-5


@pytest.mark.parametrize("arch", ["mobilenet_v2", "efficientnet_b0"])
def test_native_input_grad_matches_autograd_twin(cuda, arch):
    """Input-gradient pass of the mobile families on the kernels (depthwise dgrad = the forward kernel on flipped taps, 1x1 dgrad GEMMs,
    ReLU6 / swish / sigmoid derivatives, squeeze-excite reduction, transposed image stem) against torch.autograd on the fp64 twin
    (whose logits equal the reference classes' goldens): what autopgd_base.py:371-376 / foolbox value_and_grad ask of a source model."""
    import torch.nn.functional as F
    from robustart_b200 import nets, ops, torch_models as TM
    MEAN, STD = ops.IMAGENET_MEAN, ops.IMAGENET_STD
    # BatchNorm statistics of a calibration batch (tests/golden/calibrated_logits.npz), the synthetic head: with unrelated random
    # statistics these two networks collapse every image onto one feature and their gradients shrink ~1000x per block -- below any
    # float format's range long before the image (the fp64 twin returns 1e-40s); a network whose BN matches its data does not
    import os as _os
    import numpy as _np
    from util import calibrated_state_dict, HEAD_KEYS
    raw = nets.random_state_dict(nets._MOBILE_ARCHS[arch][1](), 0)
    cal = _np.load(_os.path.join(_os.path.dirname(__file__), "golden", "calibrated_logits.npz"))
    sd = calibrated_state_dict(arch, raw, cal)
    for k in (HEAD_KEYS[arch] + ".weight", HEAD_KEYS[arch] + ".bias"):
        sd[k] = raw[k]
    model = nets.build_model(arch, sd, device=cuda)
    twin = TM.build(arch, nets._strip_prefix(sd)).to(cuda).double().eval()
    torch.manual_seed(3)
    n = 2
    x01 = torch.rand(n, 3, 224, 224, device=cuda)
    y = torch.randint(0, 1000, (n,), device=cuda)
    logits, saved = model.forward_saved(x01)
    assert (logits - model.forward(x01)).abs().max().item() < 1e-3
    _, dlogits = ops.ce_loss_grad(logits, y)
    g = model.input_grad(dlogits, saved)
    assert torch.isfinite(g).all()
    m = torch.tensor(MEAN, device=cuda, dtype=torch.float64).view(1, 3, 1, 1)
    s = torch.tensor(STD, device=cuda, dtype=torch.float64).view(1, 3, 1, 1)
    xd = x01.double().requires_grad_(True)
    lt = twin((xd - m) / s)
    assert (lt.detach() - logits.double()).abs().max().item() < 1e-3
    (want,) = torch.autograd.grad(F.cross_entropy(lt, y, reduction="sum"), xd)
    scale = want.abs().max().item()
    assert scale > 1e-8, scale                      # the test network has a usable gradient
    cos = F.cosine_similarity(g.double().flatten(), want.flatten(), dim=0).item()
    err = (g.double() - want).abs().max().item()
    print(arch, "|grad| max %.2e" % scale, "grad max err / max %.2e  cos %.6f" % (err / scale, cos))
    assert cos > 0.999, cos
    assert err < 5e-2 * scale                       # ReLU6 / swish nets flip a few saturation masks between fp32 kernels and the fp64 twin
    big = want.abs() > 1e-2 * scale
    assert (torch.sign(g.double())[big] == torch.sign(want)[big]).float().mean().item() > 0.99


def test_mobile_sources_are_native(cuda):
    from robustart_b200 import attacks, solver
    for t in ("mobilenet_v2", "efficientnet_b0"):
        assert isinstance(solver.build_source_model({"type": t, "kwargs": {}}, None, cuda), attacks.NativeModel)


@pytest.mark.parametrize("cin,cout,act,with_res", [(16, 96, "relu6", False), (24, 144, "swish", False), (32, 16, None, False), (8, 8, "relu", True),
                                                   (32, 192, "relu6", False), (24, 24, None, True), (16, 1000, "sigmoid", False)])
def test_pointwise_smallk(cuda, cin, cout, act, with_res):
    """b200r_pointwise_smallk_nhwc (CUDA cores, exact fp32) against fp64 and against the tensor-core GEMM it replaces for narrow inputs
    (mobilenet_v2.py:52-60; efficientnet.py:312-321)."""
    from robustart_b200 import ops
    torch.manual_seed(cin * 7 + cout)
    x = torch.randn(3, 19, 23, cin, device=cuda)                  # 1311 pixels: not a multiple of the block
    w = torch.randn(cout, cin, device=cuda) / cin ** 0.5
    b = torch.randn(cout, device=cuda)
    res = torch.randn(3, 19, 23, cout, device=cuda) if with_res else None
    xp = ops.split_f32(x)
    rp = ops.split_f32(res) if with_res else None
    ref = ops.merge_f32(xp).double().view(-1, cin) @ w.double().t() + b.double()
    if with_res:
        ref = ref + ops.merge_f32(rp).double().view(-1, cout)
    fn = {"relu6": lambda v: v.clamp(0, 6), "swish": lambda v: v * torch.sigmoid(v), "relu": torch.relu, "sigmoid": torch.sigmoid, None: lambda v: v}[act]
    ref = fn(ref).view(3, 19, 23, cout)
    got = ops.merge_f32(ops.pointwise_smallk(xp, w, b, rp, act=act)).double()
    assert (got - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    gemm = ops.merge_f32(ops.conv2d_nhwc(xp, ops.split_f32(w.view(cout, 1, 1, cin).contiguous()), None, b, rp, act=act)).double()
    assert (got - gemm).abs().max().item() < 5e-5 * max(1.0, ref.abs().max().item())
