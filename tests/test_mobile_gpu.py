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
