"""ViT-B/16 and MLP-Mixer-B/16 on the B200 kernels: layer checks vs torch and golden logits from the
REFERENCE's own classes (tests/golden/make_golden_models.py --tokens).  Tolerance 1e-3 on logits."""
import os

import numpy as np
import pytest
import torch

from util import synth_images

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "token_logits.npz"))


def test_layernorm_and_attention_layers(cuda):
    from robustart_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(3 * 197, 768, device=cuda) * 2 + 0.3
    g, b = torch.rand(768, device=cuda) + 0.5, torch.randn(768, device=cuda)
    for eps in (1e-5, 1e-6):
        y = ops.merge_f32(ops.layernorm(ops.split_f32(x), g, b, eps=eps))
        xs = ops.merge_f32(ops.split_f32(x))
        ref = torch.nn.functional.layer_norm(xs.double(), (768,), g.double(), b.double(), eps)
        assert (y.double() - ref).abs().max().item() < 2e-4
    qkv = torch.randn(3 * 197, 3 * 768, device=cuda)
    out = ops.merge_f32(ops.attention(ops.split_f32(qkv), 3, 197, 12, 64, 64 ** -0.5))
    q, k, v = ops.merge_f32(ops.split_f32(qkv)).double().view(3, 197, 3, 12, 64).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 64 ** -0.5, -1) @ v).permute(0, 2, 1, 3).reshape(3 * 197, 768)
    assert (out.double() - ref).abs().max().item() < 2e-4


@pytest.mark.parametrize("n,t,heads", [(3, 197, 12), (1, 256, 2), (5, 129, 3), (2, 50, 4), (4, 128, 1), (1, 1, 2)])
def test_attention_tensor_core(cuda, n, t, heads, monkeypatch):
    """softmax(q k^T scale) v on tcgen05 (csrc/attention_sm100.cu: QK^T and PV as split-bf16 MMAs, V consumed MN-major,
    P written into the operand layout by the softmax warps) against fp64 torch on the same split-rounded inputs, and
    against the CUDA-core kernel it replaces -- vision_transformer.py:80-92."""
    from robustart_b200 import ops
    torch.manual_seed(n * 100 + t)
    qkv = torch.randn(n * t, 3 * heads * 64, device=cuda) * 1.5
    qkv[:, : heads * 64] *= 2.0                                  # sharper softmax rows
    planes = ops.split_f32(qkv)
    out = ops.merge_f32(ops.attention(planes, n, t, heads, 64, 64 ** -0.5))
    q, k, v = ops.merge_f32(planes).double().view(n, t, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 64 ** -0.5, -1) @ v).permute(0, 2, 1, 3).reshape(n * t, heads * 64)
    assert torch.isfinite(out).all()
    assert (out.double() - ref).abs().max().item() < 2e-4
    assert torch.equal(planes, ops.split_f32(qkv))              # inputs untouched


def test_token_transposes_and_patches(cuda):
    from robustart_b200 import ops
    torch.manual_seed(1)
    x = torch.randn(2, 196, 768, device=cuda)
    xp = ops.split_f32(x.view(2 * 196, 768))
    y = ops.merge_f32(ops.tokens_to_channels(xp, 2, 196, 768, 256)).view(2, 768, 256)
    xm = ops.merge_f32(xp).view(2, 196, 768)
    assert torch.equal(y[:, :, :196], xm.transpose(1, 2)) and y[:, :, 196:].abs().max().item() == 0
    back = ops.merge_f32(ops.channels_to_tokens_add(ops.split_f32(y.reshape(2 * 768, 256)), xp, 2, 196, 768, 256)).view(2, 196, 768)
    assert (back - 2 * xm).abs().max().item() < 1e-5
    img = torch.randint(0, 256, (2, 224, 224, 3), dtype=torch.uint8, device=cuda)
    cols = ops.merge_f32(ops.patch_gather(img, 16)).view(2, 196, 768)
    ref = torch.nn.functional.unfold(ops.u8nhwc_to_f32nchw(img), 16, stride=16).transpose(1, 2)   # (c, ky, kx) columns
    assert (cols - ref).abs().max().item() < 1e-4


@pytest.mark.parametrize("geo", [(3, 196, 768, 256), (2, 10, 64, 16), (1, 196, 72, 256), (2, 70, 132, 72), (2, 10, 66, 16), (1, 66, 64, 70),
                                 (2, 257, 196, 260)])
def test_token_transposes_all_geometries(cuda, geo):
    """Both transposes of the Mixer's token mixing, exact: the four-element kernels (t_pad % 4 == 0 and c % 4 == 0) over ragged tiles in
    both directions, the two-element kernels for the other geometries; the padding columns of the channel-major tensor come out zero
    even when the buffer held something else, and the way back adds an independent residual."""
    from robustart_b200 import ops
    b, t, c, tp = geo
    torch.manual_seed(sum(geo))
    x = torch.randn(b * t, c, device=cuda)
    res = torch.randn(b * t, c, device=cuda)
    xp, rp = ops.split_f32(x), ops.split_f32(res)
    yp = ops.tokens_to_channels(xp, b, t, c, tp)
    y = ops.merge_f32(yp).view(b, c, tp)
    xm = ops.merge_f32(xp).view(b, t, c)
    assert torch.equal(y[:, :, :t], xm.transpose(1, 2)) and y[:, :, t:].abs().max().item() == 0
    back = ops.merge_f32(ops.channels_to_tokens_add(yp, rp, b, t, c, tp)).view(b, t, c)
    want = xm.double() + ops.merge_f32(rp).view(b, t, c).double()
    assert (back.double() - want).abs().max().item() < 1e-6 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("arch", ["vit_b16_224", "mixer_b16_224"])
def test_token_model_logits_match_reference(cuda, arch):
    from robustart_b200 import nets
    model = nets.build_model(arch, device=cuda, seed=0)
    images = torch.from_numpy(synth_images(2, seed=9)).to(cuda)
    got = model(images).cpu().numpy()
    want = GOLD[arch]
    err = np.abs(got - want).max()
    assert err < 1e-3, (arch, err)
    assert (got.argmax(1) == want.argmax(1)).all()
    for g, w in zip(got, want):
        assert set(np.argsort(-g)[:5]) == set(np.argsort(-w)[:5])
    x01 = images.permute(0, 3, 1, 2).float().div(255).contiguous()
    assert np.abs(model(x01).cpu().numpy() - got).max() < 1e-4
