"""fp16 single-plane precision (passes = B200R_PASSES_F16): operands and stored activations are IEEE fp16, one MMA per
product, fp32 accumulation in TMEM.  Kernel-level check: the contraction of the fp16-ROUNDED operands must equal an
fp64 contraction of the same rounded operands up to fp32 accumulation error (the rounding of the inputs is the
format's, not the kernel's).  Model-level check: the north star's 1e-3 logit tolerance against the golden logits of the
reference's own classes."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import synth_images

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "resnet_logits.npz"))


def _h(x):
    return x.half().float()


def _rel_err(got, ref):
    return ((got.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def test_f16_roundtrip(cuda):
    from robustart_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(512, 256, device=cuda) * 3.0
    p = ops.to_planes(x, True)
    assert p.shape == (1, 512, 256) and p.dtype == torch.int16
    assert torch.equal(p[0].view(torch.float16), x.half())
    assert torch.equal(ops.from_planes(p), x.half().float())
    assert torch.equal(ops.from_planes(ops.to_planes(x, True, scale=4.0), scale=0.25), (x * 4).half().float() * 0.25)


@pytest.mark.parametrize("m,k,n", [(128, 64, 64), (1000, 192, 64), (4096, 512, 1000), (300, 2048, 1000), (32768, 128, 512),
                                   (20000, 320, 256)])
def test_linear_f16(cuda, m, k, n):
    from robustart_b200 import ops
    torch.manual_seed(m + k + n)
    x, w = _h(torch.randn(m, k, device=cuda)), _h(torch.randn(n, k, device=cuda) / k ** 0.5)
    b, s = torch.randn(n, device=cuda), torch.rand(n, device=cuda) + 0.5
    ref = (x.double() @ w.double().t()) * s.double() + b.double()
    out = torch.empty(m, n, device=cuda)
    ops.linear(ops.to_planes(x, True), ops.to_planes(w, True), s, b, out_f32=out, want_planes=False)
    assert _rel_err(out, ref) < 6e-6      # fp32 accumulation over K up to 2048
    # fp16 output plane + residual (LSU path, scaled) + relu: one rounding to fp16 on the way out
    r = _h(torch.randn(m, n, device=cuda))
    y = ops.linear(ops.to_planes(x, True), ops.to_planes(w, True), s, b, res=ops.to_planes(r, True), act="relu")
    assert y.shape == (1, m, n)
    ref2 = torch.relu(ref + r.double())
    got = ops.from_planes(y)
    assert ((got.double() - ref2).abs() <= ref2.abs() * 2 ** -11 + 1e-5).all()
    # unscaled + residual: the residual rides on the tensor core (fp16 identity k-blocks)
    y = ops.linear(ops.to_planes(x, True), ops.to_planes(w, True), None, b, res=ops.to_planes(r, True))
    ref3 = x.double() @ w.double().t() + b.double() + r.double()
    assert ((ops.from_planes(y).double() - ref3).abs() <= ref3.abs() * 2 ** -11 + 1e-5).all()


@pytest.mark.parametrize("shape", [(2, 56, 56, 64, 64, 3, 1, 1), (3, 28, 28, 128, 128, 3, 1, 1), (5, 7, 7, 512, 512, 3, 1, 1),
                                   (2, 56, 56, 64, 256, 1, 1, 0), (2, 56, 56, 128, 128, 3, 2, 1), (2, 56, 56, 256, 512, 1, 2, 0),
                                   (64, 14, 14, 256, 1024, 1, 1, 0)])
def test_conv_f16(cuda, shape):
    from robustart_b200 import ops
    n, h, w, cin, cout, k, stride, pad = shape
    torch.manual_seed(sum(shape))
    x = _h(torch.randn(n, cin, h, w, device=cuda))
    wt = _h(torch.randn(cout, cin, k, k, device=cuda) / (cin * k * k) ** 0.5)
    b = torch.randn(cout, device=cuda)
    ref = F.conv2d(x.double(), wt.double(), b.double(), stride, pad)
    ho, wo = ref.shape[2], ref.shape[3]
    r = _h(torch.randn(n, cout, ho, wo, device=cuda))
    ref = torch.relu(ref + r.double()).permute(0, 2, 3, 1)
    y = ops.conv2d_nhwc(ops.to_planes(x.permute(0, 2, 3, 1).contiguous(), True), ops.to_planes(wt.permute(0, 2, 3, 1).contiguous(), True),
                        None, b, ops.to_planes(r.permute(0, 2, 3, 1).contiguous(), True), stride=stride, pad=pad, act="relu")
    assert y.shape == (1, n, ho, wo, cout)
    got = ops.from_planes(y).double()
    assert ((got - ref).abs() <= ref.abs() * 2 ** -11 * 1.01 + 2e-5).all(), (got - ref).abs().max().item()


def test_pools_f16(cuda):
    from robustart_b200 import ops
    torch.manual_seed(3)
    x = _h(torch.randn(3, 64, 112, 112, device=cuda))
    p = ops.to_planes(x.permute(0, 2, 3, 1).contiguous(), True)
    got = ops.from_planes(ops.maxpool3x3s2(p)).permute(0, 3, 1, 2)
    assert torch.equal(got, F.max_pool2d(x, 3, 2, 1))
    x = _h(torch.randn(5, 2048, 7, 7, device=cuda))
    p = ops.to_planes(x.permute(0, 2, 3, 1).contiguous(), True)
    got = ops.from_planes(ops.global_avgpool(p))
    assert torch.equal(got, _h(x.mean((2, 3)))) or (got - x.mean((2, 3))).abs().max() < 1e-3


@pytest.mark.parametrize("n,h,w", [(2, 224, 224), (48, 224, 224), (3, 64, 64), (1, 40, 248), (5, 8, 8), (2, 100, 72)])
def test_stem_pool_one_launch(cuda, n, h, w):
    """conv1 + bn1 + relu + maxpool in one launch (csrc/stem_pool_sm100.cu: overlapping-descriptor implicit im2col) against
    an fp64 convolution of the same fp16-rounded operands (normalised pixels and weights), rounded once to fp16 and pooled
    -- resnet_official.py:221-227,330-334."""
    from robustart_b200 import ops
    torch.manual_seed(n * 7 + h + w)
    img = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device=cuda)
    wt = torch.randn(64, 3, 7, 7, device=cuda) / 147 ** 0.5
    s, b = torch.rand(64, device=cuda) + 0.5, torch.randn(64, device=cuda) * 0.3
    s[::5] *= -1                                  # BN scales can be negative: the scale must act before the ReLU
    wf = _h(wt * s.view(-1, 1, 1, 1))             # the kernel's contract: scale folded, then ONE rounding to fp16
    # the kernel's normalisation: fp16(fma_f32(byte, 1/(255 std), -mean/std)); fp64 emulates the fused multiply-add exactly
    k1 = torch.tensor([float(np.float32(1.0 / (255.0 * float(np.float32(sd))))) for sd in ops.IMAGENET_STD], device=cuda, dtype=torch.float64)
    k0 = torch.tensor([float(np.float32(-float(np.float32(m)) / float(np.float32(sd)))) for m, sd in zip(ops.IMAGENET_MEAN, ops.IMAGENET_STD)],
                      device=cuda, dtype=torch.float64)
    x = _h((img.permute(0, 3, 1, 2).double() * k1.view(1, 3, 1, 1) + k0.view(1, 3, 1, 1)).float())
    sel = list(range(n)) if n <= 8 else [0, 1, n // 3, n // 2, n - 2, n - 1]     # fp64 reference on a few images of the big batch
    conv = F.conv2d(x[sel].double(), wf.double(), b.double(), 2, 3)
    ref = F.max_pool2d(torch.relu(conv), 3, 2, 1).permute(0, 2, 3, 1)
    wp = ops.to_planes(ops.pack_stem_weight(wf).contiguous(), True)
    y = ops.stem_pool_u8(img, wp, b)
    assert y.shape == (1, n, h // 4, w // 4, 64)
    got = ops.from_planes(y).double()
    assert torch.isfinite(got).all()
    got = got[sel]
    # one rounding to fp16 on the way out; the bias pair and the fp32 accumulation order add ~1e-6 absolute
    assert ((got - ref).abs() <= ref.abs() * 2 ** -11 * 1.01 + 2e-5).all(), (got - ref).abs().max().item()
    # and the two-launch path it replaces (unfolded fp16 weights, scale in the epilogue, LUT normalisation) agrees to fp16 noise
    if w % 16 == 0:
        w16 = ops.to_planes(ops.pack_stem_weight(_h(wt)).contiguous(), True)
        two = ops.from_planes(ops.maxpool3x3s2(ops.stem_conv7x7_u8(img, w16, s, b, act="relu"))).double()[sel]
        assert (got - two).abs().max().item() < 2e-2 and (got - two).abs().mean().item() < 1e-3


@pytest.mark.parametrize("arch", ["resnet18", "resnet50"])
def test_logits_f16_within_tolerance(cuda, arch):
    from robustart_b200 import nets, ops
    model = nets.build_model(arch, device=cuda, seed=0, passes=ops.PASSES_F16)
    images = torch.from_numpy(synth_images(4, seed=7)).to(cuda)
    got = model(images).cpu().numpy()
    want = GOLD[arch]
    err = np.abs(got - want).max()
    assert err < 1e-3, (arch, err)          # north star tolerance; measured ~2e-4
    assert err < 5e-4, (arch, err)          # and we hold it with margin
    assert (got.argmax(1) == want.argmax(1)).all()
    for g, w in zip(got, want):
        assert set(np.argsort(-g)[:5]) == set(np.argsort(-w)[:5])
    x01 = images.permute(0, 3, 1, 2).float().div(255).contiguous()
    got2 = model(x01).cpu().numpy()
    assert np.abs(got2 - want).max() < 5e-4
    # graph replay == eager
    run = model.graphed(images)
    assert torch.equal(run(images), model(images))


@pytest.mark.parametrize("arch,n", [("resnet18", 3), ("resnet50", 2)])
def test_input_grad_f16_matches_autograd(cuda, arch, n):
    """fp16 input-gradient pass (loss-scaled) against autograd of the fp64 twin, in the L2 / cosine / sign sense
    (a random-init deep ReLU net flips a few masks between any two arithmetics, see test_backward_gpu.py)."""
    from robustart_b200 import nets, ops, torch_models
    MEAN, STD = ops.IMAGENET_MEAN, ops.IMAGENET_STD
    sd = nets.random_state_dict(nets.resnet_spec(arch), 0)
    net = nets.build_model(arch, sd, device=cuda, passes=ops.PASSES_F16)
    twin = torch_models.build(arch, sd).to(cuda).double().eval()
    torch.manual_seed(3)
    x = torch.rand(n, 3, 224, 224, device=cuda)
    y = torch.randint(0, 1000, (n,), device=cuda)
    loss, g, logits = net.loss_and_input_grad(x, y)
    assert torch.isfinite(g).all()
    xd = x.double().requires_grad_(True)
    mean = torch.tensor(MEAN, device=cuda, dtype=torch.float64).view(1, 3, 1, 1)
    std = torch.tensor(STD, device=cuda, dtype=torch.float64).view(1, 3, 1, 1)
    ref_logits = twin((xd - mean) / std)
    ref_loss = F.cross_entropy(ref_logits, y, reduction="sum")
    (ref,) = torch.autograd.grad(ref_loss, xd)
    assert (logits.double() - ref_logits).abs().max().item() < 1e-3
    gd, rd = g.double().flatten(), ref.flatten()
    rel_l2 = ((gd - rd).norm() / rd.norm()).item()
    cos = (torch.dot(gd, rd) / (gd.norm() * rd.norm())).item()
    assert rel_l2 < 8e-2 and cos > 0.997, (rel_l2, cos)
    big = ref.abs() > 1e-2 * ref.abs().max()
    agree = (torch.sign(g.double())[big] == torch.sign(ref)[big]).float().mean().item()
    assert agree > 0.985, agree


def test_pgd_f16_source(cuda):
    from robustart_b200 import attacks, nets, ops
    sd = nets.random_state_dict(nets.resnet_spec("resnet18"), 1)
    net = nets.build_model("resnet18", sd, device=cuda, passes=ops.PASSES_F16)
    ref = nets.build_model("resnet18", sd, device=cuda)
    torch.manual_seed(4)
    x = torch.rand(4, 3, 224, 224, device=cuda)
    y = torch.randint(0, 1000, (4,), device=cuda)
    eps = 4 / 255
    u = torch.rand_like(x)
    a = attacks.pgd_linf(x, y, attacks.NativeModel(net), eps, 3 / 40, 3, start_uniform=u)
    b = attacks.pgd_linf(x, y, attacks.NativeModel(ref), eps, 3 / 40, 3, start_uniform=u)
    assert (a - x).abs().max().item() <= eps + 1e-6 and a.min().item() >= 0 and a.max().item() <= 1
    assert ((a - b).abs() > 1e-6).float().mean().item() < 0.08
    la = F.cross_entropy(ref.forward(a), y, reduction="sum").item()
    l0 = F.cross_entropy(ref.forward(x), y, reduction="sum").item()
    assert la > l0
