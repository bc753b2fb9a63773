"""Native input-gradient pass (backward_layers.cu + the GEMM on transposed weights) vs torch.autograd in fp64/fp32.

The attack loops only ever need d loss / d image (autopgd_base.py:371-376; foolbox value_and_grad): each layer's
backward kernel is checked on its own, then the whole ResNet input gradient against the autograd twin built from
the same state dict."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def _planes(t):
    """float NCHW -> split planes NHWC"""
    from robustart_b200 import ops
    return ops.split_f32(t.permute(0, 2, 3, 1).contiguous())


def _nchw(p):
    from robustart_b200 import ops
    return ops.merge_f32(p).permute(0, 3, 1, 2)


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def test_relu_bwd(cuda):
    from robustart_b200 import ops
    torch.manual_seed(0)
    act = torch.relu(torch.randn(3, 16, 9, 7, device=cuda))
    act[0, 0, 0, :3] = 0.0
    dy, add = torch.randn_like(act), torch.randn_like(act)
    pa, pd, pb = _planes(act), _planes(dy), _planes(add)
    got = ops.relu_bwd(pd, pa)
    # pure masking passes the planes through bit-exactly
    m = (act > 0).permute(0, 2, 3, 1)
    assert torch.equal(got[0], torch.where(m, pd[0], torch.zeros_like(pd[0])))
    assert torch.equal(got[1], torch.where(m, pd[1], torch.zeros_like(pd[1])))
    got2 = _nchw(ops.relu_bwd(pd, pa, pb))
    ref2 = torch.where(act > 0, ops.merge_f32(pd).permute(0, 3, 1, 2), torch.zeros_like(act)) + ops.merge_f32(pb).permute(0, 3, 1, 2)
    assert (got2 - ref2).abs().max().item() < 1e-4     # the sum is re-split to 16 significant bits


def test_dilate2(cuda):
    from robustart_b200 import ops
    x = torch.randn(2, 8, 5, 6, device=cuda)
    y = _nchw(ops.dilate2(_planes(x)))
    ref = torch.zeros(2, 8, 10, 12, device=cuda)
    ref[:, :, ::2, ::2] = _nchw(_planes(x))
    assert torch.equal(y, ref)


@pytest.mark.parametrize("h,w", [(112, 112), (14, 10), (7, 9)])
def test_maxpool_bwd(cuda, h, w):
    from robustart_b200 import ops
    torch.manual_seed(1)
    x = torch.relu(torch.randn(3, 16, h, w, device=cuda))      # ReLU output: plenty of exact ties at 0
    xq = _nchw(_planes(x)).double().requires_grad_(True)        # the values the kernel sees
    dy = torch.randn_like(F.max_pool2d(xq, 3, 2, 1)).detach()
    got = _nchw(ops.maxpool3x3s2_bwd(_planes(x), _planes(dy.float())))
    dyq = _nchw(_planes(dy.float())).double()
    (ref,) = torch.autograd.grad(F.max_pool2d(xq, 3, 2, 1), xq, dyq)
    assert (got.double() - ref).abs().max().item() < 1e-4     # up to four dy summed, re-split to 16 significant bits


def test_avgpool_bwd(cuda):
    from robustart_b200 import ops
    dy = torch.randn(4, 64, device=cuda)
    got = ops.merge_f32(ops.global_avgpool_bwd(ops.split_f32(dy), 7, 7))     # [n, 7, 7, c]
    assert (got - (dy / 49).view(4, 1, 1, 64)).abs().max().item() < 1e-6


@pytest.mark.parametrize("k,stride,cin,cout,hw", [(1, 1, 64, 256, 14), (3, 1, 64, 64, 14), (3, 2, 128, 128, 28), (1, 2, 256, 512, 28),
                                                  (3, 2, 64, 128, 10), (3, 1, 256, 256, 14), (3, 1, 128, 128, 28), (3, 1, 512, 512, 7),
                                                  (1, 1, 256, 1024, 14), (1, 1, 512, 2048, 7), (3, 1, 64, 64, 56)])
def test_conv_dgrad(cuda, k, stride, cin, cout, hw):
    """_ConvBN.dgrad == autograd of conv2d (folded BN scale included) for every conv geometry of the ResNets."""
    from robustart_b200 import nets
    torch.manual_seed(2)
    sd = {"c.weight": torch.randn(cout, cin, k, k) * 0.05, "b.weight": torch.rand(cout) + 0.5, "b.bias": torch.randn(cout),
          "b.running_mean": torch.randn(cout) * 0.1, "b.running_var": torch.rand(cout) + 0.5}
    conv = nets._ConvBN(sd, "c", "b", cuda, stride, k // 2)
    x = torch.randn(2, cin, hw, hw, device=cuda, dtype=torch.float64, requires_grad=True)
    scale = (sd["b.weight"].double() / torch.sqrt(sd["b.running_var"].double() + nets.BN_EPS)).to(cuda)
    y = F.conv2d(x, sd["c.weight"].double().to(cuda) * scale.view(-1, 1, 1, 1), stride=stride, padding=k // 2)
    dy = torch.randn_like(y)
    (ref,) = torch.autograd.grad(y, x, dy)
    res = torch.randn(2, cin, hw, hw, device=cuda)
    if k == 1 and stride == 2:
        got = _nchw(conv.dgrad(_planes(dy.float())))
        assert _rel(got, ref) < 3e-5
    else:
        got = _nchw(conv.dgrad(_planes(dy.float()), res=_planes(res)))
        assert _rel(got, ref + res.double()) < 3e-5
        # fused ReLU backward: the result is zeroed wherever the activation it flows into is not positive
        act = torch.relu(torch.randn(2, cin, hw, hw, device=cuda))
        got = _nchw(conv.dgrad(_planes(dy.float()), res=_planes(res), mask=_planes(act)))
        want = (ref + res.double()) * (act > 0)
        assert _rel(got, want) < 3e-5
        assert (got[act.expand_as(got) <= 0] == 0).all()


@pytest.mark.parametrize("arch,n", [("resnet18", 3), ("resnet50", 2)])
def test_resnet_input_grad_matches_autograd(cuda, arch, n):
    from robustart_b200 import nets, ops, torch_models
    sd = nets.random_state_dict(nets.resnet_spec(arch), 0)
    net = nets.build_model(arch, sd, device=cuda)
    twin = torch_models.build(arch, sd).to(cuda).double().eval()
    torch.manual_seed(3)
    x = torch.rand(n, 3, 224, 224, device=cuda)
    y = torch.randint(0, 1000, (n,), device=cuda)
    loss, g, logits = net.loss_and_input_grad(x, y)
    xd = x.double().requires_grad_(True)
    mean = torch.tensor(MEAN, device=cuda, dtype=torch.float64).view(1, 3, 1, 1)
    std = torch.tensor(STD, device=cuda, dtype=torch.float64).view(1, 3, 1, 1)
    ref_logits = twin((xd - mean) / std)
    ref_loss = F.cross_entropy(ref_logits, y, reduction="sum")
    (ref,) = torch.autograd.grad(ref_loss, xd)
    assert (logits.double() - ref_logits).abs().max().item() < 1e-3
    assert abs(loss.double().sum().item() - ref_loss.item()) < 1e-3 * max(1.0, abs(ref_loss.item()))
    # A random-initialised deep ReLU net is chaotic in the max norm: ONE activation within 1e-5 of zero whose mask differs
    # between the fp64 twin and the kernels moves the gradient of its whole receptive cone by ~1 % of max|g| (measured:
    # resnet18 1.1 %, resnet50 5 %, from 1-3 flipped masks out of ~1e7).  So the end-to-end check is in the L2 / cosine
    # sense; test_block_backward_strict pins every block to 3e-5 given identical masks.
    gd, rd = g.double().flatten(), ref.flatten()
    rel_l2 = ((gd - rd).norm() / rd.norm()).item()
    cos = (torch.dot(gd, rd) / (gd.norm() * rd.norm())).item()
    assert rel_l2 < 5e-2 and cos > 0.999, (rel_l2, cos)
    big = ref.abs() > 1e-2 * ref.abs().max()
    agree = (torch.sign(g.double())[big] == torch.sign(ref)[big]).float().mean().item()
    assert agree > 0.99, agree


@pytest.mark.parametrize("arch", ["resnet18", "resnet50"])
def test_block_backward_strict(cuda, arch):
    """Every residual block's backward (relu masks, dgrad GEMMs on flipped weights, stride-2 dilation, downsample branch,
    residual join) against an fp64 torch restatement that uses the SAME saved activations -> no mask chaos, tight bound."""
    from robustart_b200 import nets, ops
    sd = nets.random_state_dict(nets.resnet_spec(arch), 0)
    net = nets.build_model(arch, sd, device=cuda)
    torch.manual_seed(5)
    x = torch.rand(2, 3, 224, 224, device=cuda)
    _, saved = net.forward_saved(x)
    nchw = lambda p: ops.merge_f32(p).permute(0, 3, 1, 2).double()
    wt = lambda c: c._w_folded.double().to(cuda)
    for blk, sv in zip(net.blocks, saved["blocks"]):
        y = sv[-1]
        g = _planes(torch.randn(y.shape[1], y.shape[4], y.shape[2], y.shape[3], device=cuda))
        got = nchw(nets.ResNet.block_backward(blk, sv, g))
        dz = nchw(g) * (nchw(y) > 0)
        s2 = blk["c2"].stride if blk["kind"] == "bottleneck" else blk["c1"].stride
        opad = s2 - 1
        if "down" in blk:
            r = F.conv_transpose2d(dz, wt(blk["down"]), stride=blk["down"].stride, output_padding=blk["down"].stride - 1)
        else:
            r = dz
        if blk["kind"] == "bottleneck":
            a1, a2, _ = sv
            t = F.conv_transpose2d(dz, wt(blk["c3"])) * (nchw(a2) > 0)
            t = F.conv_transpose2d(t, wt(blk["c2"]), stride=s2, padding=1, output_padding=opad) * (nchw(a1) > 0)
            ref = F.conv_transpose2d(t, wt(blk["c1"])) + r
        else:
            a1, _ = sv
            t = F.conv_transpose2d(dz, wt(blk["c2"]), padding=1) * (nchw(a1) > 0)
            ref = F.conv_transpose2d(t, wt(blk["c1"]), stride=s2, padding=1, output_padding=opad) + r
        assert _rel(got, ref) < 3e-5


def test_stem_backward_strict(cuda):
    """maxpool bwd -> relu mask -> stem dgrad GEMM -> col2im (+ 1/std) against autograd of the fp64 stem on the same masks."""
    from robustart_b200 import nets, ops
    sd = nets.random_state_dict(nets.resnet_spec("resnet18"), 0)
    net = nets.build_model("resnet18", sd, device=cuda)
    torch.manual_seed(6)
    x = torch.rand(2, 3, 224, 224, device=cuda)
    logits, saved = net.forward_saved(x)
    g = _planes(torch.randn(2, 64, 56, 56, device=cuda))
    net.input_grad(torch.zeros_like(logits), saved)                       # builds the transposed stem weights
    # the saved forward keeps the pool's arg-max codes (stem ReLU folded in), not the 112 x 112 activation
    assert saved["stem"] is None and saved["pool_codes"].shape == (2, 56, 56, 64)
    gm = ops.maxpool3x3s2_bwd_codes_hi(saved["pool_codes"], g, 112, 112)
    s0 = net._stem_f32(x, net.passes)
    two = ops.from_planes(ops.relu_bwd(ops.maxpool3x3s2_bwd(s0, g), s0))
    assert (ops.from_planes(gm) - two).abs().max().item() <= 2 ** -10 * two.abs().max().item()
    # the stem GEMM's gradient runs on single fp16 planes (hi plane of the masked gradient x fp16 weights): its [n*112*112, 192] output
    # is the largest tensor of the pass and only feeds col2im
    got = ops.stem_col2im(ops.linear(gm[:1].view(1, -1, 64), net._stem_wt, passes=ops.PASSES_F16), 2, 224, 224)
    # reference: d/dx of sum(maxpool(relu(conv(norm(x)) * s + b)) * g) in fp64
    xd = x.double().requires_grad_(True)
    mean = torch.tensor(MEAN, device=cuda, dtype=torch.float64).view(1, 3, 1, 1)
    std = torch.tensor(STD, device=cuda, dtype=torch.float64).view(1, 3, 1, 1)
    s = (sd["bn1.weight"].double() / torch.sqrt(sd["bn1.running_var"].double() + nets.BN_EPS)).to(cuda)
    b = (sd["bn1.bias"].double().to(cuda) - sd["bn1.running_mean"].double().to(cuda) * s)
    pre = F.conv2d((xd - mean) / std, sd["conv1.weight"].double().to(cuda), stride=2, padding=3) * s.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
    out = F.max_pool2d(F.relu(pre), 3, 2, 1)
    (ref,) = torch.autograd.grad(out, xd, _nchw(g).double())
    gd, rd = got.double().flatten(), ref.flatten()
    assert ((gd - rd).norm() / rd.norm()).item() < 2e-3          # 11-bit operands on this last contraction + a handful of relu / argmax ties
    assert torch.quantile((gd - rd).abs()[::97], 0.999).item() < 1e-3 * rd.abs().max().item()
    assert F.cosine_similarity(gd, rd, dim=0).item() > 0.99999


def test_pgd_native_vs_autograd_source(cuda):
    """pgd_linf driven by the native gradient lands (almost everywhere) on the same adversarial as the autograd twin."""
    from robustart_b200 import attacks, nets, torch_models
    sd = nets.random_state_dict(nets.resnet_spec("resnet18"), 1)
    net = nets.build_model("resnet18", sd, device=cuda)
    twin = torch_models.build("resnet18", sd).to(cuda).eval()
    torch.manual_seed(4)
    x = torch.rand(4, 3, 224, 224, device=cuda)
    y = torch.randint(0, 1000, (4,), device=cuda)
    eps = 4 / 255
    u = torch.rand_like(x)
    a = attacks.pgd_linf(x, y, attacks.NativeModel(net), eps, 3 / 40, 3, start_uniform=u)
    b = attacks.pgd_linf(x, y, attacks.PyTorchModel(twin, preprocessing=dict(mean=MEAN, std=STD, axis=-3)), eps, 3 / 40, 3, start_uniform=u)
    assert (a - x).abs().max().item() <= eps + 1e-6 and a.min().item() >= 0 and a.max().item() <= 1
    differ = ((a - b).abs() > 1e-6).float().mean().item()
    assert differ < 0.05, differ    # sign flips of near-zero gradient entries (fp32 twin vs split-bf16 kernels)
    with torch.no_grad():
        la = F.cross_entropy(net.forward(a), y, reduction="sum").item()
        l0 = F.cross_entropy(net.forward(x), y, reduction="sum").item()
    assert la > l0


@pytest.mark.parametrize("planes", [2, 1])
def test_maxpool_relu_bwd_fused_matches_two_passes(cuda, planes):
    """maxpool backward with the ReLU backward fused in and one fp16 plane out = relu_bwd(maxpool_bwd(x, dy), x), hi plane
    (resnet_official.py:225-227)."""
    from robustart_b200 import ops
    torch.manual_seed(planes)
    f16 = planes == 1
    x = torch.relu(torch.randn(3, 20, 28, 64, device=cuda))
    x[0, :6] = 0.0                                           # whole windows at zero: relu'(0) = 0 routes nothing
    dy = torch.randn(3, 10, 14, 64, device=cuda)
    xp, dyp = ops.to_planes(x, f16), ops.to_planes(dy, f16)
    want = ops.relu_bwd(ops.maxpool3x3s2_bwd(xp, dyp), xp)
    got = ops.maxpool3x3s2_relu_bwd_hi(xp, dyp)
    assert got.shape == (1, 3, 20, 28, 64)
    a, b = ops.from_planes(got), ops.from_planes(want)
    assert (a - b).abs().max().item() <= 2 ** -10 * b.abs().max().item()       # the hi plane of a split pair is the fp16 rounding
    assert ((a == 0) == (b == 0)).all()
    # and against autograd
    xr = ops.from_planes(xp).requires_grad_(True)
    (torch.nn.functional.max_pool2d(torch.relu(xr.permute(0, 3, 1, 2)), 3, 2, 1) * ops.from_planes(dyp).permute(0, 3, 1, 2)).sum().backward()
    ref = xr.grad
    assert (a - ref).abs().max().item() <= 2 ** -10 * ref.abs().max().item() + 1e-6


@pytest.mark.parametrize("n,ho,wo,cin,cout,with_res", [(3, 14, 14, 128, 128, False), (2, 7, 9, 64, 256, True), (5, 28, 28, 128, 128, False)])
def test_dgrad_3x3_stride2_by_parity_classes(cuda, n, ho, wo, cin, cout, with_res):
    """b200r_conv2d_dgrad3x3s2_nhwc (four parity-class convolutions of dy written through strided tensor maps) against autograd of
    the stride-2 convolution and against the dilate + 3x3 path it replaces (resnet_official.py:112)."""
    from robustart_b200 import ops
    torch.manual_seed(n + ho + cin)
    h, w = 2 * ho, 2 * wo
    wt = torch.randn(cout, cin, 3, 3, device=cuda) / (cin * 9) ** 0.5
    x = torch.randn(n, cin, h, w, device=cuda, dtype=torch.float64, requires_grad=True)
    dy = torch.randn(n, ho, wo, cout, device=cuda)
    act = torch.relu(torch.randn(n, h, w, cin, device=cuda))                  # the activation the gradient flows into
    res = torch.randn(n, h, w, cin, device=cuda) if with_res else None
    y = torch.nn.functional.conv2d(x, wt.double(), None, 2, 1)
    (ref,) = torch.autograd.grad(y, x, dy.permute(0, 3, 1, 2).double())
    ref = ref.permute(0, 2, 3, 1)
    if with_res:
        ref = ref + res.double()
    ref = ref * (act > 0)
    wflip = wt.flip(2, 3).permute(1, 2, 3, 0)                                 # [cin, ky', kx', cout]
    S = ([1], [0, 2])
    wsub = [ops.to_planes(wflip[:, S[a]][:, :, S[b]].contiguous()) for a in (0, 1) for b in (0, 1)]
    dyp, ap = ops.to_planes(dy), ops.to_planes(act)
    rp = ops.to_planes(res) if with_res else None
    got = ops.from_planes(ops.conv2d_dgrad3x3s2(dyp, wsub, rp, ap)).double()
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() < 2e-5 * scale
    old = ops.from_planes(ops.conv2d_dgrad(ops.dilate2(dyp), ops.to_planes(wflip.contiguous()), rp, ap, pad=1)).double()
    assert (got - old).abs().max().item() < 2e-5 * scale


def test_maxpool_codes_forward_and_backward(cuda):
    """b200r_maxpool3x3s2_nhwc_codes + b200r_maxpool3x3s2_bwd_codes_hi (codes written by the forward pool of the attack path's saved
    forward) = the plain pool and the fused two-pass backward, bit for bit."""
    from robustart_b200 import ops
    torch.manual_seed(3)
    x = torch.relu(torch.randn(3, 20, 28, 64, device=cuda))
    x[1, 4:9] = 0.0
    dy = torch.randn(3, 10, 14, 64, device=cuda)
    xp, dyp = ops.to_planes(x), ops.to_planes(dy)
    y, codes = ops.maxpool3x3s2_codes(xp)
    assert torch.equal(y, ops.maxpool3x3s2(xp))
    assert codes.dtype == torch.uint8 and ((codes <= 8) | (codes == 15)).all()
    got = ops.maxpool3x3s2_bwd_codes_hi(codes, dyp, 20, 28)
    assert torch.equal(got, ops.maxpool3x3s2_relu_bwd_hi(xp, dyp))


@pytest.mark.parametrize("n,h,w", [(2, 32, 32), (1, 64, 48), (3, 224, 224)])
def test_stem_col2im_against_explicit_scatter(cuda, n, h, w):
    """b200r_stem_col2im_f32_f16 (transpose of the stem's im2col composed with 1/std): dcols [n*ho*wo, 192] fp16, column = ky*24 + kx*3 + c,
    against an explicit scatter-add in fp64 (resnet_official.py:221-224 reversed)."""
    from robustart_b200 import ops
    torch.manual_seed(h + w)
    ho, wo = h // 2, w // 2
    d = torch.randn(n * ho * wo, 192, device=cuda).half()
    d.view(n, ho, wo, 8, 24)[..., 21:] = 0                      # the im2col's padding columns carry nothing
    d[:, 168:] = 0
    got = ops.stem_col2im(d.view(torch.int16).unsqueeze(0).contiguous(), n, h, w).double()
    dd = d.double().view(n, ho, wo, 192)
    ref = torch.zeros(n, 3, h + 6, w + 6, device=cuda, dtype=torch.float64)     # padded by 3: iy + 3 = 2 oy + ky
    for ky in range(7):
        for kx in range(7):
            ref[:, :, ky:ky + 2 * ho:2, kx:kx + 2 * wo:2] += dd[..., ky * 24 + kx * 3: ky * 24 + kx * 3 + 3].permute(0, 3, 1, 2)
    std = torch.tensor(ops.IMAGENET_STD, device=cuda, dtype=torch.float64).view(1, 3, 1, 1)
    ref = ref[:, :, 3:3 + h, 3:3 + w] / std
    assert (got - ref).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("n,h,w,cin,cout,masked,f16", [
    (5, 28, 28, 256, 512, True, False), (3, 14, 14, 512, 1024, True, False), (2, 56, 56, 64, 128, False, False),
    (2, 15, 13, 64, 72, True, False),                     # odd spatial size: ho = (h + 1) // 2, last row / column has no partner
    (3, 14, 14, 256, 512, True, True)])
def test_downsample_dgrad_accumulates_in_place(cuda, n, h, w, cin, cout, masked, f16):
    """b200r_conv2d_dgrad1x1s2_acc_nhwc: dx[:, 2i, 2j] = mask * (dx[:, 2i, 2j] + W^T dy[:, i, j]) with output and residual the SAME
    strided view of dx -- against fp64 and against the path it replaces (dgrad on the small map, dilate2, residual operand of the main
    branch's dgrad); positions off the stride-2 lattice must keep their bits."""
    from robustart_b200 import ops
    torch.manual_seed(n + h + cin)
    ho, wo = (h + 1) // 2, (w + 1) // 2
    wt = torch.randn(cout, cin, device=cuda) / cin ** 0.5                     # the 1x1 downsample kernel
    dy = torch.randn(n, ho, wo, cout, device=cuda)
    act = torch.relu(torch.randn(n, h, w, cin, device=cuda))
    main = torch.randn(n, h, w, cin, device=cuda) * (act > 0 if masked else 1)   # the main branch's (already masked) gradient
    dyp, mp, ap = ops.to_planes(dy, f16), ops.to_planes(main, f16), ops.to_planes(act, f16)
    wtp = ops.to_planes(wt.t().contiguous().view(cin, 1, 1, cout), f16)       # [cdx, 1, 1, cdy]
    before = mp.clone()
    passes = ops.PASSES_F16 if f16 else 3
    out = ops.conv2d_dgrad1x1s2_acc(dyp, wtp, mp, ap if masked else None, passes=passes)
    assert out.data_ptr() == mp.data_ptr()
    got = ops.from_planes(mp).double()
    ref = ops.from_planes(before).double()
    add = ops.from_planes(dyp).double() @ ops.from_planes(wtp).double().view(cin, cout).t()
    ref[:, ::2, ::2] += add
    if masked:
        ref = ref * (act > 0)
    tol = (2e-3 if f16 else 2e-5) * ref.abs().max().item()
    assert (got - ref).abs().max().item() < tol
    lattice = torch.zeros(h, w, dtype=torch.bool, device=cuda)
    lattice[::2, ::2] = True
    assert torch.equal(mp[:, :, ~lattice], before[:, :, ~lattice])            # untouched bits elsewhere
    if masked and h % 2 == 0 and w % 2 == 0:
        # the replaced path: contract on the small map, zero-insert, mask, add the (already masked) main branch
        old = ops.from_planes(ops.relu_bwd(ops.dilate2(ops.conv2d_dgrad(dyp, wtp, passes=passes)), ap, add=before))
        assert (got - old.double()).abs().max().item() < tol
