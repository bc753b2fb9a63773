"""tcgen05 implicit-GEMM kernel vs torch (float64 reference).  Tolerance: split-bf16 with 3 passes keeps
~16 mantissa bits per operand => relative error of a length-K dot product ~ 2^-16; we assert
|err| <= 3e-5 * sqrt(K) * rms(a)*rms(b) style bounds via a normalised max error."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel_err(got, ref):
    return ((got.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def test_split_merge_roundtrip(cuda):
    from robustart_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(1024, 256, device=cuda) * 37.0
    p = ops.split_f32(x)
    y = ops.merge_f32(p)
    assert ((y - x).abs() <= x.abs() * 2 ** -21 + 2 ** -24).all()      # fp16 hi + fp16 lo: 22 bits, absolute floor = fp16 subnormal spacing
    hi = p[0].view(torch.float16).float()
    assert torch.equal(hi, x.half().float())


@pytest.mark.parametrize("m,k,n", [(128, 64, 64), (256, 128, 128), (1000, 192, 64), (4096, 512, 1000), (300, 2048, 1000),
                                   (32768, 128, 512), (20000, 320, 256)])   # the last two take the BN = 256 tiles
@pytest.mark.parametrize("passes", [3, 1])
def test_linear_matches_fp64(cuda, m, k, n, passes):
    from robustart_b200 import ops
    torch.manual_seed(m + k + n)
    x = torch.randn(m, k, device=cuda)
    w = torch.randn(n, k, device=cuda) / k ** 0.5
    b = torch.randn(n, device=cuda)
    s = torch.rand(n, device=cuda) + 0.5
    ref = (x.double() @ w.double().t()) * s.double() + b.double()
    out = torch.empty(m, n, device=cuda)
    ops.linear(ops.split_f32(x), ops.split_f32(w), s, b, passes=passes, out_f32=out, want_planes=False)
    torch.cuda.synchronize()
    err = _rel_err(out, ref)
    assert err < (2e-5 if passes == 3 else 2e-2), err
    # planes output + residual + relu
    r = torch.randn(m, n, device=cuda)
    y = ops.linear(ops.split_f32(x), ops.split_f32(w), s, b, res=ops.split_f32(r), act="relu", passes=passes)
    got = ops.merge_f32(y)
    ref2 = torch.relu(ref + r.double())
    assert _rel_err(got, ref2) < (4e-5 if passes == 3 else 2e-2)


CONVS = [
    # n, h, w, cin, cout, k, stride, pad
    (2, 56, 56, 64, 64, 3, 1, 1),
    (3, 28, 28, 128, 128, 3, 1, 1),
    (5, 14, 14, 256, 256, 3, 1, 1),
    (5, 7, 7, 512, 512, 3, 1, 1),
    (2, 56, 56, 64, 256, 1, 1, 0),
    (2, 56, 56, 128, 128, 3, 2, 1),
    (3, 28, 28, 256, 256, 3, 2, 1),
    (2, 56, 56, 256, 512, 1, 2, 0),
    (3, 14, 14, 1024, 2048, 1, 2, 0),
    (2, 14, 14, 512, 512, 3, 2, 1),
    (2, 16, 16, 64, 64, 5, 1, 2),
]


@pytest.mark.parametrize("cfg", CONVS)
def test_conv_matches_torch(cuda, cfg):
    from robustart_b200 import ops
    n, h, w, cin, cout, k, stride, pad = cfg
    torch.manual_seed(sum(cfg))
    x = torch.randn(n, cin, h, w, device=cuda)
    wt = torch.randn(cout, cin, k, k, device=cuda) / (cin * k * k) ** 0.5
    s = torch.rand(cout, device=cuda) + 0.5
    b = torch.randn(cout, device=cuda)
    ref = torch.nn.functional.conv2d(x.double(), wt.double(), stride=stride, padding=pad)
    ref = ref * s.double().view(1, -1, 1, 1) + b.double().view(1, -1, 1, 1)
    ho, wo = ref.shape[2:]
    res = torch.randn(n, cout, ho, wo, device=cuda)
    ref = torch.relu(ref + res.double())
    xp = ops.split_f32(x.permute(0, 2, 3, 1).contiguous())
    wp = ops.split_f32(wt.permute(0, 2, 3, 1).contiguous())
    rp = ops.split_f32(res.permute(0, 2, 3, 1).contiguous())
    y = ops.conv2d_nhwc(xp, wp, s, b, rp, stride=stride, pad=pad, act="relu")
    torch.cuda.synchronize()
    got = ops.merge_f32(y).permute(0, 3, 1, 2)
    err = _rel_err(got, ref)
    assert err < 4e-5, (cfg, err)


def test_pool_layers(cuda):
    from robustart_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(3, 64, 112, 112, device=cuda)
    xp = ops.split_f32(x.permute(0, 2, 3, 1).contiguous())
    xm = ops.merge_f32(xp).permute(0, 3, 1, 2)   # what the planes actually hold
    mp = ops.merge_f32(ops.maxpool3x3s2(xp)).permute(0, 3, 1, 2)
    assert torch.equal(mp, torch.nn.functional.max_pool2d(xm, 3, 2, 1))
    ap = ops.merge_f32(ops.global_avgpool(xp))
    assert (ap - xm.mean(dim=(2, 3))).abs().max().item() < 1e-5


def test_stem_im2col(cuda):
    from robustart_b200 import ops
    torch.manual_seed(0)
    img = torch.randint(0, 256, (2, 224, 224, 3), dtype=torch.uint8, device=cuda)
    p = ops.stem_im2col(img)
    cols = ops.merge_f32(p).view(2, 112, 112, 192)
    xn = ops.u8nhwc_to_f32nchw(img)
    ref = torch.nn.functional.unfold(xn, 7, padding=3, stride=2)         # [n, 3*49, L] ordered (c, ky, kx)
    ref = ref.view(2, 3, 7, 7, 112, 112).permute(0, 4, 5, 2, 3, 1)         # [n, oy, ox, ky, kx, c]
    cols = cols[..., :168].reshape(2, 112, 112, 7, 8, 3)                    # K = (ky, 8 kx slots, c), then zeros
    assert (cols[..., :7, :] - ref).abs().max().item() < 1e-4
    assert cols[..., 7, :].abs().max().item() == 0
    assert ops.merge_f32(p).view(2, 112, 112, 192)[..., 168:].abs().max().item() == 0
    # float path
    x01 = torch.rand(2, 3, 224, 224, device=cuda)
    p2 = ops.merge_f32(ops.stem_im2col(x01)).view(2, 112, 112, 192)
    ref2 = torch.nn.functional.unfold(ops.normalize(x01), 7, padding=3, stride=2)
    ref2 = ref2.view(2, 3, 7, 7, 112, 112).permute(0, 4, 5, 2, 3, 1)
    assert (p2[..., :168].reshape(2, 112, 112, 7, 8, 3)[..., :7, :] - ref2).abs().max().item() < 1e-4


def test_fused_stem_matches_im2col_path(cuda):
    """The fused uint8 stem (LUT normalisation, producer-warp gather) equals conv2d on the normalised image."""
    from robustart_b200 import ops
    torch.manual_seed(0)
    img = torch.randint(0, 256, (5, 224, 224, 3), dtype=torch.uint8, device=cuda)
    w = torch.randn(64, 3, 7, 7, device=cuda) * 0.1
    s = torch.rand(64, device=cuda) + 0.5
    b = torch.randn(64, device=cuda)
    wp = ops.pack_stem_weight(w)
    y = ops.stem_conv7x7_u8(img, ops.split_f32(wp), s, b, act="relu")
    torch.cuda.synchronize()
    got = ops.merge_f32(y).permute(0, 3, 1, 2)
    xn = ops.u8nhwc_to_f32nchw(img)
    ref = torch.nn.functional.conv2d(xn.double(), w.double(), stride=2, padding=3)
    ref = torch.relu(ref * s.double().view(1, -1, 1, 1) + b.double().view(1, -1, 1, 1))
    assert _rel_err(got, ref) < 4e-5
    # batch that is not a multiple of the 128-row tile and odd geometry
    img2 = torch.randint(0, 256, (1, 32, 48, 3), dtype=torch.uint8, device=cuda)
    y2 = ops.merge_f32(ops.stem_conv7x7_u8(img2, ops.split_f32(wp), s, b, act="relu")).permute(0, 3, 1, 2)
    ref2 = torch.relu(torch.nn.functional.conv2d(ops.u8nhwc_to_f32nchw(img2).double(), w.double(), stride=2, padding=3)
                      * s.double().view(1, -1, 1, 1) + b.double().view(1, -1, 1, 1))
    assert _rel_err(y2, ref2) < 4e-5


def test_fused_stem_f32_matches_conv(cuda):
    """The float32-NCHW fused stem (attack path) == conv2d on the normalised image == the im2col + GEMM route."""
    from robustart_b200 import ops
    torch.manual_seed(1)
    w = torch.randn(64, 3, 7, 7, device=cuda) * 0.1
    s = torch.rand(64, device=cuda) + 0.5
    b = torch.randn(64, device=cuda)
    wp = ops.split_f32(ops.pack_stem_weight(w))
    for shape in [(5, 3, 224, 224), (2, 3, 32, 48), (130, 3, 64, 64)]:       # several images per CTA -> image-boundary restarts
        x = torch.rand(*shape, device=cuda)
        got = ops.merge_f32(ops.stem_conv7x7_f32(x, wp, s, b, act="relu")).permute(0, 3, 1, 2)
        ref = F.conv2d(ops.normalize(x).double(), w.double(), stride=2, padding=3)
        ref = torch.relu(ref * s.double().view(1, -1, 1, 1) + b.double().view(1, -1, 1, 1))
        assert _rel_err(got, ref) < 4e-5
        cols = ops.stem_im2col(x)
        via = ops.merge_f32(ops.linear(cols, wp, s, b, act="relu")).view(shape[0], shape[2] // 2, shape[3] // 2, 64).permute(0, 3, 1, 2)
        assert (got - via).abs().max().item() < 1e-5


PAIR_CONVS = [
    # n, h, w, cin, cout, k, stride, pad, residual
    (5, 14, 14, 256, 256, 3, 1, 1, False),     # 980 rows: 8 M tiles
    (3, 14, 14, 256, 1024, 1, 1, 0, True),     # 588 rows: 5 M tiles -- the odd tail tile's partner is out of range
    (1, 7, 7, 512, 2048, 1, 1, 0, True),       # one M tile only: no pair launch
    (3, 28, 28, 128, 128, 3, 1, 1, False),
    (2, 56, 56, 64, 64, 3, 1, 1, False),       # BN = 64 pairs (32 weight rows per CTA)
    (2, 56, 56, 64, 64, 3, 1, 1, True),
    (3, 28, 28, 256, 256, 3, 2, 1, False),
    (2, 14, 14, 1024, 2048, 1, 2, 0, False),
    (7, 8, 8, 192, 200, 1, 1, 0, True),        # ragged K (192 = 3 k-blocks) and ragged Cout (200: second N tile mostly out of range)
]


@pytest.mark.parametrize("cfg", PAIR_CONVS)
@pytest.mark.parametrize("passes", [3, 1])
def test_cta_pair_gemm_is_bit_identical(cuda, cfg, passes, monkeypatch):
    """The CTA-pair kernel (tcgen05.mma.cta_group::2: one M = 256 MMA over two M tiles, half the weight tile per CTA) adds the same
    products in the same order as the one-CTA kernel: outputs must be bit-identical, and right against fp64."""
    from robustart_b200 import ops
    n, h, w, cin, cout, k, stride, pad, with_res = cfg
    torch.manual_seed(sum(cfg[:8]) + passes)
    x = torch.randn(n, h, w, cin, device=cuda)
    wt = torch.randn(cout, k, k, cin, device=cuda) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device=cuda)
    xp, wp = ops.split_f32(x), ops.split_f32(wt)
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    res = torch.randn(n, ho, wo, cout, device=cuda) if with_res else None
    rp = ops.split_f32(res) if with_res else None
    outs = {}
    for opts in ("32", "16"):                   # pairs wherever supported / never
        monkeypatch.setenv("B200R_GEMM_OPTS", opts)
        outs[opts] = ops.conv2d_nhwc(xp, wp, None, b, rp, stride=stride, pad=pad, act="relu", passes=passes).clone()
        torch.cuda.synchronize()
    assert torch.equal(outs["32"], outs["16"])
    if passes == 3:
        ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), wt.permute(0, 3, 1, 2).double(), b.double(), stride, pad).permute(0, 2, 3, 1)
        if with_res:
            ref = ref + res.double()
        assert _rel_err(ops.merge_f32(outs["32"]), torch.relu(ref)) < 4e-5


KEEP_PRE = [
    # m, k, nout, act
    (1576, 768, 3072, "gelu_tanh"),      # ViT MLP (8 images): pairs
    (300, 256, 384, "gelu_erf"),         # Mixer token mixing: ragged rows, 3 N tiles
    (130, 64, 64, "swish"),              # BN = 64
    (128, 192, 200, "tanh"),             # one M tile (no pair), ragged K and Cout
    (517, 96, 40, "sigmoid"),
]


@pytest.mark.parametrize("cfg", KEEP_PRE)
@pytest.mark.parametrize("opts", ["0", "16", "32"])
def test_linear_keep_pre_is_linear_then_activation(cuda, cfg, opts, monkeypatch):
    """b200r_linear_keep_pre (one launch, two outputs) against the two launches it replaces -- b200r_linear(act none) and
    b200r_act_planes: the pre-activation planes bit-identical, the activation within the two implementations' rounding (the pass
    kernel and the GEMM epilogue evaluate the same formulas) and right against fp64."""
    from robustart_b200 import ops
    m, k, nout, act = cfg
    monkeypatch.setenv("B200R_GEMM_OPTS", opts)
    torch.manual_seed(m + k + nout)
    x = torch.randn(m, k, device=cuda)
    wt = torch.randn(nout, k, device=cuda) / k ** 0.5 * 2
    b = torch.randn(nout, device=cuda)
    xp, wp = ops.split_f32(x).view(2, m, k), ops.split_f32(wt).view(2, nout, k)
    y, pre = ops.linear_keep_pre(xp, wp, b, act=act)
    pre_ref = ops.linear(xp, wp, None, b, act=None, passes=3)
    fused_ref = ops.linear(xp, wp, None, b, act=act, passes=3)
    torch.cuda.synchronize()
    assert torch.equal(pre, pre_ref)
    assert torch.equal(y, fused_ref)                       # same epilogue code as the fused-activation GEMM
    via_pass = ops.merge_f32(ops.act_planes(pre_ref, act))
    assert (ops.merge_f32(y) - via_pass).abs().max().item() < 2e-6
    z = x.double() @ wt.double().t() + b.double()
    want = {"gelu_tanh": lambda t: torch.nn.functional.gelu(t, approximate="tanh"), "gelu_erf": torch.nn.functional.gelu,
            "swish": torch.nn.functional.silu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}[act](z)
    assert (ops.merge_f32(y).double() - want).abs().max().item() < 3e-5
    with pytest.raises(ValueError):
        ops.linear_keep_pre(xp, wp, b, act="relu")         # ReLU's derivative needs no pre-activation: refused
