"""The input-gradient pass of the token models (nets.ViT / nets.Mixer .forward_saved + .input_grad: which tensors are saved,
the order of the transposed layers, where the residual gradients join) against torch.autograd on the functional twins whose
logits equal the reference's classes (robustart_b200/torch_models.py, tests/golden/token_logits.npz), on CPU.

The device kernels are replaced -- in this test only -- by float64 torch statements of the FORMULAS the kernels implement
(token_backward.cu: LayerNorm backward, GELU/tanh derivatives, the two-phase attention backward, the patch scatter), so the
test also checks that mathematics; "planes" are a [2, ...] float64 tensor with the value in plane 0.  What it cannot check is
the CUDA itself: that is tests/test_tokens_gpu.py's job once the pass is switched on (B200R_NATIVE_TOKEN_GRAD=1)."""
import math

import pytest
import torch
import torch.nn.functional as F

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
DT = torch.float64


def _pl(v):
    return torch.stack([v, torch.zeros_like(v)])


def _act(v, act):
    if act == "gelu_tanh":
        return 0.5 * v * (1 + torch.tanh(0.7978845608028654 * (v + 0.044715 * v ** 3)))
    if act == "gelu_erf":
        return 0.5 * v * (1 + torch.erf(v * 0.7071067811865476))
    if act == "tanh":
        return torch.tanh(v)
    assert act is None
    return v


def _act_deriv(v, act):                                       # token_backward.cu act_deriv
    if act == "gelu_tanh":
        k = 0.7978845608028654
        t = torch.tanh(k * (v + 0.044715 * v ** 3))
        return 0.5 * (1 + t) + 0.5 * v * (1 - t * t) * k * (1 + 3 * 0.044715 * v * v)
    if act == "gelu_erf":
        return 0.5 * (1 + torch.erf(v * 0.7071067811865476)) + v * 0.3989422804014327 * torch.exp(-0.5 * v * v)
    t = torch.tanh(v)
    return 1 - t * t


def _install(monkeypatch):
    from robustart_b200 import nets
    ops = nets.ops
    launched = []

    def split_f32(x):
        return _pl(x.to(DT))

    def merge_f32(p):
        return p[0]

    def linear(x, wgt, scale=None, bias=None, res=None, *, act=None, passes=3, out=None, out_f32=None, want_planes=True):
        k = x.shape[-1]
        y = x[0].reshape(-1, k) @ wgt[0].t()
        if bias is not None:
            y = y + bias.to(DT)
        y = _act(y, act)
        if res is not None:
            y = y + res[0].reshape(y.shape)
        if out_f32 is not None:
            out_f32.copy_(y)
            return out_f32
        return _pl(y)

    def linear_keep_pre(x, wgt, bias=None, *, act):            # gemm_sm100.cu, DUAL epilogue: (act(pre), pre) from one launch
        pre = linear(x, wgt, None, bias)
        return _pl(_act(pre[0], act)), pre

    def layernorm(x, gamma, beta, eps=1e-5, out=None):
        return _pl(F.layer_norm(x[0], (x.shape[-1],), gamma.to(DT), beta.to(DT), eps))

    def layernorm_bwd(dy, x, gamma, eps=1e-5, add=None):      # token_backward.cu layernorm_bwd_kernel
        launched.append("layernorm_bwd")
        v, c = x[0], x.shape[-1]
        mean = v.mean(-1, keepdim=True)
        rstd = torch.rsqrt(((v - mean) ** 2).mean(-1, keepdim=True) + eps)
        xhat = (v - mean) * rstd
        gy = dy[0].reshape(v.shape) * gamma.to(DT)
        dx = rstd * (gy - gy.mean(-1, keepdim=True) - xhat * (gy * xhat).mean(-1, keepdim=True))
        return _pl(dx if add is None else dx + add[0].reshape(dx.shape))

    def patch_gather(img, patch=16, mean=MEAN, std=STD):
        n, _, h, w = img.shape
        m, s = torch.tensor(mean, dtype=DT).view(1, 3, 1, 1), torch.tensor(std, dtype=DT).view(1, 3, 1, 1)
        cols = F.unfold((img.to(DT) - m) / s, patch, stride=patch)          # [n, 3*p*p, np], column = c*p*p + ky*p + kx
        return _pl(cols.transpose(1, 2).reshape(-1, 3 * patch * patch))

    def patch_scatter(dcols, n, h, w, patch=16, std=STD, unscale=1.0):     # token_backward.cu patch_scatter_kernel
        launched.append("patch_scatter")
        gh, gw = h // patch, w // patch
        d = dcols[0].reshape(n, gh, gw, 3, patch, patch).permute(0, 3, 1, 4, 2, 5).reshape(n, 3, h, w)
        return (d * unscale / torch.tensor(std, dtype=DT).view(1, 3, 1, 1)).to(torch.float32)

    def assemble_tokens(x, cls, pos, n, num_patches):
        c = x.shape[-1]
        t = torch.cat([cls.to(DT).view(1, 1, c).expand(n, 1, c), x[0].view(n, num_patches, c)], 1) + pos.to(DT).view(1, -1, c)
        return _pl(t.reshape(-1, c))

    def _qkv(qkv, n, tokens, heads, hd):
        return qkv[0].view(n, tokens, 3, heads, hd).permute(2, 0, 3, 1, 4)          # "(qkv h d)"

    def attention(qkv, n, tokens, heads, head_dim, scale):
        q, k, v = _qkv(qkv, n, tokens, heads, head_dim)
        a = torch.softmax(q @ k.transpose(-1, -2) * scale, -1) @ v
        return _pl(a.permute(0, 2, 1, 3).reshape(n * tokens, heads * head_dim))

    def attention_bwd(qkv, dout, n, tokens, heads, head_dim, scale):                # token_backward.cu attention_bwd_kernel
        launched.append("attention_bwd")
        q, k, v = _qkv(qkv, n, tokens, heads, head_dim)
        do = dout[0].view(n, tokens, heads, head_dim).permute(0, 2, 1, 3)
        s = q @ k.transpose(-1, -2) * scale
        mx = s.max(-1, keepdim=True).values
        e = torch.exp(s - mx)
        inv = 1 / e.sum(-1, keepdim=True)
        dp = do @ v.transpose(-1, -2)
        delta = (e * dp).sum(-1, keepdim=True) * inv
        ds = e * inv * (dp - delta)                                                 # phase A
        dq = ds @ k * scale
        p = torch.exp(s - mx) * inv                                                 # phase B rebuilds p from (max, 1/sum)
        dv = p.transpose(-1, -2) @ do
        dk = (p * (dp - delta)).transpose(-1, -2) @ q * scale
        d = torch.stack([dq, dk, dv]).permute(1, 3, 0, 2, 4).reshape(n * tokens, 3 * heads * head_dim)
        return _pl(d)

    def tokens_to_channels(x, b, t, c, t_pad):
        y = torch.zeros(b, c, t_pad, dtype=DT)
        y[:, :, :t] = x[0].view(b, t, c).transpose(1, 2)
        return _pl(y.reshape(b * c, t_pad))

    def channels_to_tokens_add(y, res, b, t, c, t_pad):
        return _pl((res[0].view(b, t, c) + y[0].view(b, c, t_pad)[:, :, :t].transpose(1, 2)).reshape(b * t, c))

    def act_planes(pre, act):
        return _pl(_act(pre[0], act))

    def act_bwd_planes(dy, pre, act):
        launched.append("act_bwd")
        return _pl(dy[0].reshape(pre[0].shape) * _act_deriv(pre[0], act))

    def global_avgpool(x, out=None):
        P, n, h, w, c = x.shape
        return _pl(x[0].reshape(n, h * w, c).mean(1))

    def global_avgpool_bwd(dy, h, w):
        P, n, c = dy.shape
        return _pl((dy[0] / (h * w)).view(n, 1, 1, c).expand(n, h, w, c).contiguous())

    for name, fn in list(locals().items()):
        if callable(fn) and hasattr(ops, name):
            monkeypatch.setattr(ops, name, fn)
    return nets, launched


def _grad_twin(twin, x01, dlogits):
    m, s = torch.tensor(MEAN, dtype=DT).view(1, 3, 1, 1), torch.tensor(STD, dtype=DT).view(1, 3, 1, 1)
    x = x01.to(DT).requires_grad_(True)
    logits = twin.double()((x - m) / s)
    (g,) = torch.autograd.grad(logits, x, grad_outputs=dlogits.to(DT))
    return logits.detach(), g


@pytest.mark.parametrize("family", ["mixer", "vit", "vit_no_prelogits"])
def test_token_input_grad_matches_autograd(monkeypatch, family):
    nets, launched = _install(monkeypatch)
    from robustart_b200 import torch_models as TM
    torch.manual_seed(0)
    n, img, patch, classes = 3, 32, 8, 10
    if family == "mixer":
        depth, dim = 2, 64
        sd = nets.random_token_state_dict(nets.mixer_spec(depth=depth, dim=dim, patch=patch, img=img, classes=classes), 1)
        model = nets.Mixer(sd, "cpu", passes=3, depth=depth, dim=dim, patch=patch)
        twin = TM.Mixer(sd, depth=depth, dim=dim, patch=patch)
    else:
        depth, dim, heads = 2, 64, 4
        rep = dim if family == "vit" else None
        sd = nets.random_token_state_dict(nets.vit_spec(depth=depth, dim=dim, mlp=4 * dim, patch=patch, img=img, classes=classes, rep=rep), 1)
        model = nets.ViT(sd, "cpu", passes=3, depth=depth, dim=dim, heads=heads, patch=patch)
        twin = TM.ViT(sd, depth=depth, dim=dim, heads=heads, patch=patch)
    x01 = torch.rand(n, 3, img, img)
    dlogits = torch.randn(n, classes)
    logits, saved = model.forward_saved(x01)
    want_logits, want_g = _grad_twin(twin, x01, dlogits)
    assert (logits.to(DT) - want_logits).abs().max().item() < 1e-5             # logits buffer is float32
    assert (model.forward(x01).to(DT) - want_logits).abs().max().item() < 1e-5  # forward() and forward_saved() are the same network
    g = model.input_grad(dlogits, saved)
    assert g.shape == x01.shape and g.dtype == torch.float32
    scale = want_g.abs().max().item()
    assert scale > 0
    assert (g.to(DT) - want_g).abs().max().item() < 1e-6 * max(scale, 1.0) + 2e-7 * scale
    # every transposed layer ran: 2 LayerNorm backwards and one (Mixer: two) activation backward per block + the final norm
    assert launched.count("layernorm_bwd") == 2 * depth + 1
    assert launched.count("act_bwd") == (2 * depth if family == "mixer" else depth + (1 if family == "vit" else 0))
    assert launched.count("attention_bwd") == (0 if family == "mixer" else depth)
    assert launched.count("patch_scatter") == 1


def test_native_model_accepts_token_nets(monkeypatch):
    """attacks.NativeModel only needs forward_saved / input_grad: with those in place a token model can be an attack source."""
    nets, _ = _install(monkeypatch)
    from robustart_b200 import attacks as A
    sd = nets.random_token_state_dict(nets.mixer_spec(depth=1, dim=64, patch=8, img=32, classes=10), 1)
    src = A.NativeModel(nets.Mixer(sd, "cpu", passes=3, depth=1, dim=64, patch=8))
    x = torch.rand(2, 3, 32, 32)
    logits, vjp = src.forward_vjp(x)
    d = torch.zeros(2, 10)
    d[:, 3] = 1.0
    g = vjp(d)
    assert logits.shape == (2, 10) and g.shape == x.shape and torch.isfinite(g).all() and g.abs().max() > 0
    assert math.isfinite(float(logits.sum()))
