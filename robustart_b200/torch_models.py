"""Autograd-capable PyTorch twins of the kernel models -- used ONLY as the *source* model of an attack
(the input gradient needs a backward pass; dgrad kernels are the next step).  Same
state_dict keys as the reference (resnet_official.py) so one checkpoint feeds both paths."""
import torch
import torch.nn as nn
import torch.nn.functional as F

_CFG = {"resnet18": (False, [2, 2, 2, 2]), "resnet34": (False, [3, 4, 6, 3]), "resnet50": (True, [3, 4, 6, 3]),
        "resnet101": (True, [3, 4, 23, 3])}


class _Block(nn.Module):
    def __init__(self, cin, planes, stride, bottleneck):
        super().__init__()
        cout = planes * (4 if bottleneck else 1)
        if bottleneck:
            shapes = [(cin, planes, 1, 1, 0), (planes, planes, 3, stride, 1), (planes, cout, 1, 1, 0)]
        else:
            shapes = [(cin, planes, 3, stride, 1), (planes, planes, 3, 1, 1)]
        for i, (a, b, k, s, p) in enumerate(shapes, 1):
            setattr(self, "conv%d" % i, nn.Conv2d(a, b, k, s, p, bias=False))
            setattr(self, "bn%d" % i, nn.BatchNorm2d(b))
        self.n = len(shapes)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        idn = x if self.downsample is None else self.downsample(x)
        out = x
        for i in range(1, self.n + 1):
            out = getattr(self, "bn%d" % i)(getattr(self, "conv%d" % i)(out))
            if i < self.n:
                out = F.relu(out)
        return F.relu(out + idn)


class ResNet(nn.Module):
    def __init__(self, arch, num_classes=1000):
        super().__init__()
        bott, layers = _CFG[arch]
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        for li, (planes, nb) in enumerate(zip([64, 128, 256, 512], layers), 1):
            blocks = []
            for b in range(nb):
                blocks.append(_Block(cin, planes, 2 if (b == 0 and li > 1) else 1, bott))
                cin = planes * (4 if bott else 1)
            setattr(self, "layer%d" % li, nn.Sequential(*blocks))
        self.fc = nn.Linear(cin, num_classes)

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.bn1(self.conv1(x))), 3, 2, 1)
        for li in range(1, 5):
            x = getattr(self, "layer%d" % li)(x)
        return self.fc(torch.flatten(F.adaptive_avg_pool2d(x, 1), 1))


def build(arch, state_dict=None):
    m = ResNet(arch)
    if state_dict is not None:
        m.load_state_dict(state_dict, strict=True)
    return m.eval()


# ------------------------------------------------------------------------------------------------
# token models: ViT-B/16 (vision_transformer.py:44-349) and MLP-Mixer-B/16 (vit/mlp_mixer.py:7-159, vit/vit_base.py).
# Functional twins over the reference's own state_dict keys (nets.vit_spec / nets.mixer_spec): autograd source models for
# the attacks on BASELINE configs[2] / [4]; the TARGET model of an evaluation always runs on the B200 kernels.
# ------------------------------------------------------------------------------------------------
def _gelu_tanh(x):          # vision_transformer.py:19-37
    return 0.5 * x * (1.0 + torch.tanh(0.7978845608028654 * (x + 0.044715 * x ** 3)))


class _ParamModel(nn.Module):
    def __init__(self, state_dict):
        super().__init__()
        self.keys = {}
        for i, (k, v) in enumerate(state_dict.items()):
            name = "p%d" % i
            self.register_buffer(name, v.detach().clone().float())
            self.keys[k] = name

    def w(self, key):
        return getattr(self, self.keys[key])


class ViT(_ParamModel):
    def __init__(self, state_dict, depth=12, dim=768, heads=12, patch=16):
        super().__init__(state_dict)
        self.depth, self.dim, self.heads, self.patch = depth, dim, heads, patch

    def forward(self, x):                                   # normalised float32 NCHW
        n, w = x.shape[0], self.w
        x = F.conv2d(x, w("embedding.weight"), w("embedding.bias"), stride=self.patch).flatten(2).transpose(1, 2)
        x = torch.cat([w("cls_token").expand(n, -1, -1), x], 1) + w("pos_embedding")
        hd = self.dim // self.heads
        for d in range(self.depth):
            p = "transformer.encoders.encoder_%d." % d
            y = F.layer_norm(x, (self.dim,), w(p + "norm1.weight"), w(p + "norm1.bias"), 1e-5)
            qkv = F.linear(y, w(p + "attention.to_qkv.weight"), w(p + "attention.to_qkv.bias"))
            q, k, v = qkv.view(n, -1, 3, self.heads, hd).permute(2, 0, 3, 1, 4)          # "(qkv h d)" packing, :82
            a = torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5, -1) @ v
            a = a.permute(0, 2, 1, 3).reshape(n, -1, self.dim)
            x = x + F.linear(a, w(p + "attention.to_out.weight"), w(p + "attention.to_out.bias"))
            y = F.layer_norm(x, (self.dim,), w(p + "norm2.weight"), w(p + "norm2.bias"), 1e-5)
            y = _gelu_tanh(F.linear(y, w(p + "feedforward.mlp1.weight"), w(p + "feedforward.mlp1.bias")))
            x = x + F.linear(y, w(p + "feedforward.mlp2.weight"), w(p + "feedforward.mlp2.bias"))
        x = F.layer_norm(x, (self.dim,), w("transformer.encoder_norm.weight"), w("transformer.encoder_norm.bias"), 1e-5)[:, 0]
        if "pre_logits.weight" in self.keys:
            x = torch.tanh(F.linear(x, w("pre_logits.weight"), w("pre_logits.bias")))       # :289-291
        return F.linear(x, w("head.weight"), w("head.bias"))


class Mixer(_ParamModel):
    def __init__(self, state_dict, depth=12, dim=768, patch=16):
        super().__init__(state_dict)
        self.depth, self.dim, self.patch = depth, dim, patch

    def forward(self, x):
        w = self.w
        x = F.conv2d(x, w("patch_embed.proj.weight"), w("patch_embed.proj.bias"), stride=self.patch).flatten(2).transpose(1, 2)
        for d in range(self.depth):
            p = "blocks.%d." % d
            y = F.layer_norm(x, (self.dim,), w(p + "norm1.weight"), w(p + "norm1.bias"), 1e-6).transpose(1, 2)
            y = F.linear(F.gelu(F.linear(y, w(p + "token_mix.fc1.weight"), w(p + "token_mix.fc1.bias"))),
                         w(p + "token_mix.fc2.weight"), w(p + "token_mix.fc2.bias"))
            x = x + y.transpose(1, 2)
            y = F.layer_norm(x, (self.dim,), w(p + "norm2.weight"), w(p + "norm2.bias"), 1e-6)
            x = x + F.linear(F.gelu(F.linear(y, w(p + "channel_mix.fc1.weight"), w(p + "channel_mix.fc1.bias"))),
                             w(p + "channel_mix.fc2.weight"), w(p + "channel_mix.fc2.bias"))
        x = F.layer_norm(x, (self.dim,), w("norm.weight"), w("norm.bias"), 1e-6).mean(1)
        return F.linear(x, w("head.weight"), w("head.bias"))


# ------------------------------------------------------------------------------------------------
# mobile families: MobileNetV2 x1.0 (mobilenet_v2.py:31-202) and EfficientNet-B0 (efficientnet.py:289-495), functional twins
# over the reference's state_dict keys (nets.mobilenet_v2_spec / nets.efficientnet_b0_spec), eval mode (BN affine).
# ------------------------------------------------------------------------------------------------
_MBV2_SETTING = [[1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1], [6, 160, 3, 2], [6, 320, 1, 1]]
_EFFB0_BLOCKS = [(1, 3, 1, 1, 32, 16), (2, 3, 2, 6, 16, 24), (2, 5, 2, 6, 24, 40), (3, 3, 2, 6, 40, 80), (3, 5, 1, 6, 80, 112),
                 (4, 5, 2, 6, 112, 192), (1, 3, 1, 6, 192, 320)]


class _ConvNetTwin(_ParamModel):
    def bn(self, x, name):
        w = self.w
        return F.batch_norm(x, w(name + ".running_mean"), w(name + ".running_var"), w(name + ".weight"), w(name + ".bias"), False, 0.0, 1e-5)

    def cbn(self, x, conv, bn, stride=1, groups=1):
        wt = self.w(conv + ".weight")
        return self.bn(F.conv2d(x, wt, None, stride, wt.shape[-1] // 2, 1, groups), bn)


class MobileNetV2(_ConvNetTwin):
    def forward(self, x):
        x = F.relu6(self.cbn(x, "features.0.0", "features.0.1", 2))
        cin, idx = 32, 1
        for t, c, n, s in _MBV2_SETTING:
            for i in range(n):
                stride, p, j = (s if i == 0 else 1), "features.%d.conv." % idx, 0
                y = x
                if t != 1:
                    y = F.relu6(self.cbn(y, p + "0.0", p + "0.1"))
                    j = 1
                y = F.relu6(self.cbn(y, p + "%d.0" % j, p + "%d.1" % j, stride, groups=y.shape[1]))
                y = self.cbn(y, p + "%d" % (j + 1), p + "%d" % (j + 2))
                x = x + y if (stride == 1 and cin == c) else y
                cin, idx = c, idx + 1
        x = F.relu6(self.cbn(x, "features.%d.0" % idx, "features.%d.1" % idx))
        return F.linear(x.mean((2, 3)), self.w("classifier.1.weight"), self.w("classifier.1.bias"))


class EfficientNetB0(_ConvNetTwin):
    def forward(self, x):
        sw = lambda v: v * torch.sigmoid(v)                      # efficientnet.py:271-277
        w = self.w
        x = sw(self.cbn(x, "stem.0", "stem.1", 2))
        bi = 0
        for rep, k, s, e, cin, cout in _EFFB0_BLOCKS:
            for r in range(rep):
                ci, stride = (cin if r == 0 else cout), (s if r == 0 else 1)
                p, j = "blocks.%d." % bi, 0
                y = x
                if e != 1:
                    y = sw(self.cbn(y, p + "in_conv.0", p + "in_conv.1"))
                    j = 3
                y = sw(self.cbn(y, p + "in_conv.%d" % j, p + "in_conv.%d" % (j + 1), stride, groups=y.shape[1]))
                q = y.mean((2, 3), keepdim=True)                                  # squeeze on the expanded tensor
                q = sw(F.conv2d(q, w(p + "se_block.conv1.weight"), w(p + "se_block.conv1.bias")))
                q = torch.sigmoid(F.conv2d(q, w(p + "se_block.conv2.weight"), w(p + "se_block.conv2.bias")))
                y = self.cbn(y * q, p + "out_conv.0", p + "out_conv.1")
                x = x + y if (stride == 1 and ci == cout) else y
                bi += 1
        x = sw(self.cbn(x, "head.0", "head.1"))
        return F.linear(x.mean((2, 3)), w("fc.weight"), w("fc.bias"))


_TOKEN = {"vit_b16_224": ViT, "vit_base_patch16_224": ViT, "mixer_b16_224": Mixer,
          "mobilenet_v2": MobileNetV2, "mobilenet_v2_x1_0": MobileNetV2, "efficientnet_b0": EfficientNetB0}
_build_resnet_twin = build


def build(arch, state_dict=None):  # noqa: F811
    if arch in _TOKEN:
        assert state_dict is not None, "the token twins are functional over a state_dict"
        return _TOKEN[arch](state_dict).eval()
    return _build_resnet_twin(arch, state_dict)
