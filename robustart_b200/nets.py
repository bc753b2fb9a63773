"""Classifier forward passes on the B200 kernels (eval mode).

ResNet-18/50 in the reference's layout (prototype/prototype/model/resnet_official.py:40-140,221-239,
330-346 == torchvision: stride on the 3x3 of a Bottleneck, bias-free convs, BatchNorm after every conv).
A model is built from a plain state_dict with the reference's keys; BatchNorm (eval) is folded into a
per-channel scale/bias applied in the GEMM epilogue, weights are re-laid out once to [Cout, KH, KW, Cin]
split-bf16 planes.  Activations stay NHWC split-bf16 planes end to end; every layer is one launch of a
kernel in libb200robust.so; the whole forward can be captured into one CUDA graph.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import ops

BN_EPS = 1e-5  # nn.BatchNorm2d default, misc.py:115-143 get_bn


# ------------------------------------------------------------------------------------------------
# architecture description (keys and shapes of the reference's state_dict)
# ------------------------------------------------------------------------------------------------
_RESNET_CFG = {"resnet18": ("basic", [2, 2, 2, 2]), "resnet34": ("basic", [3, 4, 6, 3]),
               "resnet50": ("bottleneck", [3, 4, 6, 3]), "resnet101": ("bottleneck", [3, 4, 23, 3])}
ARCH_ALIASES = {"resnet18_official": "resnet18", "resnet34_official": "resnet34", "resnet50_official": "resnet50",
                "resnet101_official": "resnet101"}


def _bn_spec(prefix, c):
    return [(prefix + ".weight", (c,)), (prefix + ".bias", (c,)), (prefix + ".running_mean", (c,)),
            (prefix + ".running_var", (c,)), (prefix + ".num_batches_tracked", ())]


def resnet_spec(arch: str, num_classes: int = 1000) -> List[Tuple[str, Tuple[int, ...]]]:
    arch = ARCH_ALIASES.get(arch, arch)
    kind, layers = _RESNET_CFG[arch]
    exp = 4 if kind == "bottleneck" else 1
    spec = [("conv1.weight", (64, 3, 7, 7))] + _bn_spec("bn1", 64)
    inplanes = 64
    for li, (planes, blocks) in enumerate(zip([64, 128, 256, 512], layers)):
        for b in range(blocks):
            stride = 2 if (b == 0 and li > 0) else 1
            p = "layer%d.%d" % (li + 1, b)
            if kind == "bottleneck":
                spec += [(p + ".conv1.weight", (planes, inplanes, 1, 1))] + _bn_spec(p + ".bn1", planes)
                spec += [(p + ".conv2.weight", (planes, planes, 3, 3))] + _bn_spec(p + ".bn2", planes)
                spec += [(p + ".conv3.weight", (planes * 4, planes, 1, 1))] + _bn_spec(p + ".bn3", planes * 4)
            else:
                spec += [(p + ".conv1.weight", (planes, inplanes, 3, 3))] + _bn_spec(p + ".bn1", planes)
                spec += [(p + ".conv2.weight", (planes, planes, 3, 3))] + _bn_spec(p + ".bn2", planes)
            if stride != 1 or inplanes != planes * exp:
                spec += [(p + ".downsample.0.weight", (planes * exp, inplanes, 1, 1))] + _bn_spec(p + ".downsample.1", planes * exp)
            inplanes = planes * exp
    spec += [("fc.weight", (num_classes, 512 * exp)), ("fc.bias", (num_classes,))]
    return spec


def random_state_dict(spec, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights keyed by parameter name (there are no checkpoints offline).
    Conv: Kaiming-normal fan_out; BN: non-trivial affine + running stats (so the folding is exercised),
    the last BN of each residual branch damped so activations stay O(1) through 16+ blocks."""
    import zlib
    out = {}
    for key, shape in spec:
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) & 0x7FFFFFFF)
        if key.endswith("num_batches_tracked"):
            t = torch.tensor(0, dtype=torch.long)
        elif len(shape) == 4:
            fan_out = shape[0] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * (2.0 / fan_out) ** 0.5
        elif key.endswith("running_var"):
            t = torch.rand(shape, generator=g) + 0.5
        elif key.endswith("running_mean"):
            t = torch.randn(shape, generator=g) * 0.1
        elif ".bn" in key or key.startswith("bn") or "downsample.1" in key:
            if key.endswith("weight"):
                t = torch.rand(shape, generator=g) * 0.5 + 0.5
                if key.endswith("bn3.weight") or (key.endswith(".bn2.weight") and _is_basic_last(key, spec)):
                    t = t * 0.3
            else:
                t = torch.randn(shape, generator=g) * 0.1
        elif len(shape) == 2:
            bound = 1.0 / shape[1] ** 0.5
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        out[key] = t
    return out


def _is_basic_last(key, spec):
    prefix = key.rsplit(".bn2.weight", 1)[0]
    return not any(k.startswith(prefix + ".conv3") for k, _ in spec)


# ------------------------------------------------------------------------------------------------
def _strip_prefix(sd):
    """Checkpoints come as {"model": sd} or bare, with optional module./base_model. prefixes
    (benchmark_eval_adv.py:162-168, base_benchmark_eval_adv.py:166-176)."""
    if "model" in sd and isinstance(sd["model"], dict):
        sd = sd["model"]
    out = {}
    for k, v in sd.items():
        for pre in ("module.", "base_model."):
            if k.startswith(pre):
                k = k[len(pre):]
        out[k] = v
    return out


class _ConvBN:
    """conv (bias-free) + folded BN: weight planes [2, Cout, KH, KW, Cin], scale/bias float32 [Cout]."""

    def __init__(self, sd, conv, bn, device, stride, pad):
        w = sd[conv + ".weight"].float()
        gamma, beta = sd[bn + ".weight"].double(), sd[bn + ".bias"].double()
        mean, var = sd[bn + ".running_mean"].double(), sd[bn + ".running_var"].double()
        scale = gamma / torch.sqrt(var + BN_EPS)
        self.scale = scale.float().to(device).contiguous()
        self.bias = (beta - mean * scale).float().to(device).contiguous()
        self.w = ops.split_f32(w.permute(0, 2, 3, 1).contiguous().to(device))
        self.stride, self.pad = stride, pad

    def __call__(self, x, act=None, res=None, passes=3):
        return ops.conv2d_nhwc(x, self.w, self.scale, self.bias, res, stride=self.stride, pad=self.pad, act=act,
                               passes=passes)


class ResNet:
    def __init__(self, arch: str, state_dict: Dict[str, torch.Tensor], device, passes: int = 3):
        arch = ARCH_ALIASES.get(arch, arch)
        self.arch, self.device, self.passes = arch, torch.device(device), passes
        kind, layers = _RESNET_CFG[arch]
        sd = _strip_prefix(state_dict)
        dev = self.device
        # stem: 7x7/s2 as a GEMM over im2col'd patches, K = (ky, kx, c) padded 147 -> 192
        w = sd["conv1.weight"].float().permute(0, 2, 3, 1).reshape(64, 147)
        wp = torch.zeros(64, 192)
        wp[:, :147] = w
        self.stem_w = ops.split_f32(wp.to(dev).contiguous())
        g, b = sd["bn1.weight"].double(), sd["bn1.bias"].double()
        m, v = sd["bn1.running_mean"].double(), sd["bn1.running_var"].double()
        s = g / torch.sqrt(v + BN_EPS)
        self.stem_scale, self.stem_bias = s.float().to(dev), (b - m * s).float().to(dev)
        self.blocks = []
        for li, blocks in enumerate(layers):
            for bi in range(blocks):
                p = "layer%d.%d" % (li + 1, bi)
                stride = 2 if (bi == 0 and li > 0) else 1
                blk = {"kind": kind}
                if kind == "bottleneck":
                    blk["c1"] = _ConvBN(sd, p + ".conv1", p + ".bn1", dev, 1, 0)
                    blk["c2"] = _ConvBN(sd, p + ".conv2", p + ".bn2", dev, stride, 1)
                    blk["c3"] = _ConvBN(sd, p + ".conv3", p + ".bn3", dev, 1, 0)
                else:
                    blk["c1"] = _ConvBN(sd, p + ".conv1", p + ".bn1", dev, stride, 1)
                    blk["c2"] = _ConvBN(sd, p + ".conv2", p + ".bn2", dev, 1, 1)
                if (p + ".downsample.0.weight") in sd:
                    blk["down"] = _ConvBN(sd, p + ".downsample.0", p + ".downsample.1", dev, stride, 0)
                self.blocks.append(blk)
        self.fc_w = ops.split_f32(sd["fc.weight"].float().to(dev).contiguous())
        self.fc_b = sd["fc.bias"].float().to(dev).contiguous()
        self.num_classes = self.fc_w.shape[1]
        self._graphs = {}

    # -- eager launch sequence -------------------------------------------------------------------
    def forward(self, images: torch.Tensor, logits: Optional[torch.Tensor] = None) -> torch.Tensor:
        """images: uint8 NHWC [n,h,w,3] (raw pixels; ToTensor+Normalize fused into the stem gather) or
        float32 NCHW in [0,1] (attack path).  Returns float32 logits [n, classes]."""
        n = images.shape[0]
        h, w = (images.shape[1], images.shape[2]) if images.dtype == torch.uint8 else (images.shape[2], images.shape[3])
        P = self.passes
        if images.dtype == torch.uint8:
            # raw pixels: gather + ToTensor + Normalize + split fused into the stem GEMM's operand producer
            x = ops.stem_conv7x7_u8(images, self.stem_w, self.stem_scale, self.stem_bias, act="relu", passes=P)
        else:
            cols = ops.stem_im2col(images)
            x = ops.linear(cols, self.stem_w, self.stem_scale, self.stem_bias, act="relu", passes=P)
            x = x.view(2, n, h // 2, w // 2, 64)
        x = ops.maxpool3x3s2(x)
        for blk in self.blocks:
            idn = blk["down"](x, passes=P) if "down" in blk else x
            if blk["kind"] == "bottleneck":
                o = blk["c1"](x, act="relu", passes=P)
                o = blk["c2"](o, act="relu", passes=P)
                x = blk["c3"](o, act="relu", res=idn, passes=P)
            else:
                o = blk["c1"](x, act="relu", passes=P)
                x = blk["c2"](o, act="relu", res=idn, passes=P)
        pooled = ops.global_avgpool(x)
        if logits is None:
            logits = torch.empty((n, self.num_classes), dtype=torch.float32, device=self.device)
        ops.linear(pooled, self.fc_w, None, self.fc_b, passes=P, out_f32=logits, want_planes=False)
        return logits

    __call__ = forward

    # -- CUDA-graph replay for a fixed input shape -----------------------------------------------
    def graphed(self, example: torch.Tensor):
        """Returns fn(images) -> logits replaying one captured graph (static input/output buffers)."""
        key = (tuple(example.shape), example.dtype)
        if key not in self._graphs:
            static_in = example.clone()
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    self.forward(static_in)   # warm-up: cudaFuncSetAttribute, table builds, allocator
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_out = self.forward(static_in)
            self._graphs[key] = (g, static_in, static_out)
        g, static_in, static_out = self._graphs[key]

        def run(images, copy_in=True):
            if copy_in and images.data_ptr() != static_in.data_ptr():
                static_in.copy_(images)
            g.replay()
            return static_out
        run.static_in, run.static_out = static_in, static_out
        return run

    def launches_per_forward(self) -> int:
        n = 1 + 1 + 1 + 1  # fused stem, maxpool, avgpool, fc
        for blk in self.blocks:
            n += (3 if blk["kind"] == "bottleneck" else 2) + (1 if "down" in blk else 0)
        return n


def build_model(arch: str, state_dict=None, device="cuda", passes: int = 3, seed: int = 0):
    arch = ARCH_ALIASES.get(arch, arch)
    if arch not in _RESNET_CFG:
        raise NotImplementedError("architecture %r has no B200 kernel path yet (ResNet family only)" % arch)
    if state_dict is None:
        state_dict = random_state_dict(resnet_spec(arch), seed)
    return ResNet(arch, state_dict, device, passes)
