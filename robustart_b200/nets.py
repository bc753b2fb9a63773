"""Classifier forward passes on the B200 kernels (eval mode).

ResNet-18/50 in the reference's layout (prototype/prototype/model/resnet_official.py:40-140,221-239,
330-346 == torchvision: stride on the 3x3 of a Bottleneck, bias-free convs, BatchNorm after every conv).
A model is built from a plain state_dict with the reference's keys; BatchNorm (eval) is folded into a
per-channel scale/bias applied in the GEMM epilogue, weights are re-laid out once to [Cout, KH, KW, Cin]
split-fp16 planes.  Activations stay NHWC split-fp16 planes end to end; every layer is one launch of a
kernel in libb200robust.so; the whole forward can be captured into one CUDA graph.
"""
from __future__ import annotations

import os
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


GRAD_SCALE = 4096.0          # loss scale of every input-gradient pass: activations AND gradients live in fp16-coded planes (one
                             # plane, or a hi/lo pair), and a classifier's gradients sit at 1e-4 .. 1e-10, below fp16's normal range
GRAD_SCALE_F16 = GRAD_SCALE


class _ConvBN:
    """conv (bias-free) + folded BN: weight planes [2, Cout, KH, KW, Cin], scale/bias float32 [Cout]."""

    def __init__(self, sd, conv, bn, device, stride, pad, f16=False):
        self.f16 = f16
        w = sd[conv + ".weight"].float()
        gamma, beta = sd[bn + ".weight"].double(), sd[bn + ".bias"].double()
        mean, var = sd[bn + ".running_mean"].double(), sd[bn + ".running_var"].double()
        scale = gamma / torch.sqrt(var + BN_EPS)
        # the BN scale is folded into the weights (fp64 product, then the usual hi/lo split): a scale-free
        # epilogue lets the kernel add the residual on the tensor core (identity k-blocks) instead of the LSU
        self.scale = None
        self.bias = (beta - mean * scale).float().to(device).contiguous()
        w = (w.double() * scale.view(-1, 1, 1, 1)).float()
        self.w = ops.to_planes(w.permute(0, 2, 3, 1).contiguous().to(device), f16)
        self.stride, self.pad = stride, pad
        self.k = w.shape[-1]
        self._w_folded, self._w_dgrad, self._w_dgrad_s2, self._device = w, None, None, device

    def dgrad(self, dy, res=None, mask=None, passes=3):
        """Input gradient of the convolution: dy planes [2,n,ho,wo,cout] -> [mask > 0] * (planes [2,n,h,w,cin] + res).
        conv^T = stride-1 convolution of the (zero-dilated, for stride 2) output gradient with the weights
        transposed in (cout, cin) and flipped in (ky, kx); padding k-1-pad.  mask = post-ReLU activation the
        gradient flows into (ReLU backward fused into the GEMM epilogue)."""
        if self._w_dgrad is None:
            wt = self._w_folded.flip(2, 3).permute(1, 2, 3, 0).contiguous()       # [cin, ky', kx', cout]
            self._w_dgrad = ops.to_planes(wt.to(self._device), self.f16)
        if self.stride not in (1, 2):
            raise NotImplementedError("dgrad for stride %d" % self.stride)
        if self.stride == 2 and self.k == 1:
            # 1x1/s2 (downsample branch): contract on the small map, then zero-insert
            assert res is None and mask is None
            return ops.dilate2(ops.conv2d_dgrad(dy, self._w_dgrad, passes=passes))
        if self.stride == 2 and self.k == 3 and self.pad == 1 and os.environ.get("B200R_DGRAD_S2", "parity") == "parity":
            # parity classes of the output only meet 1, 2, 2 and 4 of the nine flipped taps: four small stride-1 convolutions of dy
            # itself, each written into its quarter of dx (no zero-dilated tensor, a quarter of the MMAs)
            if self._w_dgrad_s2 is None:
                wt = self._w_folded.flip(2, 3).permute(1, 2, 3, 0)                 # [cin, ky', kx', cout]
                S = ([1], [0, 2])
                self._w_dgrad_s2 = [ops.to_planes(wt[:, S[a]][:, :, S[b]].contiguous().to(self._device), self.f16) for a in (0, 1) for b in (0, 1)]
            return ops.conv2d_dgrad3x3s2(dy, self._w_dgrad_s2, res, mask, passes=passes)
        if self.stride == 2:
            dy = ops.dilate2(dy)
        return ops.conv2d_dgrad(dy, self._w_dgrad, res, mask, pad=self.k - 1 - self.pad, passes=passes)

    def dgrad_acc(self, dy, dx, mask=None, passes=3):
        """1x1 / stride 2 only: dx[..., 2i, 2j, :] += (this convolution's input gradient of dy), in place, masked like dx was."""
        assert self.stride == 2 and self.k == 1
        if self._w_dgrad is None:
            wt = self._w_folded.flip(2, 3).permute(1, 2, 3, 0).contiguous()       # [cin, 1, 1, cout]
            self._w_dgrad = ops.to_planes(wt.to(self._device), self.f16)
        return ops.conv2d_dgrad1x1s2_acc(dy, self._w_dgrad, dx, mask, passes=passes)

    def __call__(self, x, act=None, res=None, passes=3):
        return ops.conv2d_nhwc(x, self.w, self.scale, self.bias, res, stride=self.stride, pad=self.pad, act=act,
                               passes=passes)


class ResNet:
    def __init__(self, arch: str, state_dict: Dict[str, torch.Tensor], device, passes: int = 3):
        arch = ARCH_ALIASES.get(arch, arch)
        self.arch, self.device, self.passes = arch, torch.device(device), passes
        # passes = ops.PASSES_F16: activations / weights are one fp16 plane, one MMA per product (include/b200r.h)
        self.f16 = f16 = (passes == ops.PASSES_F16)
        kind, layers = _RESNET_CFG[arch]
        sd = _strip_prefix(state_dict)
        dev = self.device
        # stem: 7x7/s2 as a GEMM over im2col'd patches, K = (ky, 8 kx slots, c) = 168 padded to 192
        self.stem_w = ops.to_planes(ops.pack_stem_weight(sd["conv1.weight"]).to(dev).contiguous(), f16)
        g, b = sd["bn1.weight"].double(), sd["bn1.bias"].double()
        m, v = sd["bn1.running_mean"].double(), sd["bn1.running_var"].double()
        s = g / torch.sqrt(v + BN_EPS)
        self.stem_scale, self.stem_bias = s.float().to(dev), (b - m * s).float().to(dev)
        # one-launch stem (fp16 mode): BN scale folded into the weights before the single rounding to fp16
        self.stem_w_folded = ops.to_planes(ops.pack_stem_weight((sd["conv1.weight"].double() * s.view(-1, 1, 1, 1)).float()).to(dev).contiguous(), True) if f16 else None
        # one-launch stem (split mode): exact pixels against hi/lo weights with everything folded in on the host
        self.stem_wp, self.stem_osc = (None, 1.0) if f16 else ops.stem_pool_split_prepare(sd["conv1.weight"], self.stem_scale, self.stem_bias, dev)
        # ... and its float-input twin (the attack path): pixel as an fp16 hi/lo pair, arg-max codes out
        self.stem_wpf, self.stem_oscf = (None, 1.0) if f16 else ops.stem_pool_split_prepare(sd["conv1.weight"], self.stem_scale, self.stem_bias, dev,
                                                                                            f32_input=True)
        self.fused_stem_f32 = os.environ.get("B200R_STEM_POOL_F32", "1") != "0"
        self.blocks = []
        for li, blocks in enumerate(layers):
            for bi in range(blocks):
                p = "layer%d.%d" % (li + 1, bi)
                stride = 2 if (bi == 0 and li > 0) else 1
                blk = {"kind": kind}
                if kind == "bottleneck":
                    blk["c1"] = _ConvBN(sd, p + ".conv1", p + ".bn1", dev, 1, 0, f16)
                    blk["c2"] = _ConvBN(sd, p + ".conv2", p + ".bn2", dev, stride, 1, f16)
                    blk["c3"] = _ConvBN(sd, p + ".conv3", p + ".bn3", dev, 1, 0, f16)
                else:
                    blk["c1"] = _ConvBN(sd, p + ".conv1", p + ".bn1", dev, stride, 1, f16)
                    blk["c2"] = _ConvBN(sd, p + ".conv2", p + ".bn2", dev, 1, 1, f16)
                if (p + ".downsample.0.weight") in sd:
                    blk["down"] = _ConvBN(sd, p + ".downsample.0", p + ".downsample.1", dev, stride, 0, f16)
                self.blocks.append(blk)
        self.fc_w = ops.to_planes(sd["fc.weight"].float().to(dev).contiguous(), f16)
        self.fc_b = sd["fc.bias"].float().to(dev).contiguous()
        self.num_classes = self.fc_w.shape[1]
        self._graphs = {}
        self.fused_stem_pool = os.environ.get("B200R_STEM_POOL", "1") != "0"   # A/B switch for the one-launch stem

    # -- eager launch sequence -------------------------------------------------------------------
    def forward(self, images: torch.Tensor, logits: Optional[torch.Tensor] = None) -> torch.Tensor:
        """images: uint8 NHWC [n,h,w,3] (raw pixels; ToTensor+Normalize fused into the stem gather) or
        float32 NCHW in [0,1] (attack path).  Returns float32 logits [n, classes]."""
        n = images.shape[0]
        h, w = (images.shape[1], images.shape[2]) if images.dtype == torch.uint8 else (images.shape[2], images.shape[3])
        P = self.passes
        if images.dtype == torch.uint8 and self.fused_stem_pool and ops.stem_pool_ok(h, w):
            # raw pixels -> conv1 + bn1 + relu + maxpool in one launch (overlapping-descriptor implicit im2col)
            x = ops.stem_pool_u8(images, self.stem_w_folded, self.stem_bias) if self.f16 else ops.stem_pool_u8_split(images, self.stem_wp, self.stem_osc)
        else:
            if images.dtype != torch.uint8 and not self.f16 and self.fused_stem_pool and self.fused_stem_f32 and ops.stem_pool_ok(h, w):
                x, _ = ops.stem_pool_f32_split(images.contiguous(), self.stem_wpf, self.stem_oscf)     # float iterate: one launch too
            else:
                if images.dtype == torch.uint8:
                    # raw pixels: gather + ToTensor + Normalize + split fused into the stem GEMM's operand producer
                    x = ops.stem_conv7x7_u8(images, self.stem_w, self.stem_scale, self.stem_bias, act="relu", passes=P)
                else:
                    x = self._stem_f32(images, P)
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

    def _stem_f32(self, x01, P):
        """conv1 + bn1 + relu from a float32 NCHW image: fused producer when the geometry allows it (W <= 256, W % 16 == 0),
        else im2col + GEMM."""
        n, _, h, w = x01.shape
        if w <= 256 and w % 16 == 0 and h % 2 == 0:
            return ops.stem_conv7x7_f32(x01.contiguous(), self.stem_w, self.stem_scale, self.stem_bias, act="relu", passes=P)
        cols = ops.stem_im2col(x01)
        return ops.linear(cols, self.stem_w, self.stem_scale, self.stem_bias, act="relu", passes=P).view(-1, n, h // 2, w // 2, 64)

    __call__ = forward

    # -- forward that keeps what the input-gradient pass needs, and that pass ------------------------------
    def forward_saved(self, x01: torch.Tensor):
        """float32 NCHW [0,1] images -> (logits, saved activations).  Same launch sequence as forward()."""
        n, _, h, w = x01.shape
        P = self.passes
        if not self.f16 and self.fused_stem_pool and self.fused_stem_f32 and ops.stem_pool_ok(h, w):
            # conv1 + bn1 + relu + maxpool + arg-max codes in one launch: the 112 x 112 activation is never written
            x, codes = ops.stem_pool_f32_split(x01.contiguous(), self.stem_wpf, self.stem_oscf)
            saved = {"shape": (n, h, w), "stem": None, "pool_codes": codes, "stem_hw": (h // 2, w // 2), "blocks": []}
            return self._body_saved(x, saved, n, P)
        s0 = self._stem_f32(x01, P)
        if self.f16:
            x = ops.maxpool3x3s2(s0)
            saved = {"shape": (n, h, w), "stem": s0, "blocks": []}
        else:
            # the pool also leaves its arg-max codes (with the stem ReLU's backward folded in): the gradient pass routes through them and
            # never reads the 112 x 112 activation again, which is not kept
            x, codes = ops.maxpool3x3s2_codes(s0)
            saved = {"shape": (n, h, w), "stem": None, "pool_codes": codes, "stem_hw": (s0.shape[2], s0.shape[3]), "blocks": []}
        return self._body_saved(x, saved, n, P)

    def _body_saved(self, x, saved, n, P):
        for blk in self.blocks:
            idn = blk["down"](x, passes=P) if "down" in blk else x
            if blk["kind"] == "bottleneck":
                a1 = blk["c1"](x, act="relu", passes=P)
                a2 = blk["c2"](a1, act="relu", passes=P)
                y = blk["c3"](a2, act="relu", res=idn, passes=P)
                saved["blocks"].append((a1, a2, y))
            else:
                a1 = blk["c1"](x, act="relu", passes=P)
                y = blk["c2"](a1, act="relu", res=idn, passes=P)
                saved["blocks"].append((a1, y))
            x = y
        pooled = ops.global_avgpool(x)
        logits = torch.empty((n, self.num_classes), dtype=torch.float32, device=self.device)
        ops.linear(pooled, self.fc_w, None, self.fc_b, passes=P, out_f32=logits, want_planes=False)
        return logits, saved

    @staticmethod
    def block_backward(blk, sv, g, P=3, in_mask=None, g_is_masked=False):
        """Gradient w.r.t. a residual block's input from the gradient w.r.t. its (post-ReLU) output.
        sv = the block's saved activations (a1[, a2], y), all post-ReLU (resnet_official.py:40-140).  Every ReLU
        backward is fused into the dgrad GEMM that produces the masked tensor; `in_mask` (the block's input
        activation = the previous block's output) applies the PREVIOUS block's output ReLU to the result, so the
        chain needs no stand-alone masking pass (g_is_masked: the caller did that for this block's own output)."""
        dz = g if g_is_masked else ops.relu_bwd(g, sv[-1])
        # identity path: the 1x1 downsample's dgrad or dz itself.  A stride-2 downsample is accumulated IN PLACE into the main branch's
        # result afterwards (b200r_conv2d_dgrad1x1s2_acc_nhwc): no zero-inserted tensor, no residual operand (B200R_DS_ACC=0: old path)
        acc = "down" in blk and blk["down"].stride == 2 and os.environ.get("B200R_DS_ACC", "1") != "0"
        r = None if acc else (blk["down"].dgrad(dz, passes=P) if "down" in blk else dz)
        if blk["kind"] == "bottleneck":
            a1, a2, _ = sv
            t = blk["c3"].dgrad(dz, mask=a2, passes=P)
            t = blk["c2"].dgrad(t, mask=a1, passes=P)
        else:
            a1, _ = sv
            t = blk["c2"].dgrad(dz, mask=a1, passes=P)
        out = blk["c1"].dgrad(t, res=r, mask=in_mask, passes=P)
        if acc:
            blk["down"].dgrad_acc(dz, out, mask=in_mask, passes=P)
        return out

    def input_grad(self, dlogits: torch.Tensor, saved, passes: Optional[int] = None) -> torch.Tensor:
        """d loss / d x01 (float32 NCHW) from d loss / d logits (float32 [n, classes]) and forward_saved()'s state.
        Only input gradients are formed (what value_and_grad / autograd.grad(loss, x) return to the attacks)."""
        P = self.passes if passes is None else passes
        f16 = self.f16
        n, h, w = saved["shape"]
        if not hasattr(self, "_fc_wt"):
            self._fc_wt = ops.to_planes(ops.from_planes(self.fc_w).t().contiguous(), f16)    # [2048|512, classes]
            wt = ops.from_planes(self.stem_w) * self.stem_scale.view(-1, 1)                  # BN scale folded, [64, 192]
            # the stem GEMM's gradient runs on single fp16 planes in BOTH precisions: its [n*112*112, 192] output is the largest tensor
            # of the whole pass (1.2 GB as a hi/lo pair at batch 128) and only feeds col2im; 11 bits on the last contraction leave the
            # image gradient's direction unchanged (cosine > 0.9999, tests/test_backward_gpu.py)
            self._stem_wt = ops.to_planes(wt.t().contiguous(), True)                         # [192, 64]
        # fp16 planes: the pass runs on S * dlogits (gradients of a classifier sit at 1e-4 .. 1e-10, below fp16's
        # normal range) and 1/S is folded into the last kernel; the attacks only use the sign / direction anyway
        S = GRAD_SCALE
        g = ops.linear(ops.to_planes(dlogits.contiguous(), f16, S), self._fc_wt, passes=P)   # [P, n, c]
        last = saved["blocks"][-1][-1]
        g = ops.global_avgpool_bwd(g, last.shape[2], last.shape[3])
        g = ops.relu_bwd(g, last)                        # the last block's output ReLU; all others are fused
        for i in range(len(self.blocks) - 1, -1, -1):
            in_mask = saved["blocks"][i - 1][-1] if i > 0 else None     # block input = previous block's output
            g = self.block_backward(self.blocks[i], saved["blocks"][i], g, P, in_mask=in_mask, g_is_masked=True)
        if saved.get("pool_codes") is not None:
            g = ops.maxpool3x3s2_bwd_codes_hi(saved["pool_codes"], g, *saved["stem_hw"])
        else:
            g = ops.maxpool3x3s2_relu_bwd_hi(saved["stem"], g)   # max-pool and stem-ReLU backward in one pass, one fp16 plane out
        dcols = ops.linear(g.view(1, -1, 64), self._stem_wt, passes=ops.PASSES_F16)          # -> [1, n*ho*wo, 192]
        return ops.stem_col2im(dcols, n, h, w, unscale=1.0 / S)

    def loss_and_input_grad(self, x01: torch.Tensor, labels: torch.Tensor, reduction: str = "sum"):
        """(per-sample CE losses, d CE / d x01): the one call an attack step makes."""
        logits, saved = self.forward_saved(x01)
        loss, dlogits = ops.ce_loss_grad(logits, labels, 1.0 if reduction == "sum" else 1.0 / x01.shape[0])
        return loss, self.input_grad(dlogits, saved), logits

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
        # stem + maxpool (one launch with the fused stem), avgpool, fc
        n = (1 if self.fused_stem_pool else 2) + 1 + 1
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


# ================================================================================================
# Token models: ViT-B/16 (prototype/prototype/model/vision_transformer.py:198-349,420-436; config
# model_config.py:216-225 -> representation_size 768) and MLP-Mixer-B/16 (vit/mlp_mixer.py:7-159).
# ================================================================================================
def vit_spec(depth=12, dim=768, mlp=3072, patch=16, img=224, classes=1000, rep=768):
    np_ = (img // patch) ** 2
    spec = [("pos_embedding", (1, np_ + 1, dim)), ("cls_token", (1, 1, dim)),
            ("embedding.weight", (dim, 3, patch, patch)), ("embedding.bias", (dim,))]
    for d in range(depth):
        p = "transformer.encoders.encoder_%d." % d
        spec += [(p + "norm1.weight", (dim,)), (p + "norm1.bias", (dim,)),
                 (p + "attention.to_qkv.weight", (3 * dim, dim)), (p + "attention.to_qkv.bias", (3 * dim,)),
                 (p + "attention.to_out.weight", (dim, dim)), (p + "attention.to_out.bias", (dim,)),
                 (p + "norm2.weight", (dim,)), (p + "norm2.bias", (dim,)),
                 (p + "feedforward.mlp1.weight", (mlp, dim)), (p + "feedforward.mlp1.bias", (mlp,)),
                 (p + "feedforward.mlp2.weight", (dim, mlp)), (p + "feedforward.mlp2.bias", (dim,))]
    spec += [("transformer.encoder_norm.weight", (dim,)), ("transformer.encoder_norm.bias", (dim,))]
    if rep:
        spec += [("pre_logits.weight", (rep, dim)), ("pre_logits.bias", (rep,))]
    spec += [("head.weight", (classes, rep or dim)), ("head.bias", (classes,))]
    return spec


def mixer_spec(depth=12, dim=768, patch=16, img=224, classes=1000, token_ratio=0.5, channel_ratio=4.0):
    np_ = (img // patch) ** 2
    tok, ch = round(dim * token_ratio), round(dim * channel_ratio)
    spec = [("patch_embed.proj.weight", (dim, 3, patch, patch)), ("patch_embed.proj.bias", (dim,))]
    for d in range(depth):
        p = "blocks.%d." % d
        spec += [(p + "norm1.weight", (dim,)), (p + "norm1.bias", (dim,)), (p + "norm2.weight", (dim,)), (p + "norm2.bias", (dim,)),
                 (p + "token_mix.fc1.weight", (tok, np_)), (p + "token_mix.fc1.bias", (tok,)),
                 (p + "token_mix.fc2.weight", (np_, tok)), (p + "token_mix.fc2.bias", (np_,)),
                 (p + "channel_mix.fc1.weight", (ch, dim)), (p + "channel_mix.fc1.bias", (ch,)),
                 (p + "channel_mix.fc2.weight", (dim, ch)), (p + "channel_mix.fc2.bias", (dim,))]
    spec += [("norm.weight", (dim,)), ("norm.bias", (dim,)), ("head.weight", (classes, dim)), ("head.bias", (classes,))]
    return spec


def random_token_state_dict(spec, seed=0):
    """Synthetic weights for the token models: Linear ~ N(0, 1/fan_in) so activations keep O(1) scale through 12
    blocks, LayerNorm affine and biases non-trivial, embeddings N(0, 0.02)."""
    import zlib
    out = {}
    for key, shape in spec:
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) & 0x7FFFFFFF)
        if key in ("pos_embedding", "cls_token"):
            t = torch.randn(shape, generator=g) * 0.02
        elif len(shape) == 4:
            t = torch.randn(shape, generator=g) * (1.0 / (shape[1] * shape[2] * shape[3])) ** 0.5
        elif len(shape) == 2:
            t = torch.randn(shape, generator=g) * (1.0 / shape[1]) ** 0.5
        elif "norm" in key and key.endswith("weight"):
            t = torch.rand(shape, generator=g) * 0.5 + 0.75
        else:
            t = torch.randn(shape, generator=g) * 0.05
        out[key] = t
    return out


class _Lin:
    def __init__(self, sd, name, device, k_pad=None, n_pad=None):
        w, b = sd[name + ".weight"].float(), sd[name + ".bias"].float()
        if k_pad or n_pad:
            wp = torch.zeros(n_pad or w.shape[0], k_pad or w.shape[1])
            wp[:w.shape[0], :w.shape[1]] = w
            bp = torch.zeros(n_pad or w.shape[0])
            bp[:b.shape[0]] = b
            w, b = wp, bp
        self.w = ops.split_f32(w.contiguous().to(device))
        self.b = b.contiguous().to(device)

    def __call__(self, x, act=None, res=None, passes=3, out_f32=None):
        if out_f32 is not None:
            return ops.linear(x, self.w, None, self.b, res, act=act, passes=passes, out_f32=out_f32, want_planes=False)
        return ops.linear(x, self.w, None, self.b, res, act=act, passes=passes)

    def keep_pre(self, x, act, passes=3):
        """(act(pre), pre) for the gradient pass: one launch in split precision (b200r_linear_keep_pre), Linear + b200r_act_planes
        in the fp16 mode (or with B200R_KEEP_PRE=0)."""
        if passes == 3 and x.shape[0] == 2 and os.environ.get("B200R_KEEP_PRE", "1") != "0":
            return ops.linear_keep_pre(x, self.w, self.b, act=act)
        pre = ops.linear(x, self.w, None, self.b, None, act=None, passes=passes)
        return ops.act_planes(pre, act), pre

    def dgrad(self, g, passes=3):
        """Input gradient of the Linear: g planes [2, m, nout] -> planes [2, m, k] = g W (a GEMM with the transposed weights)."""
        if getattr(self, "_wt", None) is None:
            self._wt = ops.split_f32(ops.merge_f32(self.w).t().contiguous())
        return ops.linear(g, self._wt, passes=passes)


class _TokenModel:
    def __init__(self, device, passes):
        self.device, self.passes = torch.device(device), passes
        self._graphs = {}

    graphed = ResNet.graphed
    __call__ = lambda self, images, logits=None: self.forward(images, logits)


class ViT(_TokenModel):
    def __init__(self, state_dict, device, passes=3, depth=12, dim=768, heads=12, patch=16):
        super().__init__(device, passes)
        sd = _strip_prefix(state_dict)
        dev = self.device
        self.arch, self.depth, self.dim, self.heads, self.patch = "vit_b16_224", depth, dim, heads, patch
        self.embed = _Lin({"e.weight": sd["embedding.weight"].reshape(dim, -1), "e.bias": sd["embedding.bias"]}, "e", dev)
        self.cls = sd["cls_token"].float().reshape(-1).contiguous().to(dev)
        self.pos = sd["pos_embedding"].float().reshape(-1, dim).contiguous().to(dev)
        self.blocks = []
        for d in range(depth):
            p = "transformer.encoders.encoder_%d." % d
            self.blocks.append(dict(
                n1=(sd[p + "norm1.weight"].float().to(dev), sd[p + "norm1.bias"].float().to(dev)),
                n2=(sd[p + "norm2.weight"].float().to(dev), sd[p + "norm2.bias"].float().to(dev)),
                qkv=_Lin(sd, p + "attention.to_qkv", dev), out=_Lin(sd, p + "attention.to_out", dev),
                m1=_Lin(sd, p + "feedforward.mlp1", dev), m2=_Lin(sd, p + "feedforward.mlp2", dev)))
        self.norm = (sd["transformer.encoder_norm.weight"].float().to(dev), sd["transformer.encoder_norm.bias"].float().to(dev))
        self.pre = _Lin(sd, "pre_logits", dev) if "pre_logits.weight" in sd else None
        self.head = _Lin(sd, "head", dev)
        self.num_classes = sd["head.weight"].shape[0]

    def forward(self, images, logits=None):
        n = images.shape[0]
        P = self.passes
        x = self.embed(ops.patch_gather(images, self.patch), passes=P)              # [n*196, 768]
        npatch = x.shape[1] // n
        x = ops.assemble_tokens(x, self.cls, self.pos, n, npatch)                   # [n*197, 768]
        T = npatch + 1
        scale = (self.dim // self.heads) ** -0.5
        for b in self.blocks:
            y = ops.layernorm(x, *b["n1"], eps=1e-5)
            y = ops.attention(b["qkv"](y, passes=P), n, T, self.heads, self.dim // self.heads, scale)
            x = b["out"](y, res=x, passes=P)                                        # x = attn(norm1(x)) + x
            y = ops.layernorm(x, *b["n2"], eps=1e-5)
            y = b["m1"](y, act="gelu_tanh", passes=P)                               # tanh-approx GELU (:19-37)
            x = b["m2"](y, res=x, passes=P)
        x = ops.layernorm(x, *self.norm, eps=1e-5)
        cls = x.view(2, n, T, self.dim)[:, :, 0].contiguous()                       # x[:, 0]
        if self.pre is not None:
            cls = self.pre(cls, act="tanh", passes=P)
        if logits is None:
            logits = torch.empty((n, self.num_classes), dtype=torch.float32, device=self.device)
        self.head(cls, passes=P, out_f32=logits)
        return logits

    def launches_per_forward(self):
        return 2 + 1 + self.depth * 7 + 1 + 1 + (1 if self.pre else 0) + 1

    # -- forward that keeps what the input-gradient pass needs, and that pass (opt-in: B200R_NATIVE_TOKEN_GRAD=1) ----------
    def forward_saved(self, x01: torch.Tensor):
        """float32 NCHW [0,1] images -> (logits, saved).  Same layers as forward(); the two activations whose derivative needs
        the pre-activation (GELU in the MLP, tanh in pre_logits) run as a plain Linear + b200r_act_planes."""
        n, _, h, w = x01.shape
        P = self.passes
        x = self.embed(ops.patch_gather(x01.contiguous(), self.patch), passes=P)
        npatch = x.shape[1] // n
        x = ops.assemble_tokens(x, self.cls, self.pos, n, npatch)
        T = npatch + 1
        hd = self.dim // self.heads
        saved = {"shape": (n, h, w, T), "blocks": []}
        for b in self.blocks:
            x_in = x
            qkv = b["qkv"](ops.layernorm(x, *b["n1"], eps=1e-5), passes=P)
            x_mid = b["out"](ops.attention(qkv, n, T, self.heads, hd, hd ** -0.5), res=x, passes=P)
            a, pre = b["m1"].keep_pre(ops.layernorm(x_mid, *b["n2"], eps=1e-5), "gelu_tanh", passes=P)
            x = b["m2"](a, res=x_mid, passes=P)
            saved["blocks"].append((x_in, qkv, x_mid, pre))
        saved["final"] = x
        cls = ops.layernorm(x, *self.norm, eps=1e-5).view(2, n, T, self.dim)[:, :, 0].contiguous()
        if self.pre is not None:
            saved["pre_logits"] = self.pre(cls, passes=P)
            cls = ops.act_planes(saved["pre_logits"], "tanh")
        logits = torch.empty((n, self.num_classes), dtype=torch.float32, device=self.device)
        self.head(cls, passes=P, out_f32=logits)
        return logits, saved

    def input_grad(self, dlogits: torch.Tensor, saved, passes: Optional[int] = None) -> torch.Tensor:
        """d loss / d x01 (float32 NCHW) from d loss / d logits; input gradients only (vision_transformer.py:44-349 reversed)."""
        P = self.passes if passes is None else passes
        n, h, w, T = saved["shape"]
        hd = self.dim // self.heads
        S = GRAD_SCALE                                                                # unscaled again by patch_scatter
        g = self.head.dgrad(ops.to_planes(dlogits.contiguous(), False, S), P)         # [2, n, rep]
        if self.pre is not None:
            g = self.pre.dgrad(ops.act_bwd_planes(g, saved["pre_logits"], "tanh"), P)
        # encoder_norm only feeds x[:, 0]: LayerNorm is row-wise, so its backward runs on the class rows alone
        x_cls = saved["final"].view(2, n, T, self.dim)[:, :, 0].contiguous()
        g_cls = ops.layernorm_bwd(g, x_cls, self.norm[0], eps=1e-5)
        g = torch.zeros((2, n, T, self.dim), dtype=g_cls.dtype, device=g_cls.device)   # int16 planes: bf16 zero = 0x0000
        g[:, :, 0] = g_cls
        g = g.view(2, n * T, self.dim)
        for b, (x_in, qkv, x_mid, pre) in zip(reversed(self.blocks), reversed(saved["blocks"])):
            t = b["m1"].dgrad(ops.act_bwd_planes(b["m2"].dgrad(g, P), pre, "gelu_tanh"), P)
            g = ops.layernorm_bwd(t, x_mid, b["n2"][0], eps=1e-5, add=g)               # x = x_mid + mlp(norm2(x_mid))
            t = ops.attention_bwd(qkv, b["out"].dgrad(g, P), n, T, self.heads, hd, hd ** -0.5)
            g = ops.layernorm_bwd(b["qkv"].dgrad(t, P), x_in, b["n1"][0], eps=1e-5, add=g)  # x_mid = x_in + attn(norm1(x_in))
        g = g.view(2, n, T, self.dim)[:, :, 1:].contiguous().view(2, n * (T - 1), self.dim)   # drop the class token
        return ops.patch_scatter(self.embed.dgrad(g, P), n, h, w, self.patch, unscale=1.0 / S)


class Mixer(_TokenModel):
    T_PAD = 256   # token dimension padded so that the [b*c, t] rows are 16-byte aligned and K is a multiple of 64

    def __init__(self, state_dict, device, passes=3, depth=12, dim=768, patch=16):
        super().__init__(device, passes)
        sd = _strip_prefix(state_dict)
        dev = self.device
        self.arch, self.depth, self.dim, self.patch = "mixer_b16_224", depth, dim, patch
        self.embed = _Lin({"e.weight": sd["patch_embed.proj.weight"].reshape(dim, -1), "e.bias": sd["patch_embed.proj.bias"]}, "e", dev)
        self.blocks = []
        for d in range(depth):
            p = "blocks.%d." % d
            tok = sd[p + "token_mix.fc1.weight"].shape[0]
            self.blocks.append(dict(
                n1=(sd[p + "norm1.weight"].float().to(dev), sd[p + "norm1.bias"].float().to(dev)),
                n2=(sd[p + "norm2.weight"].float().to(dev), sd[p + "norm2.bias"].float().to(dev)),
                t1=_Lin(sd, p + "token_mix.fc1", dev, k_pad=self.T_PAD), t2=_Lin(sd, p + "token_mix.fc2", dev, n_pad=self.T_PAD),
                c1=_Lin(sd, p + "channel_mix.fc1", dev), c2=_Lin(sd, p + "channel_mix.fc2", dev), tok=tok))
        self.norm = (sd["norm.weight"].float().to(dev), sd["norm.bias"].float().to(dev))
        self.head = _Lin(sd, "head", dev)
        self.num_classes = sd["head.weight"].shape[0]

    def forward(self, images, logits=None):
        n = images.shape[0]
        P = self.passes
        x = self.embed(ops.patch_gather(images, self.patch), passes=P)              # [n*196, 768]
        T = x.shape[1] // n
        for b in self.blocks:
            y = ops.layernorm(x, *b["n1"], eps=1e-6)                                # vit_base.py:159
            y = ops.tokens_to_channels(y, n, T, self.dim, self.T_PAD)               # [n*768, 256] (zero padded)
            y = b["t1"](y, act="gelu_erf", passes=P)                                # nn.GELU (erf)
            y = b["t2"](y, passes=P)                                                # [n*768, 256], columns >= 196 are 0
            x = ops.channels_to_tokens_add(y, x, n, T, self.dim, self.T_PAD)        # x + token_mix(...)^T
            y = ops.layernorm(x, *b["n2"], eps=1e-6)
            y = b["c1"](y, act="gelu_erf", passes=P)
            x = b["c2"](y, res=x, passes=P)
        x = ops.layernorm(x, *self.norm, eps=1e-6)
        pooled = ops.global_avgpool(x.view(2, n, T, 1, self.dim))                   # x.mean(dim=1)
        if logits is None:
            logits = torch.empty((n, self.num_classes), dtype=torch.float32, device=self.device)
        self.head(pooled, passes=P, out_f32=logits)
        return logits

    def launches_per_forward(self):
        return 2 + self.depth * 8 + 1 + 1 + 1

    # -- forward that keeps what the input-gradient pass needs, and that pass (opt-in: B200R_NATIVE_TOKEN_GRAD=1) ----------
    def forward_saved(self, x01: torch.Tensor):
        """float32 NCHW [0,1] images -> (logits, saved).  Same layers as forward(); the GELUs run as a plain Linear +
        b200r_act_planes so that the pre-activations are available to the gradient pass."""
        n, _, h, w = x01.shape
        P = self.passes
        x = self.embed(ops.patch_gather(x01.contiguous(), self.patch), passes=P)
        T = x.shape[1] // n
        saved = {"shape": (n, h, w, T), "blocks": []}
        for b in self.blocks:
            x_in = x
            y = ops.tokens_to_channels(ops.layernorm(x, *b["n1"], eps=1e-6), n, T, self.dim, self.T_PAD)
            a1, p1 = b["t1"].keep_pre(y, "gelu_erf", passes=P)
            y = b["t2"](a1, passes=P)
            x_mid = ops.channels_to_tokens_add(y, x, n, T, self.dim, self.T_PAD)
            a2, p2 = b["c1"].keep_pre(ops.layernorm(x_mid, *b["n2"], eps=1e-6), "gelu_erf", passes=P)
            x = b["c2"](a2, res=x_mid, passes=P)
            saved["blocks"].append((x_in, p1, x_mid, p2))
        saved["final"] = x
        x = ops.layernorm(x, *self.norm, eps=1e-6)
        pooled = ops.global_avgpool(x.view(2, n, T, 1, self.dim))
        logits = torch.empty((n, self.num_classes), dtype=torch.float32, device=self.device)
        self.head(pooled, passes=P, out_f32=logits)
        return logits, saved

    def input_grad(self, dlogits: torch.Tensor, saved, passes: Optional[int] = None) -> torch.Tensor:
        """d loss / d x01 (float32 NCHW) from d loss / d logits; input gradients only (mlp_mixer.py:7-159 reversed)."""
        P = self.passes if passes is None else passes
        n, h, w, T = saved["shape"]
        S = GRAD_SCALE                                                                # unscaled again by patch_scatter
        g = self.head.dgrad(ops.to_planes(dlogits.contiguous(), False, S), P)         # [2, n, dim]
        g = ops.global_avgpool_bwd(g, T, 1).view(2, n * T, self.dim)                  # x.mean(dim=1) backward
        g = ops.layernorm_bwd(g, saved["final"], self.norm[0], eps=1e-6)
        zeros = torch.zeros_like(g)                                                   # residual operand of the plain transpose
        for b, (x_in, p1, x_mid, p2) in zip(reversed(self.blocks), reversed(saved["blocks"])):
            t = b["c1"].dgrad(ops.act_bwd_planes(b["c2"].dgrad(g, P), p2, "gelu_erf"), P)
            g = ops.layernorm_bwd(t, x_mid, b["n2"][0], eps=1e-6, add=g)               # x = x_mid + channel_mix(norm2(x_mid))
            t = ops.tokens_to_channels(g, n, T, self.dim, self.T_PAD)                  # [n*dim, T_PAD]
            t = b["t1"].dgrad(ops.act_bwd_planes(b["t2"].dgrad(t, P), p1, "gelu_erf"), P)
            t = ops.channels_to_tokens_add(t, zeros, n, T, self.dim, self.T_PAD)
            g = ops.layernorm_bwd(t, x_in, b["n1"][0], eps=1e-6, add=g)                # x_mid = x_in + token_mix(norm1(x_in))
        return ops.patch_scatter(self.embed.dgrad(g, P), n, h, w, self.patch, unscale=1.0 / S)


_TOKEN_ARCHS = {"vit_b16_224": (ViT, vit_spec), "vit_base_patch16_224": (ViT, vit_spec), "mixer_b16_224": (Mixer, mixer_spec)}
_build_resnet = build_model


def build_model(arch: str, state_dict=None, device="cuda", passes: int = 3, seed: int = 0):  # noqa: F811
    if arch in _TOKEN_ARCHS:
        cls, spec = _TOKEN_ARCHS[arch]
        if state_dict is None:
            state_dict = random_token_state_dict(spec(), seed)
        return cls(state_dict, device, passes)
    return _build_resnet(arch, state_dict, device, passes, seed)


# ================================================================================================
# Mobile families: MobileNetV2 x1.0 (prototype/prototype/model/mobilenet_v2.py:80-202) and EfficientNet-B0
# (prototype/prototype/model/efficientnet.py:91-125,289-369,372-495).  Pointwise convs run on the tcgen05
# GEMM (K tails zero-filled by TMA), depthwise / SE / pooling on the CUDA-core kernels of mobile_layers.cu.
# ================================================================================================
_MBV2_SETTING = [[1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1], [6, 160, 3, 2], [6, 320, 1, 1]]
# (repeat, kernel, stride, expand, in, out) of efficientnet.py:101-109 at width/depth 1.0
_EFFB0_BLOCKS = [(1, 3, 1, 1, 32, 16), (2, 3, 2, 6, 16, 24), (2, 5, 2, 6, 24, 40), (3, 3, 2, 6, 40, 80), (3, 5, 1, 6, 80, 112),
                 (4, 5, 2, 6, 112, 192), (1, 3, 1, 6, 192, 320)]


def mobilenet_v2_spec(classes=1000):
    spec = [("features.0.0.weight", (32, 3, 3, 3))] + _bn_spec("features.0.1", 32)
    cin, idx = 32, 1
    for t, c, n, s in _MBV2_SETTING:
        for i in range(n):
            hid = cin * t
            p, j = "features.%d.conv." % idx, 0
            if t != 1:
                spec += [(p + "0.0.weight", (hid, cin, 1, 1))] + _bn_spec(p + "0.1", hid)
                j = 1
            spec += [(p + "%d.0.weight" % j, (hid, 1, 3, 3))] + _bn_spec(p + "%d.1" % j, hid)
            spec += [(p + "%d.weight" % (j + 1), (c, hid, 1, 1))] + _bn_spec(p + "%d" % (j + 2), c)
            cin, idx = c, idx + 1
    spec += [("features.%d.0.weight" % idx, (1280, cin, 1, 1))] + _bn_spec("features.%d.1" % idx, 1280)
    spec += [("classifier.1.weight", (classes, 1280)), ("classifier.1.bias", (classes,))]
    return spec


def efficientnet_b0_spec(classes=1000):
    spec, bi = [], 0
    for rep, k, s, e, cin, cout in _EFFB0_BLOCKS:
        for r in range(rep):
            ci = cin if r == 0 else cout
            hid, p, j = ci * e, "blocks.%d." % bi, 0
            if e != 1:
                spec += [(p + "in_conv.0.weight", (hid, ci, 1, 1))] + _bn_spec(p + "in_conv.1", hid)
                j = 3
            spec += [(p + "in_conv.%d.weight" % j, (hid, 1, k, k))] + _bn_spec(p + "in_conv.%d" % (j + 1), hid)
            se = max(1, int(ci * 0.25))
            spec += [(p + "se_block.conv1.weight", (se, hid, 1, 1)), (p + "se_block.conv1.bias", (se,)),
                     (p + "se_block.conv2.weight", (hid, se, 1, 1)), (p + "se_block.conv2.bias", (hid,))]
            spec += [(p + "out_conv.0.weight", (cout, hid, 1, 1))] + _bn_spec(p + "out_conv.1", cout)
            bi += 1
    spec += [("stem.0.weight", (32, 3, 3, 3))] + _bn_spec("stem.1", 32)
    spec += [("head.0.weight", (1280, 320, 1, 1))] + _bn_spec("head.1", 1280)
    spec += [("fc.weight", (classes, 1280)), ("fc.bias", (classes,))]
    return spec


def _fold_bn(sd, bn, device):
    g, b = sd[bn + ".weight"].double(), sd[bn + ".bias"].double()
    m, v = sd[bn + ".running_mean"].double(), sd[bn + ".running_var"].double()
    s = g / torch.sqrt(v + BN_EPS)
    return s.float().to(device).contiguous(), (b - m * s).float().to(device).contiguous()


class _PW:
    """1x1 conv + folded BN on the tensor-core GEMM."""

    def __init__(self, sd, conv, bn, device):
        w = sd[conv + ".weight"].float()
        scale, self.bias = _fold_bn(sd, bn, device)
        self.scale = None                                    # folded into the weights (see _ConvBN)
        w = (w.reshape(w.shape[0], -1).double() * scale.double().cpu().view(-1, 1)).float()
        self.w = ops.split_f32(w.reshape(w.shape[0], 1, 1, w.shape[1]).contiguous().to(device))
        # narrow inputs (the expansion / projection layers at 112 x 112 and 56 x 56): pure output streaming, no use for a tensor-core tile
        cin = w.shape[1]
        self.w_f32 = w.contiguous().to(device) if (cin % 8 == 0 and cin <= 32 and os.environ.get("B200R_PW_SMALLK", "1") != "0") else None

    def __call__(self, x, act=None, res=None, passes=3):
        if self.w_f32 is not None and x.shape[0] == 2:
            return ops.pointwise_smallk(x, self.w_f32, self.bias, res, act=act)
        return ops.conv2d_nhwc(x, self.w, self.scale, self.bias, res, act=act, passes=passes)

    def dgrad(self, dy, res=None, passes=3):
        """Input gradient of the 1x1 convolution (BN scale folded): dy planes [2,n,h,w,cout] -> planes [2,n,h,w,cin] (+ res)."""
        if getattr(self, "_wt", None) is None:
            w = ops.merge_f32(self.w)                                  # [cout, 1, 1, cin]
            self._wt = ops.split_f32(w.permute(3, 1, 2, 0).contiguous())   # [cin, 1, 1, cout]
        return ops.conv2d_dgrad(dy, self._wt, res, None, pad=0, passes=passes)


class _DW:
    def __init__(self, sd, conv, bn, device, stride):
        w = sd[conv + ".weight"].float()                     # [c, 1, k, k]
        self.k, self.stride = w.shape[-1], stride
        self.w = w.reshape(w.shape[0], -1).t().contiguous().to(device)   # [k*k, c]
        self.scale, self.bias = _fold_bn(sd, bn, device)

    def __call__(self, x, act):
        return ops.dwconv_nhwc(x, self.w, self.scale, self.bias, k=self.k, stride=self.stride, pad=self.k // 2, act=act)

    def dgrad(self, dpre):
        """Input gradient of depthwise conv + BN from the gradient w.r.t. the BN output: the same kernel on flipped taps with the BN
        scale folded in; a stride-2 layer first re-inserts the zeros (pad k//2 = k-1-k//2 for odd k, so the padding is unchanged)."""
        if getattr(self, "_wb", None) is None:
            self._wb = (self.w.flip(0) * self.scale.view(1, -1)).contiguous()      # taps reversed = flipped in (ky, kx)
            self._one, self._zero = torch.ones_like(self.scale), torch.zeros_like(self.bias)
        if self.stride == 2:
            dpre = ops.dilate2(dpre)
        return ops.dwconv_nhwc(dpre, self._wb, self._one, self._zero, k=self.k, stride=1, pad=self.k // 2, act=None)


class _ImageStem:
    """3x3/s2 conv from the image: one direct fp32 kernel (ops.image_stem3x3s2); B200R_IMAGE_STEM=gemm keeps the former
    small im2col (K = 27 -> 32) + GEMM pair for A/B measurements."""

    def __init__(self, sd, conv, bn, device, act):
        w = sd[conv + ".weight"].float()                     # [32, 3, 3, 3]
        w27 = w.permute(0, 2, 3, 1).reshape(w.shape[0], 27)
        wp = torch.zeros(w.shape[0], 32)
        wp[:, :27] = w27
        self.w = ops.split_f32(wp.to(device).contiguous())
        self.w27 = w27.contiguous().to(device)
        self.scale, self.bias = _fold_bn(sd, bn, device)
        self.act, self.cout = act, w.shape[0]
        self.direct = os.environ.get("B200R_IMAGE_STEM", "direct") != "gemm" and self.cout % 8 == 0 and self.cout <= 64

    def dgrad(self, dpre, n, h, w, unscale):
        """d loss / d image (float32 NCHW) from the gradient w.r.t. the stem's BN output."""
        if getattr(self, "_w27s", None) is None:
            self._w27s = (self.w27 * self.scale.view(-1, 1)).contiguous()
        return ops.image_stem3x3s2_bwd(dpre, self._w27s, n, h, w, unscale=unscale)

    def pre(self, images):
        """conv + BN without the activation (the gradient pass needs swish's pre-activation)."""
        return ops.image_stem3x3s2(images, self.w27, self.scale, self.bias, act=None)

    def __call__(self, images, passes):
        if self.direct:
            return ops.image_stem3x3s2(images, self.w27, self.scale, self.bias, act=self.act)
        n = images.shape[0]
        cols, ho, wo = ops.image_im2col(images, 3, 2, 1, 32)
        x = ops.linear(cols, self.w, self.scale, self.bias, act=self.act, passes=passes)
        return x.view(2, n, ho, wo, self.cout)


class MobileNetV2(_TokenModel):
    def __init__(self, state_dict, device, passes=3):
        super().__init__(device, passes)
        sd, dev = _strip_prefix(state_dict), self.device
        self.arch = "mobilenet_v2"
        self.stem = _ImageStem(sd, "features.0.0", "features.0.1", dev, "relu6")
        self.blocks, cin, idx = [], 32, 1
        for t, c, n, s in _MBV2_SETTING:
            for i in range(n):
                stride, p, j = (s if i == 0 else 1), "features.%d.conv." % idx, 0
                blk = {"res": stride == 1 and cin == c}
                if t != 1:
                    blk["pw"] = _PW(sd, p + "0.0", p + "0.1", dev)
                    j = 1
                blk["dw"] = _DW(sd, p + "%d.0" % j, p + "%d.1" % j, dev, stride)
                blk["pl"] = _PW(sd, p + "%d" % (j + 1), p + "%d" % (j + 2), dev)
                self.blocks.append(blk)
                cin, idx = c, idx + 1
        self.last = _PW(sd, "features.%d.0" % idx, "features.%d.1" % idx, dev)
        self.fc = _Lin(sd, "classifier.1", dev)
        self.num_classes = sd["classifier.1.weight"].shape[0]

    def forward(self, images, logits=None):
        n, P = images.shape[0], self.passes
        x = self.stem(images, P)
        for b in self.blocks:
            y = b["pw"](x, act="relu6", passes=P) if "pw" in b else x
            y = b["dw"](y, "relu6")
            x = b["pl"](y, res=x if b["res"] else None, passes=P)      # linear bottleneck (+ skip)
        x = self.last(x, act="relu6", passes=P)
        pooled = ops.global_avgpool(x)
        if logits is None:
            logits = torch.empty((n, self.num_classes), dtype=torch.float32, device=self.device)
        self.fc(pooled, passes=P, out_f32=logits)
        return logits

    def launches_per_forward(self):
        return 2 + sum(2 + (1 if "pw" in b else 0) for b in self.blocks) + 3

    # -- forward that keeps what the input-gradient pass needs, and that pass (mobilenet_v2.py:31-202 reversed) -------------------
    def forward_saved(self, x01: torch.Tensor):
        """float32 NCHW [0,1] images -> (logits, saved).  Same launches as forward(); every ReLU6 output is kept (its derivative is a
        function of the output)."""
        n, _, h, w = x01.shape
        P = self.passes
        x = self.stem(x01.contiguous(), P)
        saved = {"shape": (n, h, w), "stem": x, "blocks": []}
        for b in self.blocks:
            a1 = b["pw"](x, act="relu6", passes=P) if "pw" in b else None
            a2 = b["dw"](a1 if a1 is not None else x, "relu6")
            x = b["pl"](a2, res=x if b["res"] else None, passes=P)
            saved["blocks"].append((a1, a2))
        last = self.last(x, act="relu6", passes=P)
        saved["last"] = last
        pooled = ops.global_avgpool(last)
        logits = torch.empty((n, self.num_classes), dtype=torch.float32, device=self.device)
        self.fc(pooled, passes=P, out_f32=logits)
        return logits, saved

    def input_grad(self, dlogits: torch.Tensor, saved, passes: Optional[int] = None) -> torch.Tensor:
        P = self.passes if passes is None else passes
        n, h, w = saved["shape"]
        S = GRAD_SCALE
        last = saved["last"]
        g = self.fc.dgrad(ops.to_planes(dlogits.contiguous(), False, S), P)                 # [2, n, 1280]
        g = ops.global_avgpool_bwd(g, last.shape[2], last.shape[3])
        g = self.last.dgrad(ops.act_bwd_planes(g, last, "relu6"), passes=P)
        for b, (a1, a2) in zip(reversed(self.blocks), reversed(saved["blocks"])):
            t = b["pl"].dgrad(g, passes=P)                                                  # linear bottleneck: no activation
            t = b["dw"].dgrad(ops.act_bwd_planes(t, a2, "relu6"))
            if a1 is not None:
                g = b["pw"].dgrad(ops.act_bwd_planes(t, a1, "relu6"), res=g if b["res"] else None, passes=P)
            else:
                g = ops.planes_add(t, g) if b["res"] else t
        return self.stem.dgrad(ops.act_bwd_planes(g, saved["stem"], "relu6"), n, h, w, 1.0 / S)


class EfficientNetB0(_TokenModel):
    def __init__(self, state_dict, device, passes=3):
        super().__init__(device, passes)
        sd, dev = _strip_prefix(state_dict), self.device
        self.arch = "efficientnet_b0"
        self.stem = _ImageStem(sd, "stem.0", "stem.1", dev, "swish")
        self.blocks, bi = [], 0
        for rep, k, s, e, cin, cout in _EFFB0_BLOCKS:
            for r in range(rep):
                ci, stride = (cin if r == 0 else cout), (s if r == 0 else 1)
                p, j = "blocks.%d." % bi, 0
                blk = {"res": stride == 1 and ci == cout}
                if e != 1:
                    blk["pw"] = _PW(sd, p + "in_conv.0", p + "in_conv.1", dev)
                    j = 3
                blk["dw"] = _DW(sd, p + "in_conv.%d" % j, p + "in_conv.%d" % (j + 1), dev, stride)
                hid = ci * e
                se = sd[p + "se_block.conv1.weight"].shape[0]
                sep = (se + 7) // 8 * 8                              # pad the squeeze width to a 16-byte row
                blk["se1"] = _Lin({"a.weight": sd[p + "se_block.conv1.weight"].reshape(se, hid), "a.bias": sd[p + "se_block.conv1.bias"]},
                                  "a", dev, n_pad=sep)
                blk["se2"] = _Lin({"a.weight": sd[p + "se_block.conv2.weight"].reshape(hid, se), "a.bias": sd[p + "se_block.conv2.bias"]},
                                  "a", dev, k_pad=sep)
                blk["pl"] = _PW(sd, p + "out_conv.0", p + "out_conv.1", dev)
                self.blocks.append(blk)
                bi += 1
        self.head = _PW(sd, "head.0", "head.1", dev)
        self.fc = _Lin(sd, "fc", dev)
        self.num_classes = sd["fc.weight"].shape[0]

    def forward(self, images, logits=None):
        n, P = images.shape[0], self.passes
        x = self.stem(images, P)
        for b in self.blocks:
            y = b["pw"](x, act="swish", passes=P) if "pw" in b else x
            y = b["dw"](y, "swish")
            w = ops.global_avgpool(y)                                 # squeeze
            w = b["se2"](b["se1"](w, act="swish", passes=P), act="sigmoid", passes=P)
            y = ops.channel_scale(y, w)                               # excite
            x = b["pl"](y, res=x if b["res"] else None, passes=P)
        x = self.head(x, act="swish", passes=P)
        pooled = ops.global_avgpool(x)
        if logits is None:
            logits = torch.empty((n, self.num_classes), dtype=torch.float32, device=self.device)
        self.fc(pooled, passes=P, out_f32=logits)
        return logits

    def launches_per_forward(self):
        return 2 + sum(6 + (1 if "pw" in b else 0) for b in self.blocks) + 3

    # -- forward that keeps what the input-gradient pass needs, and that pass (efficientnet.py:289-495 reversed) -------------------
    def forward_saved(self, x01: torch.Tensor):
        """float32 NCHW [0,1] images -> (logits, saved).  swish and sigmoid need their PRE-activation, so every activated layer runs
        as (conv + BN) then b200r_act_planes here, instead of the fused epilogue of forward()."""
        n, _, h, w = x01.shape
        P = self.passes
        p0 = self.stem.pre(x01.contiguous())
        x = ops.act_planes(p0, "swish")
        saved = {"shape": (n, h, w), "stem": p0, "blocks": []}
        for b in self.blocks:
            x_in = x
            ppw = b["pw"](x, passes=P) if "pw" in b else None
            y = ops.act_planes(ppw, "swish") if ppw is not None else x
            pdw = b["dw"](y, None)
            y = ops.act_planes(pdw, "swish")
            p1 = b["se1"](ops.global_avgpool(y), passes=P)
            p2 = b["se2"](ops.act_planes(p1, "swish"), passes=P)
            s = ops.act_planes(p2, "sigmoid")
            x = b["pl"](ops.channel_scale(y, s), res=x_in if b["res"] else None, passes=P)
            saved["blocks"].append((ppw, pdw, y, p1, p2, s))
        ph = self.head(x, passes=P)
        saved["head"] = ph
        pooled = ops.global_avgpool(ops.act_planes(ph, "swish"))
        logits = torch.empty((n, self.num_classes), dtype=torch.float32, device=self.device)
        self.fc(pooled, passes=P, out_f32=logits)
        return logits, saved

    def input_grad(self, dlogits: torch.Tensor, saved, passes: Optional[int] = None) -> torch.Tensor:
        P = self.passes if passes is None else passes
        n, h, w = saved["shape"]
        S = GRAD_SCALE
        ph = saved["head"]
        g = self.fc.dgrad(ops.to_planes(dlogits.contiguous(), False, S), P)
        g = ops.global_avgpool_bwd(g, ph.shape[2], ph.shape[3])
        g = self.head.dgrad(ops.act_bwd_planes(g, ph, "swish"), passes=P)
        for b, (ppw, pdw, y, p1, p2, s) in zip(reversed(self.blocks), reversed(saved["blocks"])):
            dz = b["pl"].dgrad(g, passes=P)                                                # gradient w.r.t. y * s
            # squeeze-excite: y * s with s = sigmoid(se2(swish(se1(mean(y))))) (efficientnet.py:338-360)
            ds = ops.channel_dot(dz, y, s.shape[-1])
            t = b["se2"].dgrad(ops.act_bwd_planes(ds, p2, "sigmoid"), P)
            t = b["se1"].dgrad(ops.act_bwd_planes(t, p1, "swish"), P)                      # [2, n, hid]
            dy = ops.planes_add(ops.channel_scale(dz, s), ops.global_avgpool_bwd(t, y.shape[2], y.shape[3]))
            t = b["dw"].dgrad(ops.act_bwd_planes(dy, pdw, "swish"))
            if ppw is not None:
                g = b["pw"].dgrad(ops.act_bwd_planes(t, ppw, "swish"), res=g if b["res"] else None, passes=P)
            else:
                g = ops.planes_add(t, g) if b["res"] else t
        return self.stem.dgrad(ops.act_bwd_planes(g, saved["stem"], "swish"), n, h, w, 1.0 / S)


_MOBILE_ARCHS = {"mobilenet_v2": (MobileNetV2, mobilenet_v2_spec), "mobilenet_v2_x1_0": (MobileNetV2, mobilenet_v2_spec),
                 "efficientnet_b0": (EfficientNetB0, efficientnet_b0_spec)}
_build_prev = build_model


def build_model(arch: str, state_dict=None, device="cuda", passes: int = 3, seed: int = 0):  # noqa: F811
    if arch in _MOBILE_ARCHS:
        cls, spec = _MOBILE_ARCHS[arch]
        if state_dict is None:
            state_dict = random_state_dict(spec(), seed)
        return cls(state_dict, device, passes)
    return _build_prev(arch, state_dict, device, passes, seed)
