"""Python view of the C-ABI model handles (include/b200r.h: b200r_model_*; csrc/model_handle.cu).

The handle is what a C caller uses instead of model_entry() + nn.Module (prototype/prototype/model/__init__.py:292-332,
resnet_official.py:330-346): it is built from the reference's state_dict tensors, owns its device weights and activation arena,
and runs forward / input-gradient passes as launches of the library's kernels -- no Python layer sequencing.  This class only
marshals pointers; robustart_b200.nets.ResNet issues the same launches from Python and must agree bit for bit
(tests/test_model_handle_gpu.py)."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _lib, nets, ops

ARCH_IDS = {"resnet18": 0, "resnet34": 1, "resnet50": 2, "resnet101": 3, "vit_b16_224": 4, "vit_base_patch16_224": 4, "mixer_b16_224": 5, "mobilenet_v2": 6,
            "efficientnet_b0": 7}


class _Weight(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("numel", C.c_int64)]


class ModelHandle:
    def __init__(self, arch: str, state_dict: Dict[str, torch.Tensor], device, passes: int = 3):
        arch = nets.ARCH_ALIASES.get(arch, arch)
        if arch not in ARCH_IDS:
            raise NotImplementedError("model handles cover the ResNet family, ViT-B/16, MLP-Mixer-B/16, MobileNetV2 and EfficientNet-B0, not %r" % arch)
        self.arch, self.device, self.passes = arch, torch.device(device), passes
        sd = nets._strip_prefix(state_dict)
        keep = [(k, v.detach().to("cpu", torch.float32).contiguous()) for k, v in sd.items() if not k.endswith("num_batches_tracked")]
        arr = (_Weight * len(keep))(*[_Weight(k.encode(), v.data_ptr(), v.numel()) for k, v in keep])
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().b200r_model_create(ARCH_IDS[arch], C.cast(arr, C.c_void_p), len(keep), passes, C.byref(self._h)))
        self.num_classes = _lib.load().b200r_model_num_classes(self._h)
        self.f16 = passes == ops.PASSES_F16
        self._graphs = {}
        if ARCH_IDS[arch] >= 6:
            self.feat = 1280
            self._launches = {6: 2 + 17 * 2 + 16 + 3, 7: 2 + 16 * 6 + 15 + 3}[ARCH_IDS[arch]]     # stem, blocks (pw? dw [se x4] pl), last, pool, fc
            return
        if ARCH_IDS[arch] == 5:
            # patch gather, embedding, 12 x (2 layernorm + 2 transposes + 4 linear), final norm, token mean, head
            self.feat = 768
            self._launches = 2 + 12 * 8 + 3
            return
        if ARCH_IDS[arch] == 4:
            # patch gather, embedding, token assembly, 12 x (2 layernorm + 4 linear + attention), final norm, class-token copies, head
            self.feat = 768
            self._launches = 3 + 12 * 7 + 1 + 2 + 1 + (1 if any(k.startswith("pre_logits") for k, _ in keep) else 0)
            return
        self.feat = 2048 if arch in ("resnet50", "resnet101") else 512
        kind, layers = nets._RESNET_CFG[arch]
        per_block = 3 if kind == "bottleneck" else 2
        # stem (+ maxpool unless fused), blocks, downsample convs (one per stage; ResNet-18/34's first stage has none), avgpool, fc
        self._launches = 1 + per_block * sum(layers) + (4 if kind == "bottleneck" else 3) + 2

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.load().b200r_model_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def reserve(self, n, h=224, w=224, for_input_grad=False):
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().b200r_model_reserve(self._h, n, h, w, int(for_input_grad)))

    def forward(self, images: torch.Tensor, logits: torch.Tensor = None) -> torch.Tensor:
        """uint8 NHWC [n,h,w,3] raw pixels, or float32 NCHW in [0,1] (keeps the activations for input_grad)."""
        n = images.shape[0]
        if logits is None:
            logits = torch.empty((n, self.num_classes), dtype=torch.float32, device=self.device)
        lib = _lib.load()
        with torch.cuda.device(self.device):
            if images.dtype == torch.uint8:
                ops._need_cuda(images, torch.uint8, "images")
                _lib.check(lib.b200r_model_forward_u8(self._h, images.data_ptr(), logits.data_ptr(), n, images.shape[1], images.shape[2], ops._stream()))
            else:
                ops._need_cuda(images, torch.float32, "images")
                _lib.check(lib.b200r_model_forward_f32(self._h, images.data_ptr(), logits.data_ptr(), n, images.shape[2], images.shape[3], ops._stream()))
        return logits

    __call__ = forward

    def forward_vjp(self, x01: torch.Tensor):
        """(logits, fn: dlogits -> d loss / d x01): the attack loops' model interface (attacks.forward_vjp)."""
        x01 = x01.detach().contiguous()
        logits = self.forward(x01)

        def vjp(dlogits):
            dx = torch.empty_like(x01)
            with torch.cuda.device(self.device):
                _lib.check(_lib.load().b200r_model_input_grad(self._h, dlogits.float().contiguous().data_ptr(), dx.data_ptr(), ops._stream()))
            return dx
        return logits, vjp

    bounds = (0, 1)

    graphed = nets.ResNet.graphed          # CUDA-graph replay for a fixed input shape (the arena is sized by the warm-up forwards)

    def launches_per_forward(self) -> int:
        return self._launches
