"""torch-tensor front end of the C-ABI: device memory and streams come from PyTorch, every
computation is a hand-written sm_100a kernel in libb200robust.so."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _lib

CORRUPTION_NAMES = (
    "gaussian_noise", "shot_noise", "impulse_noise", "defocus_blur", "glass_blur", "motion_blur",
    "zoom_blur", "snow", "frost", "fog", "brightness", "contrast", "elastic_transform", "pixelate",
    "jpeg_compression", "speckle_noise", "gaussian_blur", "spatter", "saturate",
)  # order of corruption_tuple, RobustART/noise/utils/imagenet_c/__init__.py:5-8
CORRUPTION_IDS = {n: i for i, n in enumerate(CORRUPTION_NAMES)}

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # benchmark_eval_adv.py:33-38
IMAGENET_STD = (0.229, 0.224, 0.225)

_workspaces = {}
_frost_loaded = {}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda(t: torch.Tensor, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError("%s must be a CUDA tensor (no CPU fallback)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)


def workspace(nbytes: int, device) -> torch.Tensor:
    key = torch.device(device).index
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def corruption_id(corruption: Union[int, str]) -> int:
    if isinstance(corruption, str):
        if corruption not in CORRUPTION_IDS:
            raise KeyError(corruption)
        return CORRUPTION_IDS[corruption]
    cid = int(corruption)
    if not 0 <= cid < len(CORRUPTION_NAMES):
        raise IndexError("corruption_number %d not in [0, 18]" % cid)
    return cid


def ext_noise_count(corruption, severity: int, n: int, h: int, w: int) -> int:
    lib = _lib.load()
    out = C.c_size_t(0)
    _lib.check(lib.b200r_corrupt_ext_noise_count(corruption_id(corruption), severity, n, h, w, C.byref(out)))
    return out.value


def ensure_frost_textures(device):
    """Upload the frost textures once per device (procedural stand-ins unless the user supplied the
    reference's frost1..6 files via robustart_b200.assets.set_frost_dir)."""
    key = torch.device(device).index
    if key in _frost_loaded:
        return
    from .assets import frost_textures
    lib = _lib.load()
    keep = []
    with torch.cuda.device(device):
        for slot, tex in enumerate(frost_textures()):
            t = torch.from_numpy(np.ascontiguousarray(tex)).to(device)
            _lib.check(lib.b200r_set_frost_texture(slot, t.data_ptr(), t.shape[0], t.shape[1]))
            keep.append(t)
    _frost_loaded[key] = keep  # the library borrows the device memory: keep it alive


def corrupt_u8(images: torch.Tensor, corruption: Union[int, str], severity: int = 1, *, seed: int = 0,
               image_offset: int = 0, ext_noise: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ImageNet-C corruption of a uint8 NHWC CUDA batch (b200r_corrupt_u8)."""
    _need_cuda(images, torch.uint8, "images")
    if images.dim() != 4 or images.shape[-1] != 3:
        raise ValueError("images must be [n,h,w,3] uint8")
    n, h, w, _ = images.shape
    cid = corruption_id(corruption)
    lib = _lib.load()
    if out is None:
        out = torch.empty_like(images)
    else:
        _need_cuda(out, torch.uint8, "out")
    if cid == CORRUPTION_IDS["frost"]:
        ensure_frost_textures(images.device)
    if ext_noise is not None:
        _need_cuda(ext_noise, torch.float32, "ext_noise")
        need = ext_noise_count(cid, severity, n, h, w)
        if ext_noise.numel() != need:
            raise ValueError("ext_noise has %d values, corruption needs %d" % (ext_noise.numel(), need))
    with torch.cuda.device(images.device):
        nbytes = C.c_size_t(0)
        _lib.check(lib.b200r_corrupt_workspace_bytes(cid, severity, n, h, w, C.byref(nbytes)))
        ws = workspace(nbytes.value, images.device) if nbytes.value else None
        _lib.check(lib.b200r_corrupt_u8(cid, severity, images.data_ptr(), out.data_ptr(), n, h, w,
                                        seed & (2 ** 64 - 1), image_offset, _ptr(ext_noise), _ptr(ws),
                                        nbytes.value, _stream()))
    return out


def u8nhwc_to_f32nchw(images: torch.Tensor, mean=IMAGENET_MEAN, std=IMAGENET_STD,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(images, torch.uint8, "images")
    n, h, w, _ = images.shape
    if out is None:
        out = torch.empty((n, 3, h, w), dtype=torch.float32, device=images.device)
    with torch.cuda.device(images.device):
        _lib.check(_lib.load().b200r_u8nhwc_to_f32nchw(images.data_ptr(), out.data_ptr(), n, h, w,
                                                       _lib.f3(mean), _lib.f3(std), _stream()))
    return out


def normalize(x: torch.Tensor, mode: str = "normal", mean=IMAGENET_MEAN, std=IMAGENET_STD,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """normalize(x) / normalize(x, 'inv') of benchmark_eval_adv.py:33-46 on float32 NCHW."""
    _need_cuda(x, torch.float32, "x")
    n, c, h, w = x.shape
    assert c == 3
    if out is None:
        out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_normalize_f32nchw(x.data_ptr(), out.data_ptr(), n, h, w, _lib.f3(mean),
                                                       _lib.f3(std), {"normal": 0, "inv": 1, "grad": 2}[mode], _stream()))
    return out


def random_start_linf(x0: torch.Tensor, eps: float, *, seed: int = 0, image_offset: int = 0,
                      u: Optional[torch.Tensor] = None, clip01: bool = True) -> torch.Tensor:
    _need_cuda(x0, torch.float32, "x0")
    x = torch.empty_like(x0)
    n = x0.shape[0]
    chw = x0[0].numel()
    if u is not None:
        _need_cuda(u, torch.float32, "u")
    with torch.cuda.device(x0.device):
        _lib.check(_lib.load().b200r_random_start_linf(x0.data_ptr(), x.data_ptr(), n, chw, eps, seed,
                                                       image_offset, _ptr(u), 1 if clip01 else 0, _stream()))
    return x


def random_start_l2(x0: torch.Tensor, eps: float, *, seed: int = 0, image_offset: int = 0) -> torch.Tensor:
    """clip01(x0 + eps * u), u uniform in the unit ball of x0[0].numel() dimensions (foolbox uniform_l2_n_balls), device Philox."""
    _need_cuda(x0, torch.float32, "x0")
    x0 = x0.contiguous()
    x = torch.empty_like(x0)
    with torch.cuda.device(x0.device):
        _lib.check(_lib.load().b200r_random_start_l2(x0.data_ptr(), x.data_ptr(), x0.shape[0], x0[0].numel(), eps, seed, image_offset, _stream()))
    return x


def pgd_step_linf_(x: torch.Tensor, g: torch.Tensor, x0: torch.Tensor, alpha: float, eps: float):
    for t, nm in ((x, "x"), (g, "g"), (x0, "x0")):
        _need_cuda(t, torch.float32, nm)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_pgd_step_linf(x.data_ptr(), g.data_ptr(), x0.data_ptr(), x.shape[0],
                                                   x[0].numel(), alpha, eps, _stream()))
    return x


def pgd_step_l2_(x: torch.Tensor, g: torch.Tensor, x0: torch.Tensor, alpha: float, eps: float):
    for t, nm in ((x, "x"), (g, "g"), (x0, "x0")):
        _need_cuda(t, torch.float32, nm)
    n = x.shape[0]
    ws = workspace(8 * n, x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_pgd_step_l2(x.data_ptr(), g.data_ptr(), x0.data_ptr(), n, x[0].numel(),
                                                 alpha, eps, ws.data_ptr(), _stream()))
    return x


def mim_step_linf_(x, momentum, g, x0, step: float, eps: float, decay: float):
    for t, nm in ((x, "x"), (momentum, "momentum"), (g, "g"), (x0, "x0")):
        _need_cuda(t, torch.float32, nm)
    n = x.shape[0]
    ws = workspace(4 * n, x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_mim_step_linf(x.data_ptr(), momentum.data_ptr(), g.data_ptr(),
                                                   x0.data_ptr(), n, x[0].numel(), step, eps, decay,
                                                   ws.data_ptr(), _stream()))
    return x


def pgd_step_l1_(x, g, x0, eps_step: float, eps: float):
    """One PGD-L1 step of ART's ProjectedGradientDescentPyTorch(norm=1), in place on x [n, ...] (b200r_pgd_step_l1)."""
    for t, nm in ((x, "x"), (g, "g"), (x0, "x0")):
        _need_cuda(t, torch.float32, nm)
    if not (x.shape == g.shape == x0.shape):
        raise ValueError("pgd_step_l1_: x, g, x0 must have the same shape")
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_pgd_step_l1(x.data_ptr(), g.data_ptr(), x0.data_ptr(), x.shape[0], x[0].numel(),
                                                 eps_step, eps, _stream()))
    return x


def l1_projection(x, y, eps: float):
    """L1_projection(x2, y2, eps1) of autopgd_base.py:19-83: delta with ||y + delta||_1 <= eps and 0 <= x + y + delta <= 1."""
    _need_cuda(x, torch.float32, "x")
    _need_cuda(y, torch.float32, "y")
    assert x.shape == y.shape
    delta = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_l1_projection(x.data_ptr(), y.data_ptr(), delta.data_ptr(), x.shape[0], x[0].numel(), eps, _stream()))
    return delta


def fab_projection_linf(t, w, b, want_dmax=False, want_passes=False):
    """projection_linf(points, w, b) of fab_projections.py:7-59 on [rows, dim] float32: the Linf-minimal step d onto the hyperplane
    inside the box.  Optionally also max |d| per row (FAB's a0) and the number of passes each row took."""
    for v, nm in ((t, "t"), (w, "w"), (b, "b")):
        _need_cuda(v, torch.float32, nm)
    rows, dim = t.shape
    assert w.shape == t.shape and b.numel() == rows
    d = torch.empty_like(t)
    dmax = torch.empty(rows, dtype=torch.float32, device=t.device) if want_dmax else None
    passes = torch.empty(rows, dtype=torch.int32, device=t.device) if want_passes else None
    with torch.cuda.device(t.device):
        _lib.check(_lib.load().b200r_fab_projection_linf(t.data_ptr(), w.data_ptr(), b.data_ptr(), d.data_ptr(),
                                                         dmax.data_ptr() if want_dmax else None, rows, dim,
                                                         passes.data_ptr() if want_passes else None, _stream()))
    out = (d,) + ((dmax,) if want_dmax else ()) + ((passes,) if want_passes else ())
    return out if len(out) > 1 else d


def fab_combine_linf_(x1, d1, x0, d2, dmax1, dmax2, eta: float, alpha_max: float):
    """FAB's convex-combination update (fab_base.py:200-232), in place on x1 [rows, ...]."""
    for v, nm in ((x1, "x1"), (d1, "d1"), (x0, "x0"), (d2, "d2"), (dmax1, "dmax1"), (dmax2, "dmax2")):
        _need_cuda(v, torch.float32, nm)
    rows = x1.shape[0]
    if not (d1.numel() == d2.numel() == x0.numel() == x1.numel() and dmax1.numel() == dmax2.numel() == rows):
        raise ValueError("fab_combine_linf_: x1, d1, x0, d2 must hold the same [rows, dim] elements and dmax1 / dmax2 one value per row")
    with torch.cuda.device(x1.device):
        _lib.check(_lib.load().b200r_fab_combine_linf(x1.data_ptr(), d1.data_ptr(), x0.data_ptr(), d2.data_ptr(), dmax1.data_ptr(),
                                                      dmax2.data_ptr(), x1.shape[0], x1[0].numel(), eta, alpha_max, _stream()))
    return x1


def ce_loss_grad(logits: torch.Tensor, labels: torch.Tensor, grad_scale: float = 1.0, want_grad=True):
    _need_cuda(logits, torch.float32, "logits")
    _need_cuda(labels, torch.int64, "labels")
    n, k = logits.shape
    loss = torch.empty(n, dtype=torch.float32, device=logits.device)
    d = torch.empty_like(logits) if want_grad else None
    with torch.cuda.device(logits.device):
        _lib.check(_lib.load().b200r_ce_loss_grad(logits.data_ptr(), labels.data_ptr(), loss.data_ptr(),
                                                  _ptr(d), n, k, grad_scale, _stream()))
    return loss, d


def softmax(logits: torch.Tensor) -> torch.Tensor:
    _need_cuda(logits, torch.float32, "logits")
    out = torch.empty_like(logits)
    with torch.cuda.device(logits.device):
        _lib.check(_lib.load().b200r_softmax(logits.data_ptr(), out.data_ptr(), logits.shape[0],
                                             logits.shape[1], _stream()))
    return out


def topk_count_(counters: torch.Tensor, logits: torch.Tensor, labels: torch.Tensor,
                pred: Optional[torch.Tensor] = None):
    """counters (int64[3], CUDA) += [top1 hits, top5 hits, n]."""
    _need_cuda(logits, torch.float32, "logits")
    _need_cuda(labels, torch.int64, "labels")
    _need_cuda(counters, torch.int64, "counters")
    with torch.cuda.device(logits.device):
        _lib.check(_lib.load().b200r_topk_count(logits.data_ptr(), labels.data_ptr(), logits.shape[0],
                                                logits.shape[1], counters.data_ptr(), _ptr(pred), _stream()))
    return counters


# ------------------------------------------------------------------------------------------------
# split-plane tensors + tensor-core contractions
# ------------------------------------------------------------------------------------------------
ACT = {None: 0, "none": 0, "relu": 1, "relu6": 2, "gelu_tanh": 3, "gelu_erf": 4, "swish": 5, "tanh": 6}


def split_f32(x: torch.Tensor) -> torch.Tensor:
    """float32 tensor of shape S -> int16 tensor [2, *S] holding the (hi, lo) fp16 planes: hi = rn16(x), lo = rn16(x - hi)."""
    _need_cuda(x, torch.float32, "x")
    planes = torch.empty((2,) + tuple(x.shape), dtype=torch.int16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_split_f32(x.data_ptr(), planes.data_ptr(), x.numel(), _stream()))
    return planes


def merge_f32(planes: torch.Tensor) -> torch.Tensor:
    _need_cuda(planes, torch.int16, "planes")
    out = torch.empty(tuple(planes.shape[1:]), dtype=torch.float32, device=planes.device)
    with torch.cuda.device(planes.device):
        _lib.check(_lib.load().b200r_merge_f32(planes.data_ptr(), out.data_ptr(), out.numel(), _stream()))
    return out


PASSES_F16 = 16      # include/b200r.h B200R_PASSES_F16: tensors are ONE plane of fp16 ([1, ...] int16) instead of two fp16 (hi, lo) planes


def to_planes(x: torch.Tensor, f16: bool = False, scale: float = 1.0) -> torch.Tensor:
    """float32 tensor -> the activation / weight format of the chosen precision: split planes [2, *S] (fp16 hi + fp16 lo) or one fp16
    plane [1, *S] (optionally pre-scaled: the loss scaling of the fp16 input-gradient pass)."""
    if not f16:
        return split_f32(x if scale == 1.0 else x * scale)      # pre-scaling: a [n, classes] tensor (loss scale of the gradient pass)
    _need_cuda(x, torch.float32, "x")
    out = torch.empty((1,) + tuple(x.shape), dtype=torch.int16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_f32_to_f16(x.data_ptr(), out.data_ptr(), x.numel(), scale, _stream()))
    return out


def from_planes(planes: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    if planes.shape[0] == 2:
        assert scale == 1.0
        return merge_f32(planes)
    _need_cuda(planes, torch.int16, "planes")
    out = torch.empty(tuple(planes.shape[1:]), dtype=torch.float32, device=planes.device)
    with torch.cuda.device(planes.device):
        _lib.check(_lib.load().b200r_f16_to_f32(planes.data_ptr(), out.data_ptr(), out.numel(), scale, _stream()))
    return out


def _passes_for(x, passes):
    """A one-plane tensor is fp16 by construction: the contraction must run in B200R_PASSES_F16 mode."""
    return PASSES_F16 if x.shape[0] == 1 else passes


def conv2d_nhwc(x, wgt, scale=None, bias=None, res=None, *, stride=1, pad=0, act=None, passes=3, out=None,
                out_f32=None, want_planes=True):
    """x: planes [P,n,h,w,cin]; wgt: planes [P,cout,kh,kw,cin]; returns planes [P,n,ho,wo,cout] (P = 2 split hi/lo, 1 fp16)."""
    _need_cuda(x, torch.int16, "x")
    _need_cuda(wgt, torch.int16, "wgt")
    P, n, h, w, cin = x.shape
    _, cout, kh, kw, cin2 = wgt.shape
    assert cin == cin2 and wgt.shape[0] == P and (res is None or res.shape[0] == P)
    passes = _passes_for(x, passes)
    ho, wo = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
    if out is None and want_planes:
        out = torch.empty((P, n, ho, wo, cout), dtype=torch.int16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_conv2d_nhwc(x.data_ptr(), wgt.data_ptr(), _ptr(scale), _ptr(bias), _ptr(res),
                                                 _ptr(out), _ptr(out_f32), n, h, w, cin, cout, kh, kw, stride, pad,
                                                 ACT[act], passes, _stream()))
    return out if out is not None else out_f32


def linear(x, wgt, scale=None, bias=None, res=None, *, act=None, passes=3, out=None, out_f32=None, want_planes=True):
    """x: planes [P,m,k]; wgt: planes [P,nout,k]."""
    _need_cuda(x, torch.int16, "x")
    _need_cuda(wgt, torch.int16, "wgt")
    m, k = x.shape[1], x.shape[-1]
    m = x[0].numel() // k
    nout = wgt.shape[1]
    assert wgt.shape[0] == x.shape[0] and (res is None or res.shape[0] == x.shape[0])
    passes = _passes_for(x, passes)
    if out is None and want_planes:
        out = torch.empty((x.shape[0], m, nout), dtype=torch.int16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_linear(x.data_ptr(), wgt.data_ptr(), _ptr(scale), _ptr(bias), _ptr(res), _ptr(out),
                                            _ptr(out_f32), m, k, nout, ACT[act], passes, _stream()))
    return out if out is not None else out_f32


def linear_keep_pre(x, wgt, bias=None, *, act):
    """One launch for the Linear a gradient pass keeps: returns (act(pre), pre), both split planes [2, m, nout] and bit-identical to
    `linear(act=None)` followed by `act_planes`.  Split precision only."""
    _need_cuda(x, torch.int16, "x")
    _need_cuda(wgt, torch.int16, "wgt")
    if x.shape[0] != 2 or wgt.shape[0] != 2:
        raise ValueError("linear_keep_pre: split planes ([2, ...]) only")
    k = x.shape[-1]
    m = x[0].numel() // k
    nout = wgt.shape[1]
    y = torch.empty((2, m, nout), dtype=torch.int16, device=x.device)
    pre = torch.empty_like(y)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_linear_keep_pre(x.data_ptr(), wgt.data_ptr(), _ptr(bias), y.data_ptr(), pre.data_ptr(),
                                                     m, k, nout, ACT[act], _stream()))
    return y, pre


def stem_im2col(img, mean=IMAGENET_MEAN, std=IMAGENET_STD, out=None):
    """uint8 NHWC or float32 NCHW image batch -> planes [2, n*ho*wo, 192] of normalised 7x7/s2 patches."""
    if img.dtype == torch.uint8:
        n, h, w, _ = img.shape
        fn = _lib.load().b200r_stem_im2col_u8
    else:
        _need_cuda(img, torch.float32, "img")
        n, _, h, w = img.shape
        fn = _lib.load().b200r_stem_im2col_f32
    rows = n * (h // 2) * (w // 2)
    if out is None:
        out = torch.empty((2, rows, 192), dtype=torch.int16, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(fn(img.data_ptr(), out.data_ptr(), n, h, w, _lib.f3(mean), _lib.f3(std), _stream()))
    return out


def pack_stem_weight(w: torch.Tensor) -> torch.Tensor:
    """conv1 weight [64, 3, 7, 7] -> float32 [64, 192] in the stem's K order: column = ky*24 + kx*3 + c, every ky
    run padded from 7 to 8 taps with zeros (include/b200r.h, b200r_stem_im2col_u8)."""
    cout = w.shape[0]
    wp = torch.zeros(cout, 7, 8, 3, dtype=torch.float32, device=w.device)
    wp[:, :, :7, :] = w.float().permute(0, 2, 3, 1)
    out = torch.zeros(cout, 192, dtype=torch.float32, device=w.device)
    out[:, :168] = wp.reshape(cout, 168)
    return out


def stem_conv7x7_u8(img, wgt, scale, bias, *, act="relu", passes=3, mean=IMAGENET_MEAN, std=IMAGENET_STD, out=None):
    """Fused 7x7/s2 stem from raw uint8 NHWC pixels: planes [P, n, h/2, w/2, 64]."""
    _need_cuda(img, torch.uint8, "img")
    n, h, w, _ = img.shape
    passes = _passes_for(wgt, passes)
    if out is None:
        out = torch.empty((wgt.shape[0], n, h // 2, w // 2, 64), dtype=torch.int16, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(_lib.load().b200r_stem_conv7x7_u8(img.data_ptr(), wgt.data_ptr(), _ptr(scale), _ptr(bias), out.data_ptr(),
                                                     n, h, w, _lib.f3(mean), _lib.f3(std), ACT[act], passes, _stream()))
    return out


def stem_pool_ok(h: int, w: int) -> bool:
    """Geometry accepted by the one-launch stem (csrc/stem_pool_sm100.cu)."""
    return h % 4 == 0 and w % 8 == 0 and 8 <= w <= 248 and h >= 8


def stem_pool_u8(img, wgt_folded, bias, *, mean=IMAGENET_MEAN, std=IMAGENET_STD, out=None):
    """conv1 7x7/s2 + BN + ReLU + MaxPool(3,2,1) from raw uint8 NHWC pixels in one launch (fp16 mode): plane [1, n, h/4, w/4, 64].
    wgt_folded: fp16 plane [1, 64, 192] = to_planes(pack_stem_weight(w * bn_scale[:, None, None, None]), True)."""
    _need_cuda(img, torch.uint8, "img")
    n, h, w, _ = img.shape
    assert wgt_folded.shape[0] == 1 and tuple(wgt_folded.shape[1:]) == (64, 192), "stem_pool_u8 takes the fp16 [1, 64, 192] stem weight"
    assert stem_pool_ok(h, w), "stem_pool_u8: unsupported geometry %dx%d" % (h, w)
    if out is None:
        out = torch.empty((1, n, h // 4, w // 4, 64), dtype=torch.int16, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(_lib.load().b200r_stem_pool_u8_f16(img.data_ptr(), wgt_folded.data_ptr(), _ptr(bias), out.data_ptr(),
                                                      n, h, w, _lib.f3(mean), _lib.f3(std), _stream()))
    return out


def stem_pool_split_prepare(conv1_w, bn_scale, bn_bias, device, *, mean=IMAGENET_MEAN, std=IMAGENET_STD, f32_input=False):
    """Operand of the split-precision one-launch stem from the float32 conv1 weight [64, 3, 7, 7] and the folded BN scale / bias
    [64] (host arithmetic in double, include/b200r.h): (device planes [2, 64, 224] int16, out_scale)."""
    import ctypes as C
    w = conv1_w.detach().float().cpu().contiguous()
    assert tuple(w.shape) == (64, 3, 7, 7), "the ResNet stem is 3 -> 64, 7x7"
    s = None if bn_scale is None else bn_scale.detach().float().cpu().contiguous()
    b = None if bn_bias is None else bn_bias.detach().float().cpu().contiguous()
    planes = torch.empty((2, 64, 224), dtype=torch.int16)
    osc = C.c_float(0.0)
    _lib.check(_lib.load().b200r_stem_pool_split_prepare(w.data_ptr(), _ptr(s), _ptr(b), _lib.f3(mean), _lib.f3(std), int(f32_input), planes.data_ptr(),
                                                         C.byref(osc)))
    return planes.to(device), float(osc.value)


def stem_pool_u8_split(img, wplanes, out_scale, *, out=None):
    """conv1 7x7/s2 + BN + ReLU + MaxPool(3,2,1) from raw uint8 NHWC pixels in one launch, split precision: planes
    [2, n, h/4, w/4, 64].  wplanes / out_scale from stem_pool_split_prepare."""
    _need_cuda(img, torch.uint8, "img")
    _need_cuda(wplanes, torch.int16, "wplanes")
    n, h, w, _ = img.shape
    assert tuple(wplanes.shape) == (2, 64, 224), "stem_pool_u8_split takes the prepared [2, 64, 224] operand"
    assert stem_pool_ok(h, w), "stem_pool_u8_split: unsupported geometry %dx%d" % (h, w)
    if out is None:
        out = torch.empty((2, n, h // 4, w // 4, 64), dtype=torch.int16, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(_lib.load().b200r_stem_pool_u8_split(img.data_ptr(), wplanes.data_ptr(), out_scale, out.data_ptr(), n, h, w, _stream()))
    return out


def stem_pool_f32_split(x01, wplanes, out_scale):
    """conv1 7x7/s2 + BN + ReLU + MaxPool(3,2,1) from a float32 NCHW image in [0,1] in one launch (split precision): (planes
    [2, n, h/4, w/4, 64], arg-max codes uint8 [n, h/4, w/4, 64] for maxpool3x3s2_bwd_codes_hi).  wplanes / out_scale from
    stem_pool_split_prepare(..., f32_input=True)."""
    _need_cuda(x01, torch.float32, "x01")
    _need_cuda(wplanes, torch.int16, "wplanes")
    n, c, h, w = x01.shape
    assert c == 3 and x01.is_contiguous() and tuple(wplanes.shape) == (2, 64, 224)
    assert stem_pool_ok(h, w), "stem_pool_f32_split: unsupported geometry %dx%d" % (h, w)
    out = torch.empty((2, n, h // 4, w // 4, 64), dtype=torch.int16, device=x01.device)
    codes = torch.empty((n, h // 4, w // 4, 64), dtype=torch.uint8, device=x01.device)
    with torch.cuda.device(x01.device):
        _lib.check(_lib.load().b200r_stem_pool_f32_split(x01.data_ptr(), wplanes.data_ptr(), out_scale, out.data_ptr(), codes.data_ptr(), n, h, w, _stream()))
    return out, codes


def stem_conv7x7_f32(img, wgt, scale, bias, *, act="relu", passes=3, mean=IMAGENET_MEAN, std=IMAGENET_STD, out=None):
    """Fused 7x7/s2 stem from a float32 NCHW image in [0,1] (attack iterates): planes [P, n, h/2, w/2, 64]."""
    _need_cuda(img, torch.float32, "img")
    n, _, h, w = img.shape
    passes = _passes_for(wgt, passes)
    if out is None:
        out = torch.empty((wgt.shape[0], n, h // 2, w // 2, 64), dtype=torch.int16, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(_lib.load().b200r_stem_conv7x7_f32(img.data_ptr(), wgt.data_ptr(), _ptr(scale), _ptr(bias), out.data_ptr(),
                                                      n, h, w, _lib.f3(mean), _lib.f3(std), ACT[act], passes, _stream()))
    return out


RESIZE_FILTERS = {"nearest": 0, "box": 1, "bilinear": 2, "hamming": 3, "bicubic": 4, "cubic": 4, "lanczos": 5}


def resize_u8(images, size, filter="bilinear", crop=None, out=None):
    """PIL.Image.resize((size[1], size[0]), FILTER) of every image of a uint8 NHWC CUDA batch, bit-exact (csrc/resize.cu);
    crop = (y0, x0, h, w) returns that window of the resized image (only it is computed)."""
    import ctypes as C
    _need_cuda(images, torch.uint8, "images")
    n, hin, win, c = images.shape
    assert c == 3 and images.is_contiguous()
    hout, wout = int(size[0]), int(size[1])
    y0, x0, ch, cw = crop if crop is not None else (0, 0, hout, wout)
    fid = RESIZE_FILTERS[filter]
    if out is None:
        out = torch.empty((n, ch, cw, 3), dtype=torch.uint8, device=images.device)
    lib = _lib.load()
    with torch.cuda.device(images.device):
        need = C.c_size_t()
        _lib.check(lib.b200r_resize_workspace_bytes(n, hin, win, hout, wout, fid, y0, x0, ch, cw, C.byref(need)))
        ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=images.device)
        _lib.check(lib.b200r_resize_u8(images.data_ptr(), out.data_ptr(), n, hin, win, hout, wout, fid, y0, x0, ch, cw,
                                       ws.data_ptr(), need.value, _stream()))
    return out


CV_INTERPOLATIONS = {"nearest": 0, "bilinear": 1, "cubic": 2, "area": 3, "lanczos": 4}      # cv2.INTER_* constants


def resize_cv_u8(images, size, interpolation="bilinear", crop=None, out=None):
    """cv2.resize(img, (size[1], size[0]), interpolation=INTER_NEAREST | INTER_LINEAR | INTER_AREA) of every image of a uint8 NHWC CUDA batch,
    bit-exact (csrc/resize_cv.cu); crop = (y0, x0, h, w) returns that window of the resized image (only it is computed)."""
    _need_cuda(images, torch.uint8, "images")
    n, hin, win, c = images.shape
    assert c == 3
    hout, wout = int(size[0]), int(size[1])
    y0, x0, ch, cw = crop if crop is not None else (0, 0, hout, wout)
    if out is None:
        out = torch.empty((n, ch, cw, 3), dtype=torch.uint8, device=images.device)
    with torch.cuda.device(images.device):
        _lib.check(_lib.load().b200r_resize_cv_u8(images.data_ptr(), out.data_ptr(), n, hin, win, hout, wout, CV_INTERPOLATIONS[interpolation],
                                                  y0, x0, ch, cw, _stream()))
    return out


def resize_center_crop_u8(images, resize=256, crop=224, filter="bilinear"):
    """torchvision Resize(resize) (short side, aspect kept: long side = int(resize * long / short)) + CenterCrop(crop) on a uint8
    NHWC batch -- the eval transform of imagenet_dataloader.py:74-80, antialiased PIL bilinear as torchvision does on PIL images."""
    n, h, w, _ = images.shape
    if h <= w:
        oh, ow = resize, int(resize * w / h)
    else:
        oh, ow = int(resize * h / w), resize
    y0, x0 = int(round((oh - crop) / 2.0)), int(round((ow - crop) / 2.0))
    return resize_u8(images, (oh, ow), filter, crop=(y0, x0, crop, crop))


def maxpool3x3s2(x, out=None):
    P, n, h, w, c = x.shape
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    if out is None:
        out = torch.empty((P, n, ho, wo, c), dtype=torch.int16, device=x.device)
    fn = _lib.load().b200r_maxpool3x3s2_nhwc_f16 if P == 1 else _lib.load().b200r_maxpool3x3s2_nhwc
    with torch.cuda.device(x.device):
        _lib.check(fn(x.data_ptr(), out.data_ptr(), n, h, w, c, _stream()))
    return out


def global_avgpool(x, out=None):
    P, n, h, w, c = x.shape
    if out is None:
        out = torch.empty((P, n, c), dtype=torch.int16, device=x.device)
    fn = _lib.load().b200r_global_avgpool_nhwc_f16 if P == 1 else _lib.load().b200r_global_avgpool_nhwc
    with torch.cuda.device(x.device):
        _lib.check(fn(x.data_ptr(), out.data_ptr(), n, h * w, c, _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# input-gradient pass (split planes in, split planes out)
# ------------------------------------------------------------------------------------------------
def conv2d_dgrad(dy, wgt_t, res=None, mask=None, *, pad=0, passes=3):
    """dx = [mask > 0] * (conv_s1(dy, wgt_t) + res); dy planes [2,n,h,w,cdy], wgt_t planes [2,cdx,kh,kw,cdy], mask = the
    post-ReLU activation planes the gradient flows into (only its hi plane is read)."""
    P, n, h, w, cdy = dy.shape
    cdx, kh, kw = wgt_t.shape[1], wgt_t.shape[2], wgt_t.shape[3]
    assert wgt_t.shape[0] == P
    passes = _passes_for(dy, passes)
    out = torch.empty((P, n, h, w, cdx), dtype=torch.int16, device=dy.device)
    if mask is not None:
        assert mask.shape == out.shape, (mask.shape, out.shape)
    if res is not None:
        assert res.shape == out.shape, (res.shape, out.shape)
    with torch.cuda.device(dy.device):
        _lib.check(_lib.load().b200r_conv2d_dgrad_nhwc(dy.data_ptr(), wgt_t.data_ptr(), _ptr(res), _ptr(mask), out.data_ptr(), n, h, w,
                                                       cdy, cdx, kh, kw, pad, passes, _stream()))
    return out


def conv2d_dgrad3x3s2(dy, wsub, res=None, mask=None, *, passes=3):
    """Input gradient of a 3x3 / stride 2 / pad 1 convolution by parity classes (include/b200r.h): dy planes [P,n,ho,wo,cdy],
    wsub = the four sub-kernels [P,cdx,1+a,1+b,cdy] in the order (0,0), (0,1), (1,0), (1,1) -> planes [P,n,2ho,2wo,cdx]."""
    _need_cuda(dy, torch.int16, "dy")
    P, n, ho, wo, cdy = dy.shape
    cdx = wsub[0].shape[1]
    for i, ws in enumerate(wsub):
        _need_cuda(ws, torch.int16, "wsub")
        if tuple(ws.shape) != (P, cdx, 1 + i // 2, 1 + i % 2, cdy) or not ws.is_contiguous():
            raise ValueError("conv2d_dgrad3x3s2: sub-kernel %d must be contiguous planes [%d, %d, %d, %d, %d]" % (i, P, cdx, 1 + i // 2, 1 + i % 2, cdy))
    passes = _passes_for(dy, passes)
    out = torch.empty((P, n, 2 * ho, 2 * wo, cdx), dtype=torch.int16, device=dy.device)
    for t, nm in ((res, "res"), (mask, "mask")):
        if t is not None and t.shape != out.shape:
            raise ValueError("conv2d_dgrad3x3s2: %s must have the output's shape %s" % (nm, tuple(out.shape)))
    with torch.cuda.device(dy.device):
        _lib.check(_lib.load().b200r_conv2d_dgrad3x3s2_nhwc(dy.data_ptr(), wsub[0].data_ptr(), wsub[1].data_ptr(), wsub[2].data_ptr(), wsub[3].data_ptr(),
                                                            _ptr(res), _ptr(mask), out.data_ptr(), n, ho, wo, cdy, cdx, passes, _stream()))
    return out


def conv2d_dgrad1x1s2_acc(dy, wgt_t, dx, mask=None, *, passes=3):
    """In place: dx[:, :, 2i, 2j] = [mask > 0] * (dx[:, :, 2i, 2j] + dy[:, :, i, j] . wgt_t^T) -- the input gradient of a 1x1 / stride-2
    convolution (the downsample branch) accumulated into the main branch's gradient (include/b200r.h).  dy planes [P,n,ho,wo,cdy],
    wgt_t planes [P,cdx,1,1,cdy], dx (and mask) planes [P,n,h,w,cdx] with ho = (h + 1) // 2.  Returns dx."""
    _need_cuda(dy, torch.int16, "dy")
    _need_cuda(wgt_t, torch.int16, "wgt_t")
    _need_cuda(dx, torch.int16, "dx")
    P, n, ho, wo, cdy = dy.shape
    _, n2, h, w, cdx = dx.shape
    if dx.shape[0] != P or n2 != n or (h + 1) // 2 != ho or (w + 1) // 2 != wo or not dx.is_contiguous():
        raise ValueError("conv2d_dgrad1x1s2_acc: dx %s does not match dy %s" % (tuple(dx.shape), tuple(dy.shape)))
    if wgt_t.shape[0] != P or wgt_t[0].numel() != cdx * cdy or not wgt_t.is_contiguous():
        raise ValueError("conv2d_dgrad1x1s2_acc: wgt_t must be contiguous planes [%d, %d, 1, 1, %d]" % (P, cdx, cdy))
    if mask is not None and (mask.shape != dx.shape or not mask.is_contiguous()):
        raise ValueError("conv2d_dgrad1x1s2_acc: mask must have dx's shape")
    passes = _passes_for(dy, passes)
    with torch.cuda.device(dy.device):
        _lib.check(_lib.load().b200r_conv2d_dgrad1x1s2_acc_nhwc(dy.data_ptr(), wgt_t.data_ptr(), _ptr(mask), dx.data_ptr(), n, h, w, cdy, cdx,
                                                                passes, _stream()))
    return dx


def relu_bwd(dy, act, add=None, out=None):
    """(act > 0 ? dy : 0) + add on split planes of any shape [2, ...]."""
    assert dy.shape == act.shape and (add is None or add.shape == dy.shape)
    count = dy[0].numel()
    if out is None:
        out = torch.empty_like(dy)
    fn = _lib.load().b200r_relu_bwd_f16 if dy.shape[0] == 1 else _lib.load().b200r_relu_bwd
    with torch.cuda.device(dy.device):
        _lib.check(fn(dy.data_ptr(), act.data_ptr(), _ptr(add), out.data_ptr(), count, _stream()))
    return out


def dilate2(x):
    P, n, h, w, c = x.shape
    out = torch.empty((P, n, 2 * h, 2 * w, c), dtype=torch.int16, device=x.device)
    fn = _lib.load().b200r_dilate2_nhwc_f16 if P == 1 else _lib.load().b200r_dilate2_nhwc
    with torch.cuda.device(x.device):
        _lib.check(fn(x.data_ptr(), out.data_ptr(), n, h, w, c, _stream()))
    return out


def maxpool3x3s2_bwd(x, dy):
    """x: the pool's forward input planes [2,n,h,w,c]; dy planes [2,n,ho,wo,c] -> dx planes like x."""
    _, n, h, w, c = x.shape
    ws = torch.empty(dy[0].numel(), dtype=torch.uint8, device=x.device)
    dx = torch.empty_like(x)
    fn = _lib.load().b200r_maxpool3x3s2_bwd_nhwc_f16 if x.shape[0] == 1 else _lib.load().b200r_maxpool3x3s2_bwd_nhwc
    with torch.cuda.device(x.device):
        _lib.check(fn(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), ws.data_ptr(), ws.numel(), n, h, w, c, _stream()))
    return dx


def maxpool3x3s2_codes(x):
    """MaxPool(3,2,1) of post-ReLU split planes that also returns the arg-max codes (uint8 [n,ho,wo,c]; 0xF = nothing to route: the
    ReLU's backward folded in) for maxpool3x3s2_bwd_codes_hi."""
    _need_cuda(x, torch.int16, "x")
    P, n, h, w, c = x.shape
    if P != 2:
        raise ValueError("maxpool3x3s2_codes takes split planes")
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    out = torch.empty((2, n, ho, wo, c), dtype=torch.int16, device=x.device)
    codes = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_maxpool3x3s2_nhwc_codes(x.data_ptr(), out.data_ptr(), codes.data_ptr(), n, h, w, c, _stream()))
    return out, codes


def maxpool3x3s2_bwd_codes_hi(codes, dy, h, w):
    """Routes dy planes [P,n,ho,wo,c] back through the codes of maxpool3x3s2_codes: ONE fp16 plane [1,n,h,w,c]."""
    _need_cuda(dy, torch.int16, "dy")
    P, n, ho, wo, c = dy.shape
    if codes.dtype != torch.uint8 or tuple(codes.shape) != (n, ho, wo, c) or ho != (h - 1) // 2 + 1 or wo != (w - 1) // 2 + 1:
        raise ValueError("maxpool3x3s2_bwd_codes_hi: codes must be the uint8 [n, ho, wo, c] tensor of the forward pool of an h x w input")
    dx = torch.empty((1, n, h, w, c), dtype=torch.int16, device=dy.device)
    with torch.cuda.device(dy.device):
        _lib.check(_lib.load().b200r_maxpool3x3s2_bwd_codes_hi(codes.data_ptr(), dy.data_ptr(), dx.data_ptr(), n, h, w, c, P, _stream()))
    return dx


def maxpool3x3s2_relu_bwd_hi(x, dy):
    """MaxPool(3,2,1) backward with the backward of the ReLU that produced x fused in; ONE fp16 plane [1,n,h,w,c] out (the stem's
    gradient GEMM reads a single plane).  x: the pool's forward input planes [P,n,h,w,c]; dy planes [P,n,ho,wo,c]."""
    _need_cuda(x, torch.int16, "x")
    _need_cuda(dy, torch.int16, "dy")
    P, n, h, w, c = x.shape
    if dy.shape[0] != P or tuple(dy.shape[1:]) != (n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c):
        raise ValueError("maxpool3x3s2_relu_bwd_hi: dy must be the pooled shape of x in the same precision")
    ws = torch.empty(dy[0].numel(), dtype=torch.uint8, device=x.device)
    dx = torch.empty((1, n, h, w, c), dtype=torch.int16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_maxpool3x3s2_relu_bwd_hi(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), ws.data_ptr(), ws.numel(), n, h, w, c, P, _stream()))
    return dx


def global_avgpool_bwd(dy, h, w):
    P, n, c = dy.shape
    dx = torch.empty((P, n, h, w, c), dtype=torch.int16, device=dy.device)
    fn = _lib.load().b200r_global_avgpool_bwd_nhwc_f16 if P == 1 else _lib.load().b200r_global_avgpool_bwd_nhwc
    with torch.cuda.device(dy.device):
        _lib.check(fn(dy.data_ptr(), dx.data_ptr(), n, h * w, c, _stream()))
    return dx


def stem_col2im(dcols, n, h, w, std=IMAGENET_STD, out=None, unscale: float = 1.0):
    """dcols planes [P, n*(h/2)*(w/2), 192] -> float32 NCHW gradient w.r.t. the [0,1] image (fp16 planes: times
    `unscale`, the inverse of the loss scale the pass ran with)."""
    if out is None:
        out = torch.empty((n, 3, h, w), dtype=torch.float32, device=dcols.device)
    with torch.cuda.device(dcols.device):
        if dcols.shape[0] == 1:
            _lib.check(_lib.load().b200r_stem_col2im_f32_f16(dcols.data_ptr(), out.data_ptr(), n, h, w, _lib.f3(std), unscale, _stream()))
        else:
            # split planes: the kernel divides by std, so the inverse loss scale rides on it
            std_u = tuple(float(s) / unscale for s in std)
            _lib.check(_lib.load().b200r_stem_col2im_f32(dcols.data_ptr(), out.data_ptr(), n, h, w, _lib.f3(std_u), _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# token-model layers (ViT / MLP-Mixer)
# ------------------------------------------------------------------------------------------------
def layernorm(x, gamma, beta, eps=1e-5, out=None):
    """x: planes [2, rows, c]."""
    _need_cuda(x, torch.int16, "x")
    c = x.shape[-1]
    rows = x[0].numel() // c
    if out is None:
        out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_layernorm(x.data_ptr(), out.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rows, c, eps, _stream()))
    return out


def patch_gather(img, patch=16, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """uint8 NHWC or float32 NCHW batch -> planes [2, n*(h/p)*(w/p), 3*p*p]."""
    if img.dtype == torch.uint8:
        n, h, w, _ = img.shape
        fn = _lib.load().b200r_patch_gather_u8
    else:
        _need_cuda(img, torch.float32, "img")
        n, _, h, w = img.shape
        fn = _lib.load().b200r_patch_gather_f32
    out = torch.empty((2, n * (h // patch) * (w // patch), 3 * patch * patch), dtype=torch.int16, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(fn(img.data_ptr(), out.data_ptr(), n, h, w, patch, _lib.f3(mean), _lib.f3(std), _stream()))
    return out


def assemble_tokens(x, cls, pos, n, num_patches):
    c = x.shape[-1]
    out = torch.empty((2, n * (num_patches + 1), c), dtype=torch.int16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_assemble_tokens(x.data_ptr(), cls.data_ptr(), pos.data_ptr(), out.data_ptr(), n, num_patches, c, _stream()))
    return out


def attention(qkv, n, tokens, heads, head_dim, scale):
    out = torch.empty((2, n * tokens, heads * head_dim), dtype=torch.int16, device=qkv.device)
    with torch.cuda.device(qkv.device):
        _lib.check(_lib.load().b200r_attention(qkv.data_ptr(), out.data_ptr(), n, tokens, heads, head_dim, scale, _stream()))
    return out


def tokens_to_channels(x, b, t, c, t_pad):
    out = torch.empty((2, b * c, t_pad), dtype=torch.int16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_tokens_to_channels(x.data_ptr(), out.data_ptr(), b, t, c, t_pad, _stream()))
    return out


def channels_to_tokens_add(y, res, b, t, c, t_pad):
    out = torch.empty((2, b * t, c), dtype=torch.int16, device=y.device)
    with torch.cuda.device(y.device):
        _lib.check(_lib.load().b200r_channels_to_tokens_add(y.data_ptr(), res.data_ptr(), out.data_ptr(), b, t, c, t_pad, _stream()))
    return out


# input-gradient pass of the token models (token_backward.cu) ---------------------------------------
def layernorm_bwd(dy, x, gamma, eps=1e-5, add=None):
    """dy, x (the LayerNorm's INPUT), add (optional residual gradient): planes [2, rows, c] -> dx planes."""
    _need_cuda(dy, torch.int16, "dy")
    _need_cuda(x, torch.int16, "x")
    c = x.shape[-1]
    rows = x[0].numel() // c
    assert dy.shape == x.shape and (add is None or add.shape == x.shape)
    dx = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_layernorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), _ptr(add), dx.data_ptr(), rows, c, eps,
                                                   _stream()))
    return dx


def act_planes(pre, act):
    """act(pre) on split planes (the forward of a Linear whose pre-activation is kept for the gradient pass)."""
    _need_cuda(pre, torch.int16, "pre")
    out = torch.empty_like(pre)
    with torch.cuda.device(pre.device):
        _lib.check(_lib.load().b200r_act_planes(pre.data_ptr(), out.data_ptr(), pre[0].numel(), ACT[act], _stream()))
    return out


def act_bwd_planes(dy, pre, act):
    _need_cuda(dy, torch.int16, "dy")
    _need_cuda(pre, torch.int16, "pre")
    assert dy.shape == pre.shape
    dx = torch.empty_like(pre)
    with torch.cuda.device(pre.device):
        _lib.check(_lib.load().b200r_act_bwd_planes(dy.data_ptr(), pre.data_ptr(), dx.data_ptr(), pre[0].numel(), ACT[act], _stream()))
    return dx


def patch_scatter(dcols, n, h, w, patch=16, std=IMAGENET_STD, unscale: float = 1.0):
    """dcols planes [2, n*(h/p)*(w/p), 3*p*p] -> float32 NCHW gradient w.r.t. the [0,1] image (transpose of patch_gather),
    times `unscale` (the inverse loss scale of the pass; it rides on the 1/std the kernel applies)."""
    _need_cuda(dcols, torch.int16, "dcols")
    std = tuple(float(s) / unscale for s in std)
    assert dcols[0].numel() == n * 3 * h * w
    dx = torch.empty((n, 3, h, w), dtype=torch.float32, device=dcols.device)
    with torch.cuda.device(dcols.device):
        _lib.check(_lib.load().b200r_patch_scatter_f32(dcols.data_ptr(), dx.data_ptr(), n, h, w, patch, _lib.f3(std), _stream()))
    return dx


def attention_bwd(qkv, dout, n, tokens, heads, head_dim, scale):
    _need_cuda(qkv, torch.int16, "qkv")
    _need_cuda(dout, torch.int16, "dout")
    dqkv = torch.empty_like(qkv)
    ws = torch.empty(n * heads * 3 * tokens, dtype=torch.float32, device=qkv.device)       # per-query softmax statistics (tensor-core path)
    with torch.cuda.device(qkv.device):
        _lib.check(_lib.load().b200r_attention_bwd_ws(qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), ws.data_ptr(), ws.numel() * 4,
                                                      n, tokens, heads, head_dim, scale, _stream()))
    return dqkv


# ------------------------------------------------------------------------------------------------
# mobile-family layers
# ------------------------------------------------------------------------------------------------
ACT["sigmoid"] = 7


def dwconv_nhwc(x, wgt, scale, bias, *, k, stride, pad, act=None):
    """x: planes [2,n,h,w,c]; wgt float32 [k*k, c]."""
    _, n, h, w, c = x.shape
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    out = torch.empty((2, n, ho, wo, c), dtype=torch.int16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_dwconv_nhwc(x.data_ptr(), wgt.data_ptr(), scale.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                                 n, h, w, c, k, stride, pad, ACT[act], _stream()))
    return out


def pointwise_smallk(x, wgt, bias=None, res=None, *, act=None):
    """1x1 convolution with cin in {8, 16, 24, 32} on CUDA cores in exact fp32 (include/b200r.h): x planes [2, ..., cin],
    wgt float32 [cout, cin] (BN scale folded), bias float32 [cout], res planes [2, ..., cout] -> planes [2, ..., cout]."""
    _need_cuda(x, torch.int16, "x")
    _need_cuda(wgt, torch.float32, "wgt")
    cin, cout = x.shape[-1], wgt.shape[0]
    if x.shape[0] != 2 or tuple(wgt.shape) != (cout, cin) or not wgt.is_contiguous():
        raise ValueError("pointwise_smallk: x must be split planes [2, ..., cin] and wgt contiguous float32 [cout, cin]")
    m = x[0].numel() // cin
    out = torch.empty((2,) + tuple(x.shape[1:-1]) + (cout,), dtype=torch.int16, device=x.device)
    if res is not None and res.shape != out.shape:
        raise ValueError("pointwise_smallk: res must have the output's shape")
    if bias is not None:
        _need_cuda(bias, torch.float32, "bias")
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_pointwise_smallk_nhwc(x.data_ptr(), wgt.data_ptr(), _ptr(bias), _ptr(res), out.data_ptr(), m, cin, cout, ACT[act], _stream()))
    return out


def channel_scale(x, s):
    """x: planes [2,n,h,w,c]; s: planes [2,n,c_stride] -> x * s[n, :c]."""
    _, n, h, w, c = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200r_channel_scale(x.data_ptr(), s.data_ptr(), out.data_ptr(), n, h * w, c, s.shape[-1], _stream()))
    return out


def image_stem3x3s2(img, wgt, scale, bias, *, act="relu6", mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """3x3/s2/p1 conv from the image + folded BN + activation in one launch (CUDA cores, fp32): planes [2, n, ho, wo, cout].
    img: uint8 NHWC or float32 NCHW in [0,1]; wgt: float32 [cout, 27], column = (ky*3 + kx)*3 + c."""
    if img.dtype == torch.uint8:
        n, h, w, _ = img.shape
        fn = _lib.load().b200r_image_stem3x3s2_u8
    else:
        _need_cuda(img, torch.float32, "img")
        n, _, h, w = img.shape
        fn = _lib.load().b200r_image_stem3x3s2_f32
    cout = wgt.shape[0]
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    out = torch.empty((2, n, ho, wo, cout), dtype=torch.int16, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(fn(img.data_ptr(), wgt.data_ptr(), _ptr(scale), _ptr(bias), out.data_ptr(), n, h, w, cout, ACT[act],
                      _lib.f3(mean), _lib.f3(std), _stream()))
    return out


def image_im2col(img, k, stride, pad, kpad, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    if img.dtype == torch.uint8:
        n, h, w, _ = img.shape
        fn = _lib.load().b200r_image_im2col_u8
    else:
        _need_cuda(img, torch.float32, "img")
        n, _, h, w = img.shape
        fn = _lib.load().b200r_image_im2col_f32
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    out = torch.empty((2, n * ho * wo, kpad), dtype=torch.int16, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(fn(img.data_ptr(), out.data_ptr(), n, h, w, k, stride, pad, kpad, _lib.f3(mean), _lib.f3(std), _stream()))
    return out, ho, wo


# ------------------------------------------------------------------------------------------------
# input-gradient pieces of the mobile families
# ------------------------------------------------------------------------------------------------
def channel_dot(a, b, s_stride=None):
    """planes [2, n, s_stride]: sum over pixels of a * b (a, b planes [2,n,h,w,c]) -- d loss / d (squeeze-excite scale)."""
    _, n, h, w, c = a.shape
    assert a.shape == b.shape
    s_stride = c if s_stride is None else s_stride
    out = torch.empty((2, n, s_stride), dtype=torch.int16, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().b200r_channel_dot(a.data_ptr(), b.data_ptr(), out.data_ptr(), n, h * w, c, s_stride, _stream()))
    return out


def planes_add(a, b):
    assert a.shape == b.shape
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().b200r_planes_add(a.data_ptr(), b.data_ptr(), out.data_ptr(), a[0].numel(), _stream()))
    return out


def image_stem3x3s2_bwd(dy, wgt_scaled, n, h, w, std=IMAGENET_STD, unscale: float = 1.0):
    """dy planes [2, n, ho, wo, cout] -> float32 NCHW gradient w.r.t. the [0,1] image (transposed 3x3/s2 stem, 1/std, unscale)."""
    cout = wgt_scaled.shape[0]
    dx = torch.empty((n, 3, h, w), dtype=torch.float32, device=dy.device)
    with torch.cuda.device(dy.device):
        _lib.check(_lib.load().b200r_image_stem3x3s2_bwd(dy.data_ptr(), wgt_scaled.data_ptr(), dx.data_ptr(), n, h, w, cout, _lib.f3(std),
                                                         unscale, _stream()))
    return dx
