"""Noise registry of the B200-native AddNoise.

Same names, default configs and dispatch table as the reference's
RobustART/noise/utils/add_noise_utils.py:7-18,41-50; every entry of `function_dict` lands in
hand-written sm_100a kernels through robustart_b200 (no CPU fallback).
"""
from __future__ import annotations

import numpy as np
import torch

from robustart_b200 import attacks as _atk
from robustart_b200 import ops as _ops

noise_list = ['imagenet-s', 'imagenet-c', 'pgd_linf', 'pgd_l2', 'fgsm', 'autoattack_linf', 'mim_linf', 'pgd_l1']

default_config = {
    'imagenet-s': {'decoder_type': 'pil', 'resize_type': 'pil-bilinear', 'transform_type': 'val'},
    'imagenet-c': {'severity': 1, 'corruption_name': None, 'corruption_number': -1},
    'pgd_linf': {'f_model': None, 'eps': 8 / 255, 'rel_stepsize': 3 / 40, 'steps': 20},
    'pgd_l2': {'f_model': None, 'eps': 8.0, 'rel_stepsize': 3 / 40, 'steps': 20},
    'fgsm': {'f_model': None, 'eps': 8 / 255},
    'autoattack_linf': {'model': None, 'norm': 'Linf', 'eps': 8 / 255, 'version': 'standard', 'verbose': False},
    'mim_linf': {'model': None, 'eps': 8 / 255, 'num_steps': 20, 'step_size': 0.002, 'decay_factor': 1.0},
    'pgd_l1': {'model': None, 'eps': 1600.0, 'input_size': 224, 'eps_step': 120, 'max_iter': 20, 'batch_size': 16},
}

# RNG bookkeeping of the imagenet-c path: the reference pulls from numpy's global generator, so
# np.random.seed(k) makes a run repeatable.  We keep that property: the Philox key is drawn from
# numpy's global generator the first time it is needed and every image gets its own counter stream.
_rng_state = {'seed': None, 'images_done': 0}


def reseed(seed=None):
    """Fix the Philox key of the imagenet-c corruptions (None: re-draw from np.random on next use)."""
    _rng_state['seed'] = None if seed is None else int(seed)
    _rng_state['images_done'] = 0


def _seed():
    if _rng_state['seed'] is None:
        _rng_state['seed'] = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 31)
    return _rng_state['seed']


def add_noise_for_imagenet_c(image, severity=1, corruption_name=None, corruption_number=-1):
    """Batch corruption on the GPU.  `image`:
      * uint8 numpy [n,h,w,3]  -> corrupted IN PLACE and returned (reference contract,
        add_noise_utils.py:27-31); staged through pinned memory;
      * uint8 CUDA tensor [n,h,w,3] -> corrupted in place on the device, no host round trip;
      * a file path -> [h,w,3] numpy (the reference's assert at add_noise.py:36-38 is inverted and makes
        this unreachable there; the documented behaviour is implemented here)."""
    if corruption_name:
        cid = _ops.corruption_id(corruption_name)
    elif corruption_number != -1:
        cid = _ops.corruption_id(int(corruption_number))
    else:
        raise ValueError("Either corruption_name or corruption_number must be passed")
    single = False
    if isinstance(image, str):
        from PIL import Image
        image = np.array(Image.open(image).convert('RGB'))[None]
        single = True
    seed = _seed()
    offset = _rng_state['images_done']
    if isinstance(image, torch.Tensor):
        if not image.is_cuda:
            raise TypeError("torch input must live on the GPU")
        _ops.corrupt_u8(image, cid, severity, seed=seed, image_offset=offset, out=image)
        _rng_state['images_done'] += image.shape[0]
        return image
    arr = np.ascontiguousarray(image)
    if arr.dtype != np.uint8 or arr.ndim != 4 or arr.shape[-1] != 3:
        raise ValueError("imagenet-c expects a uint8 array of shape (n, h, w, 3)")
    dev = torch.device('cuda', torch.cuda.current_device())
    host = torch.from_numpy(arr)
    d = host.pin_memory().to(dev, non_blocking=True) if arr.size else host.to(dev)
    _ops.corrupt_u8(d, cid, severity, seed=seed, image_offset=offset, out=d)
    result = d.cpu().numpy()
    _rng_state['images_done'] += arr.shape[0]
    if single:
        return result[0]
    image[...] = result
    return image


def random_resized_crop_params(height, width, scale=(0.08, 1.0), ratio=(3. / 4., 4. / 3.)):
    """ImageTransfer.get_params (imagenet_s_gen.py:199-239): (i, j, h, w) of a random-sized crop, ten attempts drawn from
    Python's `random` in the reference's order (uniform area, uniform log-ratio, randint i, randint j), then the centre
    fallback at the nearest admissible aspect ratio."""
    import math
    import random
    area = height * width
    for _ in range(10):
        target_area = random.uniform(*scale) * area
        log_ratio = (math.log(ratio[0]), math.log(ratio[1]))
        aspect_ratio = math.exp(random.uniform(*log_ratio))
        w = int(round(math.sqrt(target_area * aspect_ratio)))
        h = int(round(math.sqrt(target_area / aspect_ratio)))
        if 0 < w <= width and 0 < h <= height:
            return random.randint(0, height - h), random.randint(0, width - w), h, w
    in_ratio = float(width) / float(height)
    if in_ratio < min(ratio):
        w = width
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = height
        w = int(round(h * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w


_PIL_RESIZE_TYPES = {'pil-bilinear': 'bilinear', 'pil-nearest': 'nearest', 'pil-box': 'box', 'pil-hamming': 'hamming',
                     'pil-cubic': 'bicubic', 'pil-lanczos': 'lanczos'}


_CV_RESIZE_TYPES = {'opencv-nearest': 'nearest', 'opencv-bilinear': 'bilinear', 'opencv-area': 'area', 'opencv-cubic': 'cubic',
                    'opencv-lanczos': 'lanczos'}


def _cv_resize_enabled():
    # csrc/resize_cv.cu: bit-exact against cv2.resize on the GPU (tests/test_resize_cv_gpu.py); B200R_CV_RESIZE=0 switches it off
    import os
    return os.environ.get('B200R_CV_RESIZE', '1') != '0'


def _imagenet_s_resize(batch, size_hw, resize_type, crop=None):
    """One launch: Image.resize (pil-*) or cv2.resize (opencv-*) of a uint8 NHWC CUDA batch, optionally cropped."""
    if resize_type in _CV_RESIZE_TYPES:
        return _ops.resize_cv_u8(batch, size_hw, _CV_RESIZE_TYPES[resize_type], crop=crop)
    return _ops.resize_u8(batch, size_hw, _PIL_RESIZE_TYPES[resize_type], crop=crop)


def add_noise_for_imagenet_s(image, decoder_type='pil', resize_type='pil-bilinear', transform_type='val', size=224):
    """ImageTransfer(..., return_online=True).getimage() (imagenet_s_gen.py:83-141) for the PIL decoder and the six `pil-*`
    resize types, transform 'val': decode on the host (file parsing, as the reference), then Image.resize to
    (size*8/7, size*8/7) and the centre crop as ONE bit-exact resize kernel launch (b200r_resize_u8).  `image` is a file
    path (the reference's contract) or an already decoded uint8 [h, w, 3] / [n, h, w, 3] array or CUDA tensor.
    transform 'train' = the reference's random resized crop (parameters from Python's `random`) + Image.resize to (size, size).
    Also the `opencv` decoder (cv2.imdecode on the host + BGR->RGB, imagenet_s_gen.py:193-202) and the
    five `opencv-*` resize types (cv2.resize bit for bit; opencv-cubic as IPP's float cubic, b200r_resize_cv_u8).  The ffmpeg
    decoder is not implemented."""
    cv_ok = _cv_resize_enabled()
    if decoder_type != 'pil' and not (cv_ok and decoder_type == 'opencv'):
        raise NotImplementedError("imagenet-s decoder_type=%r: only the PIL decoder is implemented" % decoder_type)
    if resize_type not in _PIL_RESIZE_TYPES and not (cv_ok and resize_type in _CV_RESIZE_TYPES):
        raise NotImplementedError("imagenet-s resize_type=%r: only the pil-* resize types are implemented" % resize_type)
    if transform_type not in ('val', 'train'):
        raise NotImplementedError("imagenet-s transform_type=%r: 'val' (resize + centre crop) and 'train' (random resized crop)" % transform_type)
    if isinstance(image, str):
        if decoder_type == 'opencv':
            import cv2
            image = cv2.cvtColor(cv2.imdecode(np.fromfile(image, dtype=np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
        else:
            from PIL import Image
            with Image.open(image) as im:
                image = np.array(im.convert('RGB'))
    if transform_type == 'train':
        # imagenet_s_gen.py:120-129: crop the box get_params draws from Python's `random`, then resize to (size, size)
        single = image.ndim == 3 if not isinstance(image, torch.Tensor) else image.dim() == 3
        batch = image if not single else image[None]
        outs = []
        for img in batch:
            y0, x0, h, w = random_resized_crop_params(int(img.shape[0]), int(img.shape[1]))
            crop = img[y0:y0 + h, x0:x0 + w]
            d = (crop if isinstance(crop, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(crop)).to(
                torch.device('cuda', torch.cuda.current_device())))
            if not d.is_cuda:
                raise TypeError("torch input must live on the GPU")
            outs.append(_imagenet_s_resize(d.contiguous()[None], (size, size), resize_type)[0])
        out = torch.stack(outs)
        if not isinstance(image, torch.Tensor):
            out = out.cpu().numpy()
        return out[0] if single else out
    first = int(size * 8 / 7)
    i = int(round((first - size) / 2.))
    if isinstance(image, torch.Tensor):
        if not image.is_cuda:
            raise TypeError("torch input must live on the GPU")
        batch = image if image.dim() == 4 else image[None]
        out = _imagenet_s_resize(batch.contiguous(), (first, first), resize_type, crop=(i, i, size, size))
        return out if image.dim() == 4 else out[0]
    arr = np.ascontiguousarray(image)
    if arr.dtype != np.uint8 or arr.ndim not in (3, 4) or arr.shape[-1] != 3:
        raise ValueError("imagenet-s expects a file path or a uint8 array of shape ([n,] h, w, 3)")
    dev = torch.device('cuda', torch.cuda.current_device())
    d = torch.from_numpy(arr if arr.ndim == 4 else arr[None]).to(dev)
    out = _imagenet_s_resize(d, (first, first), resize_type, crop=(i, i, size, size)).cpu().numpy()
    return out if arr.ndim == 4 else out[0]


def pgd_l1(input, label, model, eps, input_size, eps_step, max_iter, batch_size):
    # attack.py:44-49 delegates to ART's ProjectedGradientDescentPyTorch(norm=1) (not vendored, unpinned): same update rules on the
    # device, no numpy round trip (robustart_b200.attacks.pgd_l1)
    return _atk.pgd_l1(input, label, model, eps, input_size, eps_step, max_iter, batch_size)


def autoattack_linf(input, label, model, norm, eps, version, verbose):
    from robustart_b200 import autoattack as _aa
    return _aa.autoattack_linf(input, label, model, norm, eps, version, verbose)


function_dict = {
    'imagenet-s': add_noise_for_imagenet_s,
    'imagenet-c': add_noise_for_imagenet_c,
    'pgd_l1': pgd_l1,
    'pgd_linf': _atk.pgd_linf,
    'pgd_l2': _atk.pgd_l2,
    'fgsm': _atk.fgsm,
    'autoattack_linf': autoattack_linf,
    'mim_linf': _atk.mim_linf,
}
