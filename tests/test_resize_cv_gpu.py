"""cv2.resize (INTER_NEAREST / INTER_LINEAR / INTER_AREA / INTER_LANCZOS4 bit for bit, INTER_CUBIC within IPP's bar) on the GPU against cv2 itself (the call of imagenet_s_gen.py:120-148).

GATED like tests/test_token_grad_gpu.py: csrc/resize_cv.cu has run from source on the host emulator only
(tests/test_kernel_emulation_cpu.py); B200R_CV_RESIZE=1 enables the opencv-* ImageNet-S types and these tests."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("hin,win,hout,wout", [(375, 500, 256, 256), (64, 64, 256, 256), (500, 333, 299, 299), (224, 224, 256, 256),
                                               (31, 500, 256, 256), (512, 512, 256, 256), (768, 512, 256, 256), (515, 770, 256, 256)])
def test_resize_cv_matches_cv2(cuda, hin, win, hout, wout):
    cv2 = pytest.importorskip("cv2")
    from robustart_b200 import ops
    img = np.random.RandomState(hin + win).randint(0, 256, (3, hin, win, 3), dtype=np.uint8)
    d = torch.from_numpy(img).to(cuda)
    for name, inter in (("bilinear", cv2.INTER_LINEAR), ("nearest", cv2.INTER_NEAREST), ("area", cv2.INTER_AREA)):
        full = ops.resize_cv_u8(d, (hout, wout), name).cpu().numpy()
        crop = ops.resize_cv_u8(d, (hout, wout), name, crop=(16, 8, 224, 200)).cpu().numpy()
        for i in range(3):
            want = cv2.resize(img[i], (wout, hout), interpolation=inter)
            assert np.array_equal(full[i], want), (name, i)
            assert np.array_equal(crop[i], want[16:240, 8:208]), (name, i)
    full = ops.resize_cv_u8(d, (hout, wout), "lanczos").cpu().numpy()
    cub = ops.resize_cv_u8(d, (hout, wout), "cubic").cpu().numpy()
    for i in range(3):
        assert np.array_equal(full[i], cv2.resize(img[i], (wout, hout), interpolation=cv2.INTER_LANCZOS4)), i
        # INTER_CUBIC runs through closed-source IPP in the opencv-python wheels (a float32 cubic; oracle/cv_resize.py): 1 LSB on <= 1e-4
        # of the pixels here in the container -- the host CPU of the GPU box may pick another IPP code path, hence the wider bar
        dd = np.abs(cub[i].astype(int) - cv2.resize(img[i], (wout, hout), interpolation=cv2.INTER_CUBIC).astype(int))
        assert dd.max() <= 1 and (dd > 0).mean() <= 1e-3, (i, dd.max(), (dd > 0).mean())


def test_imagenet_s_opencv_types(cuda, tmp_path):
    cv2 = pytest.importorskip("cv2")
    from RobustART.noise import AddNoise
    img = np.random.RandomState(9).randint(0, 256, (375, 500, 3), dtype=np.uint8)
    path = str(tmp_path / "a.png")
    cv2.imwrite(path, cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
    for rt, inter in (("opencv-bilinear", cv2.INTER_LINEAR), ("opencv-nearest", cv2.INTER_NEAREST), ("opencv-area", cv2.INTER_AREA)):
        gen = AddNoise("imagenet-s")
        gen.set_config(resize_type=rt, decoder_type="opencv")
        out = gen.add_noise(path)
        assert np.array_equal(out, cv2.resize(img, (256, 256), interpolation=inter)[16:240, 16:240]), rt
