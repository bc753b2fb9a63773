"""What bounds the split-precision convolution GEMM: operand delivery or the tensor pipe?  Times a few ResNet-50 layer shapes
(batch 256) with passes = 3 (three MMAs per k-block, hi and lo planes loaded) and passes = 1 (one MMA per k-block, hi plane only:
half the operand bytes, a third of the MMA work).   python tools/gemm_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustart_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
SHAPES = [(56, 64, 64, 3, False), (28, 128, 128, 3, False), (14, 256, 256, 3, False), (7, 512, 512, 3, False),
          (14, 256, 1024, 1, True), (14, 1024, 256, 1, False), (28, 128, 512, 1, True), (7, 512, 2048, 1, True), (56, 64, 256, 1, True)]


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


n = 256
for hw, cin, cout, k, res in SHAPES:
    x = ops.to_planes(torch.randn(n, hw, hw, cin, device=dev))
    w = ops.to_planes(torch.randn(cout, k, k, cin, device=dev) * 0.05)
    r = ops.to_planes(torch.randn(n, hw, hw, cout, device=dev)) if res else None
    b = torch.zeros(cout, device=dev)
    out = torch.empty((2, n, hw, hw, cout), dtype=torch.int16, device=dev)
    t3 = timed(lambda: ops.conv2d_nhwc(x, w, None, b, r, pad=k // 2, act="relu", passes=3, out=out))
    t1 = timed(lambda: ops.conv2d_nhwc(x, w, None, b, r, pad=k // 2, act="relu", passes=1, out=out))
    flops = 2.0 * n * hw * hw * cin * cout * k * k
    tiles = (n * hw * hw / 128.0) * max(1, cout / 128.0)
    kb = k * k * (cin // 64)
    bn = 64 if cout <= 64 else 128
    op3 = tiles * kb * (128 * 64 * 2 * 2 + bn * 64 * 2 * 2)
    print("%dx%d %4d->%4d @%2d res=%d   passes3 %.3f ms (%.0f TF/s alg, operands %.1f TB/s)   passes1 %.3f ms (operands %.1f TB/s)   ratio %.2f" % (
        k, k, cin, cout, hw, res, t3, flops / t3 / 1e9, op3 / t3 / 1e9, t1, op3 / 2 / t1 / 1e9, t3 / t1))
