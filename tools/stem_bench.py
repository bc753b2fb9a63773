"""Stem timing, fp16 mode, batch 256: one-launch stem (conv1+bn+relu+maxpool) vs the two-launch path it replaces."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustart_b200 import ops
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
imgs = [torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device=dev) for _ in range(4)]
wt = (torch.randn(64, 3, 7, 7, device=dev) / 147 ** 0.5)
wp = ops.to_planes(ops.pack_stem_weight(wt).contiguous(), True)
s, b = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.3
out1 = torch.empty((1, n, 56, 56, 64), dtype=torch.int16, device=dev)
mid = torch.empty((1, n, 112, 112, 64), dtype=torch.int16, device=dev)
out2 = torch.empty_like(out1)


def timed(fn, reps=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


one = timed(lambda i: ops.stem_pool_u8(imgs[i % 4], wp, b, out=out1))
def two_fn(i):
    ops.stem_conv7x7_u8(imgs[i % 4], wp, s, b, act="relu", out=mid)
    ops.maxpool3x3s2(mid, out=out2)
two = timed(two_fn)
alg = n * (150528 + 56 * 56 * 64 * 2)
print(json.dumps({"batch": n, "one_launch_us": one, "two_launch_us": two, "one_launch_GBs": alg / one / 1e3,
                  "issued_TFs": n * 2 * 112 * 112 * 64 * 224 / one / 1e6}))
