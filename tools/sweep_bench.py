"""BASELINE configs[3]: EfficientNet-B0 / MobileNetV2 + the full ImageNet-C 15 x 5 sweep on one GPU (the 8-GPU run shards
images, no data-path collective).  Per cell: corruption kernel -> forward (CUDA graph) -> counters, batch 256; reports
corrupted images/s over the whole sweep and the share of time spent in the corruption kernels.
  python tools/sweep_bench.py [batch]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustart_b200 import nets, ops  # noqa: E402
from robustart_b200.solver import EvalSolver  # noqa: E402

dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
groups = EvalSolver.IMAGENET_C_GROUPS
cells = [(t, s) for g in ("noise", "blur", "weather", "digital") for t in groups[g] for s in (1, 2, 3, 4, 5)]
img = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device=dev)
labels = torch.randint(0, 1000, (n,), device=dev)
work = torch.empty_like(img)
out = {}
ev = lambda: torch.cuda.Event(enable_timing=True)
# corruption kernels alone (same for both models)
for t, s in cells:
    ops.corrupt_u8(img, t, s, seed=1, image_offset=0, out=work)          # first use builds per-severity tables
torch.cuda.synchronize()
a, b = ev(), ev()
a.record()
for t, s in cells:
    ops.corrupt_u8(img, t, s, seed=1, image_offset=0, out=work)
b.record()
torch.cuda.synchronize()
corrupt_ms = a.elapsed_time(b)
for arch in ("mobilenet_v2", "efficientnet_b0"):
    model = nets.build_model(arch, device=dev, seed=0)
    run = model.graphed(work)
    counters = torch.zeros((len(cells), 3), dtype=torch.int64, device=dev)
    for _ in range(2):
        for ci, (t, s) in enumerate(cells[:3]):
            ops.corrupt_u8(img, t, s, seed=1, image_offset=0, out=run.static_in)
            ops.topk_count_(counters[ci], run(run.static_in, copy_in=False), labels)
    torch.cuda.synchronize()
    counters.zero_()
    t0 = time.perf_counter()
    a, b = ev(), ev()
    a.record()
    for ci, (t, s) in enumerate(cells):
        ops.corrupt_u8(img, t, s, seed=1, image_offset=0, out=run.static_in)
        ops.topk_count_(counters[ci], run(run.static_in, copy_in=False), labels)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    out[arch] = {"batch": n, "cells": len(cells), "sweep_ms": ms, "wall_ms": (time.perf_counter() - t0) * 1e3,
                 "corrupted_images_per_s": n * len(cells) / ms * 1e3, "corruption_kernels_ms": corrupt_ms,
                 "corruption_share": corrupt_ms / ms, "count_check": int(counters[:, 2].sum().item())}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep_bench.json", "w"), indent=1)
