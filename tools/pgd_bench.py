"""PGD-Linf k-step eval loop: images/s and tensor-pipe fraction, native dgrad vs autograd twin.

  python tools/pgd_bench.py [--n 128] [--steps 10] [--arch resnet50|resnet18|vit_b16_224|mixer_b16_224] [--autograd]

The token archs (BASELINE configs[2] / [4]) use the input-gradient pass of DESIGN.md section 4d as the native source.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustart_b200 import attacks, nets, ops, torch_models  # noqa: E402

FWD_GFLOP = {"resnet50": 8.18, "resnet18": 3.62, "vit_b16_224": 35.1, "mixer_b16_224": 25.2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--arch", default="resnet50")
    ap.add_argument("--autograd", action="store_true")
    ap.add_argument("--passes-bwd", type=int, default=3)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--passes", type=int, default=3, help="3 = split-bf16, 16 = fp16 single plane")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    if a.arch in nets._TOKEN_ARCHS:
        sd = nets.random_token_state_dict(nets._TOKEN_ARCHS[a.arch][1](), 0)
    else:
        sd = nets.random_state_dict(nets.resnet_spec(a.arch), 0)
    net = nets.build_model(a.arch, sd, device=dev, passes=a.passes)
    if a.autograd:
        twin = torch_models.build(a.arch, nets._strip_prefix(sd)).to(dev).eval()
        src = attacks.PyTorchModel(twin, preprocessing=dict(mean=ops.IMAGENET_MEAN, std=ops.IMAGENET_STD, axis=-3))
    else:
        src = attacks.NativeModel(net, passes_bwd=a.passes_bwd)
    x = torch.rand(a.n, 3, 224, 224, device=dev)
    y = torch.randint(0, 1000, (a.n,), device=dev)
    counters = torch.zeros(3, dtype=torch.int64, device=dev)

    def loop():
        adv = attacks.pgd_linf(x, y, src, 4 / 255, 3 / 40, a.steps, seed=0)
        ops.topk_count_(counters, net.forward(adv), y)

    loop()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        loop()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    gflop = (2 * a.steps + 1) * FWD_GFLOP[a.arch] * a.n      # k*(fwd + dgrad) + 1 fwd, SURVEY 8(d)
    print(json.dumps({"loop": "pgd_linf %d-step eval, %s, batch %d" % (a.steps, a.arch, a.n),
                      "source": "autograd twin" if a.autograd else "native dgrad (passes=%d)" % a.passes,
                      "ms_per_batch": ms, "images_per_s": a.n / ms * 1e3, "algorithmic_tflops": gflop / ms,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))


if __name__ == "__main__":
    main()
