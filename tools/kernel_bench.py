"""Per-kernel device timing of the corruption / attack-step kernels (CUDA events, rotating buffers > L2).
Writes gpurun_out/kernel_bench.json.  Not the headline bench (bench.py is)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustart_b200 import ops  # noqa: E402

PEAKS = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}


def timeit(fn, iters=16, warm=3):
    """Device time per launch: the launches are captured into a CUDA graph (a 256-image corruption is ~12 us at HBM
    speed, shorter than a Python call), the graph is replayed 5 times between CUDA events."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(iters):
                fn(i)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (5 * iters) * 1e-3


def main():
    dev = torch.device("cuda", 0)
    N, R = 256, 8   # 8 rotating batches of 38.5 MB in + out => 616 MB > 126 MB L2
    g = torch.Generator(device=dev).manual_seed(0)
    ins = [torch.randint(0, 256, (N, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g) for _ in range(R)]
    outs = [torch.empty_like(ins[0]) for _ in range(R)]
    res = {}
    names = sys.argv[1:] or ["gaussian_noise", "shot_noise", "impulse_noise", "speckle_noise", "brightness",
                             "saturate", "contrast", "frost", "fog", "pixelate", "jpeg_compression", "gaussian_blur",
                             "defocus_blur", "zoom_blur", "motion_blur", "snow", "glass_blur", "elastic_transform",
                             "spatter"]
    # the ceiling a kernel of this size can reach in this timing harness: a plain device copy of the same 38.5 MB batch
    t = timeit(lambda i: outs[i % R].copy_(ins[i % R]))
    res["copy_256_images"] = dict(us=t * 1e6, gbs=2 * N * 150528 / t / 1e9, frac=2 * N * 150528 / t / 1e9 / PEAKS["hbm_gbs"])
    print("copy", res["copy_256_images"], flush=True)
    for name in names:
        for sev in (1, 3, 5):
            try:
                t = timeit(lambda i: ops.corrupt_u8(ins[i % R], name, sev, seed=i, out=outs[i % R]))
            except NotImplementedError:
                continue
            bytes_alg = 2 * N * 150528
            res["%s/s%d" % (name, sev)] = dict(us=t * 1e6, img_per_s=N / t, gbs=bytes_alg / t / 1e9,
                                               frac=bytes_alg / t / 1e9 / PEAKS["hbm_gbs"])
            print(name, sev, res["%s/s%d" % (name, sev)], flush=True)
    # attack steps
    x0 = [torch.rand(N, 3, 224, 224, device=dev) for _ in range(2)]
    gr = [torch.randn(N, 3, 224, 224, device=dev) for _ in range(2)]
    xs = [t.clone() for t in x0]
    mom = [torch.zeros_like(t) for t in x0]
    eps = 4 / 255
    chw = 3 * 224 * 224 * 4
    for nm, fn, mult in (("pgd_linf_step", lambda i: ops.pgd_step_linf_(xs[i % 2], gr[i % 2], x0[i % 2], eps / 4, eps), 4),
                         ("pgd_l2_step", lambda i: ops.pgd_step_l2_(xs[i % 2], gr[i % 2], x0[i % 2], 0.1, 2.0), 5),
                         ("mim_step", lambda i: ops.mim_step_linf_(xs[i % 2], mom[i % 2], gr[i % 2], x0[i % 2], 0.002, eps, 1.0), 5),
                         ("u8_to_f32nchw", lambda i: ops.u8nhwc_to_f32nchw(ins[i % R], out=x0[i % 2]), None)):
        t = timeit(fn)
        b = N * chw * mult if mult else N * (150528 + chw)
        res[nm] = dict(us=t * 1e6, img_per_s=N / t, gbs=b / t / 1e9, frac=b / t / 1e9 / PEAKS["hbm_gbs"])
        print(nm, res[nm], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/kernel_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
