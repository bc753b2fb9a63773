"""Forward throughput of the five benchmark architectures on the sm_100a kernels (CUDA-event timed, uint8 NHWC
batches rotating over > L2).  Writes gpurun_out/model_bench.json.  Not the headline bench (bench.py is).

  python tools/model_bench.py [arch ...] [--n 256]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustart_b200 import nets  # noqa: E402

GFLOP = {"resnet18": 3.62, "resnet50": 8.18, "mobilenet_v2": 0.60, "efficientnet_b0": 0.78,
         "vit_b16_224": 35.1, "mixer_b16_224": 25.2}     # SURVEY 8(d), per image


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(sys.argv[sys.argv.index("--n") + 1]) if "--n" in sys.argv else 256
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    ins = [torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g) for _ in range(4)]
    res = {}
    for arch in args or list(GFLOP):
        try:
            m = nets.build_model(arch, device=dev, seed=0)
        except Exception as ex:  # unknown name on this build
            print(arch, "skipped:", ex, flush=True)
            continue
        for i in range(3):
            m.forward(ins[i % 4])
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 8
        s.record()
        for i in range(reps):
            m.forward(ins[i % 4])
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / reps
        res[arch] = dict(batch=n, ms_per_batch=ms, img_per_s=n / ms * 1e3, algorithmic_tflops=GFLOP[arch] * n / ms,
                         launches=m.launches_per_forward() if hasattr(m, "launches_per_forward") else None)
        print(arch, res[arch], flush=True)
        del m
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/model_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
