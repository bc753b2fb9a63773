"""Small driver for ncu: a few gaussian_noise launches (N=256) and, with --model, eager ResNet-50 forwards (N=256) in the
precision given by --passes (16 = fp16 single plane, 3 = split-bf16)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustart_b200 import nets, ops
dev = torch.device("cuda", 0)
imgs = [torch.randint(0, 256, (256, 224, 224, 3), dtype=torch.uint8, device=dev) for _ in range(4)]
out = torch.empty_like(imgs[0])
names = [a for a in sys.argv[1:] if not a.startswith("--") and not a.isdigit()] or ["gaussian_noise"]
for name in names:
    for i in range(int(os.environ.get("NREP", "6"))):
        ops.corrupt_u8(imgs[i % 4], name, 1 + i % 5, seed=i, out=out)
if "--model" in sys.argv:
    passes = int(sys.argv[sys.argv.index("--passes") + 1]) if "--passes" in sys.argv else 16
    model = nets.build_model("resnet50", device=dev, passes=passes)
    for _ in range(2):
        model.forward(imgs[0])
torch.cuda.synchronize()
print("done")
