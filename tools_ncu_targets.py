"""Small driver for `ncu --set full`: a few gaussian_noise launches (N=256) and one eager ResNet-50 forward (N=256)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from robustart_b200 import nets, ops
dev = torch.device("cuda", 0)
imgs = [torch.randint(0, 256, (256, 224, 224, 3), dtype=torch.uint8, device=dev) for _ in range(4)]
out = torch.empty_like(imgs[0])
for i in range(6):
    ops.corrupt_u8(imgs[i % 4], "gaussian_noise", 1 + i % 5, seed=i, out=out)
if "--model" in sys.argv:
    model = nets.build_model("resnet50", device=dev)
    for _ in range(2):
        model.forward(imgs[0])
torch.cuda.synchronize()
print("done")
