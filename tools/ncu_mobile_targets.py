"""Driver for ncu: two eager forwards of a mobile-family model (batch 256, split precision).
  ncu ... python tools/ncu_mobile_targets.py [mobilenet_v2|efficientnet_b0] [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustart_b200 import nets  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "mobilenet_v2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
model = nets.build_model(arch, device=dev, seed=0)
img = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device=dev)
for _ in range(2):
    model.forward(img)
torch.cuda.synchronize()
print("done")
