"""Driver for ncu captures of the gradient-pass / attack kernels: one ResNet-50 step, one ViT-B/16 step, one MobileNetV2 step, the FAB /
L1 projections.  ncu -k regex:... python tools/ncu_backward_targets.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustart_b200 import nets, ops  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
for arch, n in (("resnet50", 32), ("vit_b16_224", 16), ("mobilenet_v2", 16)):
    m = nets.build_model(arch, device=dev, passes=3)
    x = torch.rand(n, 3, 224, 224, device=dev)
    y = torch.randint(0, 1000, (n,), device=dev)
    for _ in range(2):
        lg, saved = m.forward_saved(x)
        g = m.input_grad(ops.ce_loss_grad(lg, y)[1], saved)
        ops.pgd_step_linf_(x.clone(), g, x, 0.001, 4 / 255)
    del m, saved
t = torch.rand(32, 150528, device=dev)
w = torch.randn(32, 150528, device=dev)
b = (w * torch.rand(32, 150528, device=dev)).sum(1)
for _ in range(2):
    d, dm = ops.fab_projection_linf(t, w, b, want_dmax=True)
    ops.fab_combine_linf_(t[:16].clone(), d[:16], t[16:], d[16:], dm[:16], dm[16:], 1.05, 0.1)
    ops.l1_projection(t, (w * 0.05).contiguous(), 12.0)
    ops.pgd_step_l1_(t.clone(), w, t, 120.0, 1600.0)
torch.cuda.synchronize()
print("done")
