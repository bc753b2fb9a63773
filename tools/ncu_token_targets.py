"""Launch targets for `ncu --set full`: the HBM-bound passes of the token models' gradient step at bench sizes.

  ncu --set full --clock-control none -o gpurun_out/token_passes python tools/ncu_token_targets.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustart_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    b, t, c, tp = 64, 196, 768, 256                       # Mixer-B/16, batch 64
    x = ops.split_f32(torch.randn(b * t, c, device=dev))
    g = torch.ones(c, device=dev)
    for _ in range(2):
        y = ops.tokens_to_channels(x, b, t, c, tp)
        ops.channels_to_tokens_add(y, x, b, t, c, tp)
    rows = 128 * 197                                      # ViT-B/16, batch 128
    xv = ops.split_f32(torch.randn(rows, c, device=dev))
    dy = ops.split_f32(torch.randn(rows, c, device=dev))
    for _ in range(2):
        ops.layernorm(xv, g, g, eps=1e-5)
        ops.layernorm_bwd(dy, xv, g, eps=1e-5, add=dy)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
