"""Oracle restatements of the AutoAttack loss formulas -- TEST INFRASTRUCTURE ONLY (see oracle/imagenet_c.py).

  dlr_loss            RobustART/noise/utils/adv/Attacks/autoattack/autopgd_base.py:198-204
  dlr_loss_targeted   autopgd_base.py:599-604
Pinned by tests/test_oracle_cpu.py::test_attack_pieces_match_reference_code against values produced by the reference's own
vendored functions (tests/golden/make_golden_attacks.py -> attack_pieces.npz).  tests/test_autoattack_gpu.py holds the same
formulas for the device kernels' checks."""
import torch


def dlr_loss(x, y):
    xs, ind = x.sort(dim=1)
    i = (ind[:, -1] == y).float()
    u = torch.arange(x.shape[0])
    return -(x[u, y] - xs[:, -2] * i - xs[:, -1] * (1. - i)) / (xs[:, -1] - xs[:, -3] + 1e-12)


def dlr_loss_targeted(x, y, t):
    xs, _ = x.sort(dim=1)
    u = torch.arange(x.shape[0])
    return -(x[u, y] - x[u, t]) / (xs[:, -1] - .5 * (xs[:, -3] + xs[:, -4]) + 1e-12)
