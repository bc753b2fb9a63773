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


def projection_linf(t, w, b):
    """fab_projections.py:7-59 restated (argsort / cumulative sums / index bisection): project the rows of t onto
    {x: <w,x> = b} intersected with the box [0,1]^d, minimising the Linf norm of the step.  The product solves the same
    problem without a sort (csrc/attack_proj.cu); this statement is its checker.  Pinned against the reference's own function
    on tests/golden/attack_pieces.npz (tests/test_oracle_cpu.py)."""
    import math
    w, b = w.clone(), b.clone()
    sign = 2 * ((w * t).sum(1) - b >= 0) - 1
    w.mul_(sign.unsqueeze(1))
    b.mul_(sign)
    a = (w < 0).float()
    d = (a - t) * (w != 0).float()
    p = a - t * (2 * a - 1)
    indp = torch.argsort(p, dim=1)
    b = b - (w * t).sum(1)
    b0 = (w * d).sum(1)
    indp2 = indp.flip((1,))
    ws = w.gather(1, indp2)
    bs2 = -ws * d.gather(1, indp2)
    s = torch.cumsum(ws.abs(), dim=1)
    sb = torch.cumsum(bs2, dim=1) + b0.unsqueeze(1)
    b2 = sb[:, -1] - s[:, -1] * p.gather(1, indp[:, 0:1]).squeeze(1)
    c_l = b - b2 > 0
    c2 = (b - b0 > 0) & (~c_l)
    lb = torch.zeros(int(c2.sum()), device=t.device)
    ub = torch.full_like(lb, w.shape[1] - 1)
    indp_, sb_, s_, p_, b_ = indp[c2], sb[c2], s[c2], p[c2], b[c2]
    for _ in range(math.ceil(math.log2(w.shape[1]))):
        c4 = torch.floor((lb + ub) / 2)
        c2i = c4.long().unsqueeze(1)
        indcurr = indp_.gather(1, indp_.size(1) - 1 - c2i)
        bb = (sb_.gather(1, c2i) - s_.gather(1, c2i) * p_.gather(1, indcurr)).squeeze(1)
        cc = b_ - bb > 0
        lb = torch.where(cc, c4, lb)
        ub = torch.where(cc, ub, c4)
    lb = lb.long()
    if c_l.any():
        lm = torch.clamp_min((b[c_l] - sb[c_l, -1]) / (-s[c_l, -1]), min=0).unsqueeze(-1)
        d[c_l] = (2 * a[c_l] - 1) * lm
    u = torch.arange(lb.shape[0], device=t.device)
    lm = torch.clamp_min((b[c2] - sb[c2][u, lb]) / (-s[c2][u, lb]), min=0).unsqueeze(-1)
    d[c2] = torch.min(lm, d[c2]) * a[c2] + torch.max(-lm, d[c2]) * (1 - a[c2])
    return d * (w != 0).float()


def fab_combine(x1, d1, x0, d2, eta, alpha_max):
    """fab_base.py:200-232: step-size rule from the two projection lengths + convex combination + clamp."""
    bs = x1.shape[0]
    a1 = d1.reshape(bs, -1).abs().max(dim=1)[0].clamp_min(1e-8).view(-1, *([1] * (x1.dim() - 1)))
    a2 = d2.reshape(bs, -1).abs().max(dim=1)[0].clamp_min(1e-8).view(-1, *([1] * (x1.dim() - 1)))
    alpha = torch.min(torch.max(a1 / (a1 + a2), torch.zeros_like(a1)), alpha_max * torch.ones_like(a1))
    return ((x1 + eta * d1.reshape(x1.shape)) * (1 - alpha) + (x0 + d2.reshape(x1.shape) * eta) * alpha).clamp(0.0, 1.0)


def l1_projection(x2, y2, eps1):
    """L1_projection of autopgd_base.py:19-83 restated with the sort the reference uses: delta such that ||y2 + delta||_1 <= eps1 and
    0 <= x2 + y2 + delta <= 1.  Pinned against the reference's own function on tests/golden/l1_projection.npz."""
    x = x2.clone().float().view(x2.shape[0], -1)
    y = y2.clone().float().view(y2.shape[0], -1)
    sigma = y.clone().sign()
    u = torch.min(1 - x - y, x + y)
    u = torch.min(torch.zeros_like(y), u)
    l = -torch.clone(y).abs()
    d = u.clone()
    bs, indbs = torch.sort(-torch.cat((u, l), 1), dim=1)
    bs2 = torch.cat((bs[:, 1:], torch.zeros(bs.shape[0], 1)), 1)
    inu = 2 * (indbs < u.shape[1]).float() - 1
    size1 = inu.cumsum(dim=1)
    s1 = -u.sum(dim=1)
    c = eps1 - y.clone().abs().sum(dim=1)
    c5 = s1 + c < 0
    c2 = c5.nonzero().squeeze(1)
    s = s1.unsqueeze(-1) + torch.cumsum((bs2 - bs) * size1, dim=1)
    if c2.numel() != 0:
        lb = torch.zeros_like(c2).float()
        ub = torch.ones_like(lb) * (bs.shape[1] - 1)
        nitermax = int(torch.ceil(torch.log2(torch.tensor(bs.shape[1]).float())))
        for _ in range(nitermax):
            counter4 = torch.floor((lb + ub) / 2.)
            counter2 = counter4.long()
            c8 = s[c2, counter2] + c[c2] < 0
            lb = torch.where(c8, counter4, lb)
            ub = torch.where(c8, ub, counter4)
        lb2 = lb.long()
        alpha = (-s[c2, lb2] - c[c2]) / size1[c2, lb2 + 1] + bs2[c2, lb2]
        d[c2] = -torch.min(torch.max(-u[c2], alpha.unsqueeze(-1)), -l[c2])
    return (sigma * d).view(x2.shape)
