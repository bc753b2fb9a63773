"""AutoAttack pieces: device kernels vs torch restatements of the vendored reference formulas
(autopgd_base.py:198-204,332-338,599-604; square.py:246-254; fab_projections.py:7-59) and end-to-end
invariants of the 'standard' Linf pipeline.  The control flow around these kernels (APGD, FAB-T, the driver) is pinned to
runs of the reference's own vendored code in tests/test_autoattack_cpu.py (goldens from tests/golden/make_golden_attacks.py);
here the kernels are checked against the same torch statements that stand in for them there."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dlr_ref(x, y):
    xs, ind = x.sort(dim=1)
    i = (ind[:, -1] == y).float()
    u = torch.arange(x.shape[0])
    return -(x[u, y] - xs[:, -2] * i - xs[:, -1] * (1. - i)) / (xs[:, -1] - xs[:, -3] + 1e-12)


def _dlr_t_ref(x, y, t):
    xs, _ = x.sort(dim=1)
    u = torch.arange(x.shape[0])
    return -(x[u, y] - x[u, t]) / (xs[:, -1] - .5 * (xs[:, -3] + xs[:, -4]) + 1e-12)


def test_dlr_loss_and_grad(cuda):
    from robustart_b200 import autoattack as AA
    torch.manual_seed(0)
    z = torch.randn(300, 1000, device=cuda) * 3
    y = torch.randint(0, 1000, (300,), device=cuda)
    y[:100] = z[:100].argmax(1)
    t = z.sort(1)[1][:, -3].contiguous()
    for tgt in (None, t):
        zr = z.clone().requires_grad_(True)
        ref = _dlr_ref(zr, y) if tgt is None else _dlr_t_ref(zr, y, tgt)
        ref.sum().backward()
        loss, d = AA._dlr(z, y, tgt)
        assert (loss - ref.detach()).abs().max().item() < 1e-5
        assert (d - zr.grad).abs().max().item() < 1e-5


def test_apgd_step_and_square_kernels(cuda):
    from robustart_b200 import autoattack as AA, _lib
    import ctypes as C
    torch.manual_seed(1)
    n, eps = 5, 4 / 255
    x0 = torch.rand(n, 3, 224, 224, device=cuda)
    xa = (x0 + eps * (2 * torch.rand_like(x0) - 1)).clamp(0, 1)
    xo = (x0 + eps * (2 * torch.rand_like(x0) - 1)).clamp(0, 1)
    g = torch.randn_like(x0)
    step = torch.rand(n, device=cuda) * 2 * eps
    for a in (1.0, 0.75):
        ra, ro = xa.clone(), xo.clone()
        grad2 = ra - ro
        x1 = ra + step.view(-1, 1, 1, 1) * torch.sign(g)
        x1 = torch.clamp(torch.min(torch.max(x1, x0 - eps), x0 + eps), 0.0, 1.0)
        x1 = torch.clamp(torch.min(torch.max(ra + (x1 - ra) * a + grad2 * (1 - a), x0 - eps), x0 + eps), 0.0, 1.0)
        ka, ko = xa.clone(), xo.clone()
        AA._apgd_step_(ka, ko, g, x0, step, eps, a)
        assert torch.equal(ko, xa) and (ka - x1).abs().max().item() < 1e-6
    # square proposal
    xb = (x0 + eps * torch.sign(torch.randn(n, 3, 1, 224, device=cuda))).clamp(0, 1).contiguous()
    signs, vh, vw, s = [1.0, -1.0, 1.0], 17, 100, 40
    nd = torch.zeros(3, 224, 224, device=cuda)
    nd[:, vh:vh + s, vw:vw + s] = 2 * eps * torch.tensor(signs, device=cuda).view(3, 1, 1)
    ref = torch.clamp(torch.min(torch.max(xb + nd, x0 - eps), x0 + eps), 0., 1.)
    out = torch.empty_like(xb)
    _lib.check(_lib.load().b200r_square_propose_linf(xb.data_ptr(), x0.data_ptr(), out.data_ptr(), n, 3, 224, 224, vh, vw, s,
                                                     (C.c_float * 3)(*signs), eps, torch.cuda.current_stream().cuda_stream))
    assert (out - ref).abs().max().item() < 1e-6
    m = torch.tensor([1, 0, 1, 0, 0], dtype=torch.bool, device=cuda)
    dst = x0.clone()
    AA._masked_rows_(dst, xb, m)
    assert torch.equal(dst[m], xb[m]) and torch.equal(dst[~m], x0[~m])


def _proj_cases(cuda):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "attack_pieces.npz"))
    yield "golden", torch.from_numpy(g["proj_t"]).to(cuda), torch.from_numpy(g["proj_w"]).to(cuda), torch.from_numpy(g["proj_b"]).to(cuda), g["proj"]
    gen = torch.Generator().manual_seed(2)
    for rows, dim, kind in ((6, 150528, "inside"), (4, 150528, "far"), (5, 3001, "inside"), (3, 7, "near")):
        t = torch.rand(rows, dim, generator=gen)
        w = torch.randn(rows, dim, generator=gen)
        w[:, ::11] = 0
        if kind == "inside":        # a hyperplane through the box: the threshold search (c2) or the no-saturation case (c_l)
            b = (w * torch.rand(rows, dim, generator=gen)).sum(1)
        elif kind == "near":
            b = (w * t).sum(1) + 0.01 * torch.randn(rows, generator=gen)
        else:                       # a hyperplane the box cannot reach: every coordinate goes to its bound
            b = (w * t).sum(1) + w.abs().sum(1) * 2
        t[0, :5] = 0.0              # coordinates already at a bound (p = 0)
        yield "%s %dx%d" % (kind, rows, dim), t.to(cuda), w.to(cuda), b.to(cuda), None


def test_fab_projection_kernel_matches_reference_statement(cuda):
    """K-FABproj (csrc/attack_proj.cu, sort-free) against the reference's function: its golden output (attack_pieces.npz, produced
    by fab_projections.projection_linf itself) and, at full D = 150 528 and odd sizes, against the reference-pinned restatement
    (oracle.autoattack.projection_linf) run on the CPU.  Bar: 1e-6 absolute (the reference's own float32 cumulative sums)."""
    from oracle import autoattack as OAA
    from robustart_b200 import ops
    for name, t, w, b, want in _proj_cases(cuda):
        d, dmax, passes = ops.fab_projection_linf(t, w, b, want_dmax=True, want_passes=True)
        ref = want if want is not None else OAA.projection_linf(t.cpu(), w.cpu(), b.cpu()).numpy()
        err = np.abs(d.cpu().numpy() - ref).max()
        assert err <= 1e-6, (name, err)
        assert np.abs(dmax.cpu().numpy() - np.abs(ref).max(1)).max() <= 1e-6
        assert int(passes.max()) <= 30, (name, passes.tolist())
        p = t + d
        assert p.min().item() >= -1e-6 and p.max().item() <= 1 + 1e-6
    # FAB's update in one launch (fab_base.py:200-232)
    torch.manual_seed(5)
    x1, x0 = torch.rand(6, 3, 32, 32, device=cuda), torch.rand(6, 3, 32, 32, device=cuda)
    d1, d2 = torch.randn(6, 3072, device=cuda) * 0.05, torch.randn(6, 3072, device=cuda) * 0.02
    d2[0] = 0
    want = OAA.fab_combine(x1.cpu(), d1.cpu(), x0.cpu(), d2.cpu(), 1.05, 0.1)
    got = ops.fab_combine_linf_(x1.clone(), d1, x0, d2, d1.abs().max(1)[0], d2.abs().max(1)[0], 1.05, 0.1)
    assert (got.cpu() - want).abs().max().item() <= 1e-6


def test_l1_projection_kernel_matches_reference_statement(cuda):
    """K-L1proj (sort-free) against the reference's own vendored L1_projection: golden outputs (tests/golden/l1_projection.npz)
    and the pinned restatement at image size."""
    from oracle import autoattack as OAA
    from robustart_b200 import ops
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "l1_projection.npz"))
    for k in range(4):
        x, y, eps = torch.from_numpy(g["x%d" % k]).to(cuda), torch.from_numpy(g["y%d" % k]).to(cuda), float(g["eps%d" % k])
        d = ops.l1_projection(x, y, eps).cpu().numpy()
        assert np.abs(d - g["d%d" % k]).max() <= 2e-6, (k, np.abs(d - g["d%d" % k]).max())
    gen = torch.Generator().manual_seed(8)
    x = torch.rand(3, 150528, generator=gen)
    y = torch.randn(3, 150528, generator=gen) * 0.05
    for eps in (12.0, 1600.0, 1e6):
        d = ops.l1_projection(x.to(cuda), y.to(cuda), eps).cpu()
        ref = OAA.l1_projection(x, y, eps)
        assert (d - ref).abs().max().item() <= 2e-6, eps
        z = y + d
        assert z.abs().sum(1).max().item() <= eps * (1 + 1e-5) and (x + z).min().item() >= -1e-6 and (x + z).max().item() <= 1 + 1e-6


def test_autoattack_standard_pipeline(cuda):
    from RobustART.noise import AddNoise
    from robustart_b200 import autoattack as AA, ops
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.AdaptiveAvgPool2d(8), torch.nn.Flatten(), torch.nn.Linear(192, 64), torch.nn.Tanh(),
                              torch.nn.Linear(64, 10)).to(cuda).eval()
    x = torch.rand(12, 3, 224, 224, device=cuda)
    y = net(ops.normalize(x)).argmax(1)                      # all clean-correct
    eps = 8 / 255
    aa = AA.AutoAttack(net, eps=eps, verbose=False, n_iter=20, n_queries=60, n_target_classes=3)
    adv = aa.run_standard_evaluation(x, y, bs=12)
    assert (adv - x).abs().max().item() <= eps + 1e-6 and adv.min().item() >= 0 and adv.max().item() <= 1
    accs = [a for _, a in aa.history]
    assert accs[0] == 1.0 and all(a2 <= a1 + 1e-9 for a1, a2 in zip(accs, accs[1:]))
    final = (net(ops.normalize(adv)).argmax(1) == y).float().mean().item()
    assert abs(final - accs[-1]) < 1e-6                      # robust flags agree with the returned adversarials
    assert aa.m.forwards > 0 and aa.m.backwards > 0
    # plugin surface (attack.py:35-38 via AddNoise)
    gen = AddNoise("autoattack_linf")
    gen.set_config(model=net, eps=eps)
    gen.config.update(verbose=False)
    with pytest.raises(NotImplementedError):
        AA.AutoAttack(net, norm="L2", eps=0.5)
