"""The PRODUCT's APGD control flow (robustart_b200/autoattack.py: checkpoints, oscillation test, step halving, restart from the
best point, best-adversarial bookkeeping) against the reference's own `attack_single_run` (autopgd_base.py:208-448), on CPU.

Goldens: tests/golden/attack_pieces.npz, produced by running the reference's vendored APGDAttack / APGDAttack_targeted with
device='cpu' on a tiny seeded CNN (tests/golden/make_golden_attacks.py).  The three device kernels the loop launches are
replaced here -- in the test only -- by the torch statements they implement (each kernel is checked against the same
statements on the GPU in tests/test_autoattack_gpu.py); everything else is the code that runs in production."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "attack_pieces.npz"))


def _tiny_model():
    torch.manual_seed(123)
    m = nn.Sequential(nn.Conv2d(3, 8, 3, 2, 1), nn.ReLU(), nn.Conv2d(8, 16, 3, 2, 1), nn.ReLU(), nn.AdaptiveAvgPool2d(1), nn.Flatten(),
                      nn.Linear(16, 10)).eval()
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def _step(x_adv, x_old, grad, x0, step, eps, a):          # b200r_apgd_step_linf: autopgd_base.py:332-338, in place
    s = step.view(-1, 1, 1, 1)
    grad2 = x_adv - x_old
    x_old.copy_(x_adv)
    x1 = x_adv + s * torch.sign(grad)
    x1 = torch.clamp(torch.min(torch.max(x1, x0 - eps), x0 + eps), 0.0, 1.0)
    x1 = torch.clamp(torch.min(torch.max(x_adv + (x1 - x_adv) * a + grad2 * (1 - a), x0 - eps), x0 + eps), 0.0, 1.0)
    x_adv.copy_(x1)


def _masked(dst, src, mask):                               # b200r_masked_rows_copy
    dst[mask] = src[mask]


class _FakeLib:                                            # b200r_square_propose_linf: square.py:246-254 on the product's buffers
    def b200r_square_propose_linf(self, xb_p, xc_p, xn_p, n, c, h, w, vh, vw, s, signs, eps, stream):
        import ctypes as C
        view = lambda p: torch.frombuffer((C.c_float * (n * c * h * w)).from_address(p), dtype=torch.float32).view(n, c, h, w)
        xb, xc, xn = view(xb_p), view(xc_p), view(xn_p)
        delta = torch.zeros(c, h, w)
        delta[:, vh:vh + s, vw:vw + s] = 2. * eps * torch.tensor(list(signs)).view(c, 1, 1)
        xn.copy_(torch.clamp(torch.min(torch.max(xb + delta, xc - eps), xc + eps), 0., 1.))
        return 0


class _TorchModel:                                         # autoattack._Model.loss_and_grad with autograd instead of the kernels
    def __init__(self, model):
        self.model = model

    def loss_and_grad(self, x, y, kind, target=None):
        from oracle import autoattack as OAA
        x = x.clone().requires_grad_(True)
        lg = self.model(x)
        if kind == "ce":
            li = F.cross_entropy(lg, y, reduction="none")
        elif kind == "dlr":
            li = OAA.dlr_loss(lg, y)
        else:
            li = OAA.dlr_loss_targeted(lg, y, target)
        (g,) = torch.autograd.grad(li.sum(), x)
        return lg.detach(), li.detach(), g


@pytest.mark.parametrize("loss,key", [("ce", "ce"), ("dlr", "dlr"), ("dlr-targeted", "t")])
def test_apgd_control_flow_matches_reference_run(monkeypatch, loss, key):
    from robustart_b200 import autoattack as AA
    monkeypatch.setattr(AA, "_apgd_step_", _step)
    monkeypatch.setattr(AA, "_masked_rows_", _masked)
    eps, n_iter, seed = G["apgd_cfg"].tolist()
    x, y = torch.from_numpy(G["apgd_x"]), torch.from_numpy(G["apgd_y"])
    target = torch.from_numpy(G["apgd_t_target"]) if loss == "dlr-targeted" else None
    apgd = AA.APGD(_TorchModel(_tiny_model()), eps, n_iter=int(n_iter), loss=loss)
    assert (apgd.n_iter_2, apgd.n_iter_min, apgd.size_decr) == (4, 1, 1)          # autopgd_base.py:163-165 at 20 iterations
    torch.manual_seed(int(seed))
    x_best, acc, loss_best, x_best_adv = apgd.single_run(x.clone(), y, target)
    assert torch.equal(acc, torch.from_numpy(G["apgd_%s_acc" % key]))
    assert acc.any() and not acc.all()                                              # the golden has robust and fooled samples
    assert np.abs(x_best.numpy() - G["apgd_%s_x_best" % key]).max() <= 1e-6
    assert np.abs(x_best_adv.numpy() - G["apgd_%s_x_best_adv" % key]).max() <= 1e-6
    assert np.abs(loss_best.numpy() - G["apgd_%s_loss" % key]).max() <= 1e-5
    assert (x_best_adv - x).abs().max().item() <= eps + 1e-6 and x_best_adv.min() >= 0 and x_best_adv.max() <= 1


class _TorchFabModel:
    def __init__(self, model):
        self.model = model

    def logits(self, x):
        with torch.no_grad():
            return self.model(x)

    def loss_and_grad(self, x, y, kind, target=None):      # "fab-diff": -(z_y - z_t), fab_pt.py:102-117
        assert kind == "fab-diff"
        x = x.clone().requires_grad_(True)
        lg = self.model(x)
        u = torch.arange(lg.shape[0])
        li = -(lg[u, y] - lg[u, target])
        (g,) = torch.autograd.grad(li.sum(), x)
        return lg.detach(), li.detach(), g


def _fab_kernels_on_host(monkeypatch):
    """b200r_fab_projection_linf / b200r_fab_combine_linf replaced -- in these CPU tests only -- by the reference-pinned torch
    statements they are checked against on the GPU (oracle.autoattack.projection_linf / fab_combine)."""
    from oracle import autoattack as OAA
    from robustart_b200 import ops

    def proj(t, w, b, want_dmax=False, want_passes=False):
        d = OAA.projection_linf(t, w, b)
        return (d, d.abs().max(dim=1)[0]) if want_dmax else d

    def combine_(x1, d1, x0, d2, dmax1, dmax2, eta, alpha_max):
        x1.copy_(OAA.fab_combine(x1, d1, x0, d2, eta, alpha_max))
        return x1
    monkeypatch.setattr(ops, "fab_projection_linf", proj)
    monkeypatch.setattr(ops, "fab_combine_linf_", combine_)


@pytest.mark.parametrize("tc", [2, 3])
def test_fab_targeted_single_run_matches_reference_run(tc, monkeypatch):
    """The product's FAB-T iteration (linearised boundary, projection_linf of x1 and x0, convex combination, overshoot,
    backward step, best-so-far bookkeeping) against FABAttack_PT.attack_single_run (fab_base.py:84-270) on CPU."""
    from robustart_b200 import autoattack as AA
    _fab_kernels_on_host(monkeypatch)
    eps = G["apgd_cfg"].tolist()[0]
    x, y = torch.from_numpy(G["apgd_x"]), torch.from_numpy(G["apgd_y"])
    out = AA.FABT(_TorchFabModel(_tiny_model()), eps, n_iter=15).single_run(x.clone(), y.clone(), tc)
    want = G["fab_t%d" % tc]
    assert np.abs(out.numpy() - want).max() <= 1e-6
    moved = np.abs(want - G["apgd_x"]).reshape(8, -1).max(1)
    assert (moved > 0).sum() >= 4 and moved[7] == 0          # the sample that starts misclassified is left alone


def test_autoattack_driver_matches_reference_run(monkeypatch):
    """The product's AutoAttack driver (robust-set bookkeeping, per-attack seeding, APGD-CE -> APGD-T over 9 target classes ->
    FAB-T over 9 target classes -> Square) against the reference's run_standard_evaluation on CPU: same adversarials, same robust
    accuracy after every stage (autoattack.py:90-211; autopgd_base.py:450-529,610-690; fab_base.py:274-334)."""
    from oracle import autoattack as OAA
    from robustart_b200 import autoattack as AA
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    net = _tiny_model()

    class TorchNormalizedModel:                  # stands in for autoattack._Model (NormalizeModel + kernels) in this test only
        def __init__(self, _):
            self.forwards = self.backwards = 0

        def _f(self, x):
            return net((x - mean) / std)

        def logits(self, x):
            with torch.no_grad():
                return self._f(x)

        def loss_and_grad(self, x, y, kind, target=None):
            x = x.clone().requires_grad_(True)
            lg = self._f(x)
            u = torch.arange(lg.shape[0])
            li = (F.cross_entropy(lg, y, reduction="none") if kind == "ce" else OAA.dlr_loss(lg, y) if kind == "dlr"
                  else OAA.dlr_loss_targeted(lg, y, target) if kind == "dlr-targeted" else -(lg[u, y] - lg[u, target]))
            (g,) = torch.autograd.grad(li.sum(), x)
            return lg.detach(), li.detach(), g

    _fab_kernels_on_host(monkeypatch)
    monkeypatch.setattr(AA, "_apgd_step_", _step)
    monkeypatch.setattr(AA, "_masked_rows_", _masked)
    monkeypatch.setattr(AA, "_Model", TorchNormalizedModel)
    monkeypatch.setattr(AA._lib, "load", lambda: _FakeLib())
    monkeypatch.setattr(AA, "_stream", lambda: None)
    eps, n_iter, seed = G["aa_cfg"].tolist()
    x, y = torch.from_numpy(G["aa_x"]), torch.from_numpy(G["aa_y"])
    aa = AA.AutoAttack(net, norm="Linf", eps=eps, seed=int(seed), verbose=False, n_iter=int(n_iter), n_queries=200)   # the standard four
    adv = aa.run_standard_evaluation(x.clone(), y.clone(), bs=10)
    assert np.abs(adv.numpy() - G["aa_adv"]).max() <= 1e-6
    assert [h[0] for h in aa.history] == ["clean", "apgd-ce", "apgd-t", "fab-t", "square"]
    assert np.allclose([h[1] for h in aa.history[1:]], G["aa_robust_after"], atol=1e-6)
    assert G["aa_robust_after"][0] > G["aa_robust_after"][1]          # the golden exercises the shrinking robust set
    assert (adv - x).abs().max().item() <= eps + 1e-6


def test_square_attack_matches_reference_run(monkeypatch):
    """The product's Square attack (stripe initialisation, p schedule, window draws from the torch generator in the reference's
    order, acceptance rule, active-set bookkeeping) against SquareAttack.perturb (square.py:221-294) on CPU with the same seed.
    The proposal kernel is replaced -- in this test only -- by the torch statements of square.py:246-254 acting on the
    buffers the product passes."""
    from robustart_b200 import autoattack as AA
    monkeypatch.setattr(AA._lib, "load", lambda: _FakeLib())
    monkeypatch.setattr(AA, "_stream", lambda: None)
    monkeypatch.setattr(AA, "_masked_rows_", _masked)
    eps, n_queries, seed = G["square_cfg"].tolist()
    x, y = torch.from_numpy(G["apgd_x"]), torch.from_numpy(G["apgd_y"])
    out = AA.Square(_TorchFabModel(_tiny_model()), eps, n_queries=int(n_queries)).perturb(x.clone(), y.clone(), int(seed))
    assert np.abs(out.numpy() - G["square_adv"]).max() <= 1e-7
    moved = np.abs(G["square_adv"] - G["apgd_x"]).reshape(8, -1).max(1)
    assert 0 < (moved > 0).sum() < 7                           # some samples fooled, some robust within the budget
