"""The PRODUCT's PGD / FGSM / MI-FGSM host loops (robustart_b200/attacks.py: start, gradient scaling, preprocessing wrapper,
step order, number of model calls) against runs of the reference's own code, on CPU.

Goldens: tests/golden/attack_pieces.npz (tests/golden/make_golden_attacks.py) -- `_mim_whitebox` (imfgsm_attack.py:62-93) and the
in-repo `pgd_linf_attack` (adv_cls_solver_train_pgd_new.py:67-105) executed from /root/reference on a tiny seeded CNN.  The five
device kernels the loops launch are replaced -- in this test only -- by the torch statements they implement; each kernel is
checked against the same statements on the GPU (tests/test_attacks_gpu.py, tests/test_metrics_gpu.py).  Everything else is the
code that runs in production."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "attack_pieces.npz"))
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def _tiny_model():
    torch.manual_seed(123)
    m = nn.Sequential(nn.Conv2d(3, 8, 3, 2, 1), nn.ReLU(), nn.Conv2d(8, 16, 3, 2, 1), nn.ReLU(), nn.AdaptiveAvgPool2d(1), nn.Flatten(),
                      nn.Linear(16, 10)).eval()
    for p in m.parameters():
        p.requires_grad_(False)
    return m


class _Calls:
    def __init__(self):
        self.n = {}

    def hit(self, k):
        self.n[k] = self.n.get(k, 0) + 1


def _install(monkeypatch, calls):
    from robustart_b200 import attacks as A

    def random_start_linf(x0, eps, *, seed=0, image_offset=0, u=None, clip01=True):    # b200r_random_start_linf with given draws
        assert u is not None, "the CPU stand-in has no device RNG"
        calls.hit("start")
        x = x0 + ((eps - (-eps)) * u + (-eps))
        return x.clamp(0, 1) if clip01 else x

    def pgd_step_linf_(x, g, x0, alpha, eps):                                             # b200r_pgd_step_linf, in place
        calls.hit("pgd_step")
        x.copy_((x0 + (x + alpha * g.sign() - x0).clamp(-eps, eps)).clamp(0, 1))
        return x

    def mim_step_linf_(x, momentum, g, x0, step, eps, decay):                             # b200r_mim_step_linf, in place
        calls.hit("mim_step")
        momentum.copy_(decay * momentum + g / g.abs().mean(dim=[1, 2, 3], keepdim=True))
        x.copy_((x0 + (x + step * momentum.sign() - x0).clamp(-eps, eps)).clamp(0, 1.0))
        return x

    def ce_loss_grad(logits, labels, grad_scale=1.0, want_grad=True):                     # b200r_ce_loss_grad
        calls.hit("ce")
        loss = F.cross_entropy(logits, labels, reduction="none")
        d = (torch.softmax(logits, 1) - F.one_hot(labels, logits.shape[1]).to(logits.dtype)) * grad_scale
        return loss, d

    def normalize(x, mode="normal", mean=MEAN, std=STD, out=None):                        # b200r_normalize_f32nchw
        calls.hit("normalize_" + mode)
        m = torch.tensor(mean, dtype=x.dtype).view(1, 3, 1, 1)
        s = torch.tensor(std, dtype=x.dtype).view(1, 3, 1, 1)
        return {"normal": (x - m) / s, "inv": x * s + m, "grad": x / s}[mode]

    for name, fn in dict(random_start_linf=random_start_linf, pgd_step_linf_=pgd_step_linf_, mim_step_linf_=mim_step_linf_,
                         ce_loss_grad=ce_loss_grad, normalize=normalize).items():
        monkeypatch.setattr(A.ops, name, fn)
    # the product refuses host tensors (attacks._prep); the stand-ins above are what makes a host run meaningful here
    monkeypatch.setattr(A, "_prep", lambda inp, lab: (inp.detach().to(torch.float32).contiguous(),
                                                      lab.detach().to(torch.int64).view(-1).contiguous()))
    return A


def test_product_refuses_host_tensors():
    from robustart_b200 import attacks as A
    with pytest.raises(TypeError):
        A.pgd_linf(torch.zeros(1, 3, 8, 8), torch.zeros(1, dtype=torch.int64), _tiny_model(), 8 / 255, 0.1, 1)


def test_mim_loop_matches_reference_run(monkeypatch):
    calls = _Calls()
    A = _install(monkeypatch, calls)
    eps, steps, step_size, decay = G["mim_cfg"].tolist()
    X, y, u = (torch.from_numpy(G[k]) for k in ("mim_X", "mim_y", "mim_u"))
    adv = A.mim_linf(X, y, _tiny_model(), eps, int(steps), step_size, decay, start_uniform=u)
    want = G["mim_adv"]
    # dL/dx comes from softmax-onehot pushed through the vjp instead of loss.backward(): same mathematics, a few ulp apart, so a
    # coordinate whose momentum is ~0 may step the other way (2*step_size); everything else agrees to rounding
    diff = np.abs(adv.numpy() - want)
    assert (diff > 1e-6).mean() < 2e-3, (diff > 1e-6).mean()
    assert diff.max() <= 2 * int(steps) * step_size + 1e-6
    assert np.abs(adv.numpy() - G["mim_X"]).max() <= eps + 1e-6 and adv.min() >= 0 and adv.max() <= 1
    # imfgsm_attack.py:62-93: one start (not clipped), then per step one normalised forward, one CE gradient (mean loss), one update
    assert calls.n == {"start": 1, "normalize_normal": int(steps), "normalize_grad": int(steps), "ce": int(steps), "mim_step": int(steps)}
    # the start is NOT clipped to [0,1] (imfgsm_attack.py:70-72) -- visible when no step follows
    far = torch.ones_like(X)
    out0 = A.mim_linf(far, y, _tiny_model(), eps, 0, step_size, decay, start_uniform=torch.ones_like(X))
    assert out0.max().item() > 1.0


def test_pgd_linf_loop_matches_reference_loop(monkeypatch):
    calls = _Calls()
    A = _install(monkeypatch, calls)
    eps, rel, steps = G["pgd_cfg"].tolist()
    x, y, u = (torch.from_numpy(G[k]) for k in ("pgd_x", "pgd_y", "pgd_u"))
    net = _tiny_model()
    fmodel = A.PyTorchModel(net, bounds=(0, 1))                    # the in-repo loop feeds the model [0,1] input directly
    adv = A.pgd_linf(x, y, fmodel, eps, rel, int(steps), start_uniform=u)
    diff = np.abs(adv.numpy() - G["pgd_adv"])
    assert (diff > 1e-6).mean() < 2e-3, (diff > 1e-6).mean()
    assert np.abs(adv.numpy() - G["pgd_x"]).max() <= eps + 1e-6 and adv.min() >= 0 and adv.max() <= 1
    assert calls.n == {"start": 1, "ce": int(steps), "pgd_step": int(steps)}
    # the same predictions on the adversarials as the reference loop's
    assert torch.equal(net(adv).argmax(1), net(torch.from_numpy(G["pgd_adv"])).argmax(1))


def test_fgsm_is_one_full_step_without_start(monkeypatch):
    """attack.py:30-33: foolbox LinfFastGradientAttack = one sign step of size eps from the clean image."""
    calls = _Calls()
    A = _install(monkeypatch, calls)
    x, y = torch.from_numpy(G["pgd_x"]), torch.from_numpy(G["pgd_y"])
    net = _tiny_model()
    eps = 8 / 255
    adv = A.fgsm(x, y, A.PyTorchModel(net, preprocessing=dict(mean=MEAN, std=STD, axis=-3)), eps)
    xg = x.clone().requires_grad_(True)
    m, s = torch.tensor(MEAN).view(1, 3, 1, 1), torch.tensor(STD).view(1, 3, 1, 1)
    (g,) = torch.autograd.grad(F.cross_entropy(net((xg - m) / s), y, reduction="sum"), xg)
    want = (x + (eps * g.sign()).clamp(-eps, eps)).clamp(0, 1)
    assert ((adv - want).abs() > 1e-6).float().mean().item() < 2e-3
    assert calls.n == {"normalize_normal": 1, "normalize_grad": 1, "ce": 1, "pgd_step": 1}


def test_model_bounds_and_preprocessing_contract():
    """benchmark_eval_adv.py:198-207 builds fb.PyTorchModel(model, bounds=(0, 1), preprocessing=dict(mean, std, axis=-3)); other
    bounds / axes are refused instead of silently mis-normalised."""
    from robustart_b200 import attacks as A
    with pytest.raises(AssertionError):
        A.PyTorchModel(_tiny_model(), preprocessing=dict(mean=MEAN, std=STD, axis=-1))
    with pytest.raises(NotImplementedError):
        A._bounds01(A.PyTorchModel(_tiny_model(), bounds=(0, 255)))
    with pytest.raises(NotImplementedError):
        A.NativeModel(nn.Linear(2, 2))
    f = A.as_f_model(_tiny_model())
    assert f.preprocessing["mean"] == A.ops.IMAGENET_MEAN and f.bounds == (0, 1)
