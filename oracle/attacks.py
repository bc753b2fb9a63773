"""Reference restatement of the gradient attacks -- TEST INFRASTRUCTURE ONLY (see oracle/imagenet_c.py).

foolbox 3.3.1 and ART are third-party dependencies of the reference (requirements.txt:13,25) that are
NOT vendored in /root/reference and not installed here.  This file restates the published algorithm of
foolbox 3.3.1 `attacks/gradient_descent_base.py` (BaseGradientDescent.run, Linf/L2 mixins) and
`attacks/projected_gradient_descent.py` / `fast_gradient_method.py`, anchored on the reference's call
sites RobustART/noise/utils/adv/attack.py:20-33 and the in-repo Linf loop
prototype/prototype/solver/adv_cls_solver_train_pgd_new.py:84-103, in plain torch fp32.
PARITY UNPINNED: the reference holds no test or golden vector for these attacks.

MI-FGSM follows RobustART/noise/utils/adv/Attacks/imfgsm_attack.py:62-93 line by line (minus .cuda()).
"""
import torch
import torch.nn.functional as F


def _grad(model_fn, x, y, reduction="sum"):
    x = x.clone().requires_grad_(True)
    loss = F.cross_entropy(model_fn(x), y, reduction=reduction)
    (g,) = torch.autograd.grad(loss, x)
    return g


def pgd_linf(model_fn, x0, y, eps, rel_stepsize, steps, start_u=None, random_start=True):
    """model_fn maps [0,1] images to logits (preprocessing inside, like fb.PyTorchModel)."""
    alpha = rel_stepsize * eps
    if random_start:
        u = torch.rand_like(x0) if start_u is None else start_u
        x = x0 + ((eps - (-eps)) * u + (-eps))          # torch.uniform_: (hi-lo)*u + lo
        x = x.clamp(0, 1)
    else:
        x = x0
    for _ in range(steps):
        g = _grad(model_fn, x, y)
        x = x + alpha * g.sign()
        x = x0 + (x - x0).clamp(-eps, eps)
        x = x.clamp(0, 1)
    return x


def fgsm(model_fn, x0, y, eps):
    return pgd_linf(model_fn, x0, y, eps, 1.0, 1, random_start=False)


def _l2(v):
    return v.flatten(1).norm(dim=1).view(-1, 1, 1, 1)


def pgd_l2(model_fn, x0, y, eps, rel_stepsize, steps, start_direction=None):
    alpha = rel_stepsize * eps
    x = x0 if start_direction is None else (x0 + eps * start_direction).clamp(0, 1)
    for _ in range(steps):
        g = _grad(model_fn, x, y)
        g = g * (1.0 / _l2(g).clamp_min(1e-12))
        x = x + alpha * g
        d = x - x0
        d = d * torch.minimum(torch.ones_like(_l2(d)), eps / _l2(d).clamp_min(1e-12))
        x = (x0 + d).clamp(0, 1)
    return x


def mim_linf(model_norm_fn, X, y, epsilon, num_steps, step_size, decay_factor=1.0, start_u=None,
             mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """model_norm_fn takes NORMALISED input (imfgsm_attack.py:14-23 normalize())."""
    m = torch.tensor(mean, dtype=X.dtype, device=X.device).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=X.dtype, device=X.device).view(1, 3, 1, 1)
    u = torch.rand_like(X) if start_u is None else start_u
    X_pgd = X + ((2 * epsilon) * u - epsilon)
    previous_grad = torch.zeros_like(X)
    for _ in range(num_steps):
        xg = X_pgd.clone().requires_grad_(True)
        loss = F.cross_entropy(model_norm_fn((xg - m) / s), y)      # mean reduction
        (g,) = torch.autograd.grad(loss, xg)
        grad = g / g.abs().mean(dim=[1, 2, 3], keepdim=True)
        previous_grad = decay_factor * previous_grad + grad
        X_pgd = X_pgd + step_size * previous_grad.sign()
        eta = (X_pgd - X).clamp(-epsilon, epsilon)
        X_pgd = (X + eta).clamp(0, 1.0)
    return X_pgd


def pgd_l1_step(x, g, x0, eps_step, eps, tol=1e-7):
    """One step of ART's ProjectedGradientDescentPyTorch(norm=1) (third-party, NOT vendored and not pinned by the reference's
    requirements.txt:25 -- `adversarial-robustness-toolbox`; the statement follows art/attacks/evasion/projected_gradient_descent/
    projected_gradient_descent_pytorch.py of ART 1.7-1.16: _compute_perturbation, _apply_perturbation, _projection).  PARITY
    UNPINNED: ART is absent here; anchored on the call site adv/attack.py:44-49 (norm=1, clip_values (0,1))."""
    n = x.shape[0]
    gn = g / (g.reshape(n, -1).abs().sum(1).view(-1, *([1] * (x.dim() - 1))) + tol)
    x = torch.clamp(x + eps_step * gn, 0.0, 1.0)
    delta = (x - x0).reshape(n, -1)
    delta = delta * torch.clamp(eps / (delta.abs().sum(1, keepdim=True) + tol), max=1.0)
    return x0 + delta.reshape(x.shape)
