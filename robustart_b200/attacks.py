"""Gradient attacks of RobustART.noise on B200: same call signatures as
RobustART/noise/utils/adv/attack.py:20-42, same update rules as foolbox 3.3.1 (third-party, pinned in
the reference's requirements.txt:13) and the vendored MI-FGSM (Attacks/imfgsm_attack.py:62-93).

The model is whatever the caller passes.  A `NativeModel` (a robustart_b200.nets network) runs forward AND
input gradient on the sm_100a kernels; an arbitrary nn.Module (the reference hands those in) goes through
torch.autograd.  Everything around it -- random start, softmax-CE
gradient, sign/normalise/step/project/clip, (de)normalisation -- is one fused sm_100a kernel each
(robustart_b200/csrc/attack_steps.cu, loss_metrics.cu).  Unlike eagerpy's loss.backward() only the
input gradient is requested, so no weight gradients are computed.
"""
from __future__ import annotations

import itertools
import os
from typing import Optional

import torch

from . import ops

_call_counter = itertools.count()


class _NormalizeFn(torch.autograd.Function):
    """(x - mean)/std with both directions on our kernels (b200r_normalize_f32nchw)."""

    @staticmethod
    def forward(ctx, x, mean, std):
        ctx.ms = (mean, std)
        return ops.normalize(x.contiguous(), "normal", mean, std)

    @staticmethod
    def backward(ctx, g):
        mean, std = ctx.ms
        return ops.normalize(g.contiguous(), "grad", mean, std), None, None


class PyTorchModel:
    """Minimal stand-in for foolbox.PyTorchModel as the solvers build it
    (benchmark_eval_adv.py:198-207): inputs in `bounds`, `preprocessing` = dict(mean, std, axis=-3).
    A real foolbox model object works too: the attacks only need `model(x) -> logits` and `.bounds`."""

    def __init__(self, model, bounds=(0, 1), device=None, preprocessing=None):
        self.model = model
        self.bounds = tuple(bounds)
        self.device = device
        self.preprocessing = preprocessing
        if preprocessing is not None:
            assert preprocessing.get("axis", -3) == -3, "only channel-first preprocessing is supported"
            self._mean = tuple(float(v) for v in preprocessing["mean"])
            self._std = tuple(float(v) for v in preprocessing["std"])

    def __call__(self, x):
        if self.preprocessing is not None:
            x = _NormalizeFn.apply(x, self._mean, self._std)
        return self.model(x)

    def forward_vjp(self, x):
        return _autograd_forward_vjp(self, x)


class NativeModel:
    """A robustart_b200.nets model as the SOURCE model of an attack: forward and input gradient both run on the
    sm_100a kernels (nets.ResNet.forward_saved / input_grad) -- no autograd graph, no weight gradients.  Takes
    [0,1] float32 NCHW like a foolbox model with ImageNet preprocessing (Normalize is fused into the stem), so it
    can be handed to every attack in this package, including the ones whose reference signature expects a raw
    nn.Module on normalised input (mim_linf, autoattack_linf)."""

    bounds = (0, 1)

    def __init__(self, net, passes_bwd: Optional[int] = None, use_graphs: Optional[bool] = None):
        if not hasattr(net, "forward_saved"):
            raise NotImplementedError("%s has no native input-gradient pass" % type(net).__name__)
        self.net, self.passes_bwd = net, passes_bwd
        # one whole attack step (forward_saved + loss gradient + input_grad + step kernel) is captured into a CUDA graph per
        # (attack, batch shape) and replayed: ~130 launches of 5-100 us each are launch-bound from Python (B200R_ATTACK_GRAPHS=0: eager)
        self.use_graphs = (os.environ.get("B200R_ATTACK_GRAPHS", "1") != "0") if use_graphs is None else use_graphs
        self._step_graphs = {}

    def graphed_step(self, key, x0, y, build):
        """Static buffers + captured graph of one attack step for this (attack kind, shape, hyper-parameters).
        build(st) launches the step on the static tensors st.x / st.x0 / st.y (+ whatever it adds to st).  Returns st with
        st.replay()."""
        st = self._step_graphs.get(key)
        if st is None:
            st = _StepGraph(x0, y)
            side = torch.cuda.Stream(device=x0.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):                      # warm-up: lazy weight layouts, function attributes, workspaces, allocator
                    build(st)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            st.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(st.graph):
                build(st)
            self._step_graphs[key] = st
        return st

    def __call__(self, x):
        return self.net.forward(x.contiguous())

    def forward_vjp(self, x):
        logits, saved = self.net.forward_saved(x.detach().contiguous())
        return logits, lambda d: self.net.input_grad(d.float().contiguous(), saved, self.passes_bwd)


class _StepGraph:
    def __init__(self, x0, y):
        self.x, self.x0, self.y = torch.empty_like(x0), torch.empty_like(x0), torch.empty_like(y)
        self.graph = None

    def replay(self):
        self.graph.replay()


def _autograd_forward_vjp(f_model, x):
    x = x.detach().requires_grad_(True)
    with torch.enable_grad():
        logits = f_model(x)
    if hasattr(logits, "raw"):  # eagerpy tensor from a real foolbox model
        logits = logits.raw

    def vjp(d):
        (g,) = torch.autograd.grad(logits, x, grad_outputs=d.to(logits.dtype))
        return g.contiguous()
    return logits.detach().float().contiguous(), vjp


def forward_vjp(f_model, x):
    """(logits, fn: dlogits -> dx) for any model object the attacks accept."""
    fn = getattr(f_model, "forward_vjp", None)
    return fn(x) if fn is not None else _autograd_forward_vjp(f_model, x)


def as_f_model(model):
    """mim / autoattack receive a raw module on NORMALISED input (benchmark_eval_adv.py:201-203)."""
    if isinstance(model, NativeModel):
        return model
    return PyTorchModel(model, preprocessing=dict(mean=ops.IMAGENET_MEAN, std=ops.IMAGENET_STD, axis=-3))


def _bounds01(f_model):
    b = getattr(f_model, "bounds", (0, 1))
    lo, hi = float(b[0]), float(b[1])
    if (lo, hi) != (0.0, 1.0):
        raise NotImplementedError("only model bounds (0, 1) are supported, got %r" % (b,))


def _input_grad(f_model, x, label, grad_scale=1.0):
    """d/dx sum_i CE(f(x)_i, y_i) * grad_scale, using our CE kernel for dL/dlogits."""
    logits, vjp = forward_vjp(f_model, x)
    _, dlogits = ops.ce_loss_grad(logits, label, grad_scale)
    return vjp(dlogits)


def _prep(input, label):
    if not input.is_cuda:
        raise TypeError("adversarial noise needs CUDA tensors (the reference is GPU-only here too)")
    x0 = input.detach().to(torch.float32).contiguous()
    y = label.detach().to(torch.int64).view(-1).contiguous()
    return x0, y


def pgd_linf(input, label, f_model, eps, rel_stepsize, steps, *, random_start=True, seed: Optional[int] = None,
             start_uniform: Optional[torch.Tensor] = None):
    """foolbox LinfProjectedGradientDescentAttack(rel_stepsize, steps)(f_model, input, label, epsilons=eps)[0]
    (attack.py:20-23): raw adversarials."""
    _bounds01(f_model)
    x0, y = _prep(input, label)
    alpha = float(rel_stepsize) * float(eps)
    if random_start:
        s = next(_call_counter) if seed is None else seed
        x = ops.random_start_linf(x0, float(eps), seed=s, u=start_uniform, clip01=True)
    else:
        x = x0.clone()
    if isinstance(f_model, NativeModel) and f_model.use_graphs and int(steps) > 1:
        def build(st):
            ops.pgd_step_linf_(st.x, _input_grad(f_model, st.x, st.y), st.x0, alpha, float(eps))
        st = f_model.graphed_step(("pgd_linf", tuple(x0.shape), alpha, float(eps)), x0, y, build)
        st.x0.copy_(x0); st.y.copy_(y); st.x.copy_(x)
        for _ in range(int(steps)):
            st.replay()
        return st.x.clone()
    for _ in range(int(steps)):
        g = _input_grad(f_model, x, y)
        ops.pgd_step_linf_(x, g, x0, alpha, float(eps))
    return x


def fgsm(input, label, f_model, eps):
    """foolbox LinfFastGradientAttack(): one step of size eps, no random start (attack.py:30-33)."""
    return pgd_linf(input, label, f_model, eps, rel_stepsize=1.0, steps=1, random_start=False)


def pgd_l2(input, label, f_model, eps, rel_stepsize, steps, *, random_start=True, seed: Optional[int] = None,
           start_direction: Optional[torch.Tensor] = None):
    """foolbox L2ProjectedGradientDescentAttack (attack.py:25-28).  Random start: a uniform draw from
    the eps-ball (foolbox uniform_l2_n_balls: the first n coordinates of a uniform point on the
    (n+1)-sphere), drawn on the device by b200r_random_start_l2; `start_direction` (tests) overrides it."""
    _bounds01(f_model)
    x0, y = _prep(input, label)
    alpha = float(rel_stepsize) * float(eps)
    n = x0.shape[0]
    if random_start:
        if start_direction is None:
            x = ops.random_start_l2(x0, float(eps), seed=next(_call_counter) if seed is None else int(seed))
        else:
            x = (x0 + float(eps) * start_direction).clamp_(0, 1).contiguous()
    else:
        x = x0.clone()
    for _ in range(int(steps)):
        g = _input_grad(f_model, x, y)
        ops.pgd_step_l2_(x, g, x0, alpha, float(eps))
    return x


def pgd_l1(input, label, model, eps, input_size=None, eps_step=120.0, max_iter=20, batch_size=16, *, seed: Optional[int] = None,
           start: Optional[torch.Tensor] = None):
    """attack.py:44-49: ART's ProjectedGradientDescentPyTorch(norm=1, eps, eps_step, max_iter, num_random_init=1, batch_size) on a
    PyTorchClassifier with clip_values (0, 1) and ImageNet preprocessing -- `model` takes NORMALISED input, like mim / autoattack.
    ART is not vendored by the reference and not version-pinned: the update rules follow ART 1.7-1.16 (oracle.attacks.pgd_l1_step
    states them; PARITY UNPINNED).  What ART does per internal batch of `batch_size` (16) samples on the host -- numpy round trip
    included -- runs here on the whole device batch: the loss is a mean over the batch, but the L1-normalised step makes the
    update invariant to that scale, so the result does not depend on the batching.
      start: x_adv = clip(x + r), r = R * (Dirichlet(1..1) spacings) * random signs, R = sqrt(U(0, eps^2))   (ART random_sphere, norm 1)
      step:  g /= ||g||_1 + 1e-7; x = clip(x + eps_step g); delta = (x - x0) * min(1, eps / (||x - x0||_1 + 1e-7)); x = x0 + delta
    `start` overrides the random perturbation r (tests)."""
    x0, y = _prep(input, label)
    n = x0.shape[0]
    f_model = as_f_model(model)
    if start is None:
        gen = None
        if seed is not None:
            gen = torch.Generator(device=x0.device)
            gen.manual_seed(seed)
        dim = x0[0].numel()
        e = torch.empty(n, dim, device=x0.device).exponential_(generator=gen)      # spacings of sorted uniforms = normalised exponentials
        r = torch.sqrt(torch.rand(n, 1, device=x0.device, generator=gen) * float(eps) ** 2)
        sgn = torch.randint(0, 2, (n, dim), device=x0.device, generator=gen).float() * 2 - 1
        start = (e / e.sum(1, keepdim=True) * r * sgn).reshape(x0.shape)
    x = (x0 + start).clamp_(0, 1).contiguous()
    for _ in range(int(max_iter)):
        g = _input_grad(f_model, x, y, grad_scale=1.0 / n)
        ops.pgd_step_l1_(x, g, x0, float(eps_step), float(eps))
    return x


def mim_linf(input, label, model, eps, num_steps, step_size, decay_factor, *, seed: Optional[int] = None,
             start_uniform: Optional[torch.Tensor] = None):
    """_mim_whitebox (imfgsm_attack.py:62-93): `model` takes NORMALISED input; CE *mean* loss; the random
    start is not clipped to [0,1]; the two diagnostic forwards (err, err_pgd) are not needed for the
    result and are skipped."""
    x0, y = _prep(input, label)
    n = x0.shape[0]
    s = next(_call_counter) if seed is None else seed
    x = ops.random_start_linf(x0, float(eps), seed=s, u=start_uniform, clip01=False)
    momentum = torch.zeros_like(x0)
    f_model = as_f_model(model)
    for _ in range(int(num_steps)):
        g = _input_grad(f_model, x, y, grad_scale=1.0 / n)
        ops.mim_step_linf_(x, momentum, g, x0, float(step_size), float(eps), float(decay_factor))
    return x
