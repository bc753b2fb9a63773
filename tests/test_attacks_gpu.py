"""Fused attack-step kernels and attack loops vs the torch fp32 restatement (oracle/attacks.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def _data(cuda, n=6, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x0 = torch.rand(n, 3, 224, 224, generator=g).to(cuda)
    grad = torch.randn(n, 3, 224, 224, generator=g).to(cuda)
    grad[0, 0, :4, :4] = 0.0   # sign(0) = 0 must be honoured
    y = torch.randint(0, 10, (n,), generator=g).to(cuda)
    return x0, grad, y


def test_linf_step_bit_exact(cuda):
    from robustart_b200 import ops
    x0, g, _ = _data(cuda)
    eps, alpha = 4 / 255, 3 / 40 * 4 / 255
    u = torch.rand_like(x0)
    x = ops.random_start_linf(x0, eps, u=u)
    ref = (x0 + ((eps - (-eps)) * u + (-eps))).clamp(0, 1)
    assert torch.equal(x, ref)
    for _ in range(3):
        ref = ref + alpha * g.sign()
        ref = (x0 + (ref - x0).clamp(-eps, eps)).clamp(0, 1)
        ops.pgd_step_linf_(x, g, x0, alpha, eps)
    assert torch.equal(x, ref)
    assert (x - x0).abs().max().item() <= eps + 1e-7
    assert x.min().item() >= 0 and x.max().item() <= 1


def test_random_start_device_rng(cuda):
    from robustart_b200 import ops
    x0 = torch.full((4, 3, 224, 224), 0.5, device=cuda)
    eps = 8 / 255
    a = ops.random_start_linf(x0, eps, seed=1)
    b = ops.random_start_linf(x0, eps, seed=1)
    c = ops.random_start_linf(x0, eps, seed=2)
    assert torch.equal(a, b) and not torch.equal(a, c)
    d = (a - x0)
    assert d.abs().max().item() <= eps + 1e-7
    assert abs(d.mean().item()) < 1e-4 and abs(d.std().item() - eps / 3 ** 0.5) < 1e-4


def test_l2_step(cuda):
    from robustart_b200 import ops
    from oracle import attacks as OA
    x0, g, _ = _data(cuda)
    eps, alpha = 3.0, 3 / 40 * 3.0
    x = x0.clone()
    ref = x0.clone()
    for _ in range(3):
        ops.pgd_step_l2_(x, g, x0, alpha, eps)
        gn = g * (1.0 / OA._l2(g).clamp_min(1e-12))
        ref = ref + alpha * gn
        d = ref - x0
        d = d * torch.minimum(torch.ones_like(OA._l2(d)), eps / OA._l2(d).clamp_min(1e-12))
        ref = (x0 + d).clamp(0, 1)
    assert (x - ref).abs().max().item() < 2e-6
    assert OA._l2(x - x0).max().item() <= eps * (1 + 1e-5)


def test_mim_step(cuda):
    from robustart_b200 import ops
    x0, g, _ = _data(cuda)
    eps, step, decay = 8 / 255, 0.002, 1.0
    x, m = x0.clone(), torch.zeros_like(x0)
    rx, rm = x0.clone(), torch.zeros_like(x0)
    for it in range(3):
        gi = g * (it + 1)
        ops.mim_step_linf_(x, m, gi, x0, step, eps, decay)
        grad = gi / gi.abs().mean(dim=[1, 2, 3], keepdim=True)
        rm = decay * rm + grad
        rx = rx + step * rm.sign()
        rx = (x0 + (rx - x0).clamp(-eps, eps)).clamp(0, 1)
    assert (m - rm).abs().max().item() < 1e-4 * rm.abs().max().item()
    mism = (x != rx).float().mean().item()
    assert mism < 1e-4, mism


def _mlp(cuda, seed=0):
    torch.manual_seed(seed)
    net = torch.nn.Sequential(torch.nn.AdaptiveAvgPool2d(8), torch.nn.Flatten(), torch.nn.Linear(192, 64),
                              torch.nn.Tanh(), torch.nn.Linear(64, 10)).to(cuda).eval()
    return net


@pytest.mark.parametrize("attack", ["pgd_linf", "fgsm", "pgd_l2", "mim_linf"])
def test_attack_loops_match_restatement(cuda, attack):
    from robustart_b200 import attacks as A
    from oracle import attacks as OA
    net = _mlp(cuda)
    x0, _, y = _data(cuda, n=8, seed=3)
    m = torch.tensor(MEAN, device=cuda).view(1, 3, 1, 1)
    s = torch.tensor(STD, device=cuda).view(1, 3, 1, 1)
    fmodel = A.PyTorchModel(net, bounds=(0, 1), preprocessing=dict(mean=MEAN, std=STD, axis=-3))
    ref_fn = lambda x: net((x - m) / s)
    u = torch.rand_like(x0)
    if attack == "pgd_linf":
        eps = 4 / 255
        got = A.pgd_linf(x0, y, fmodel, eps, 3 / 40, 10, start_uniform=u)
        want = OA.pgd_linf(ref_fn, x0, y, eps, 3 / 40, 10, start_u=u)
        tol_frac, bound = 5e-3, eps
    elif attack == "fgsm":
        eps = 8 / 255
        got = A.fgsm(x0, y, fmodel, eps)
        want = OA.fgsm(ref_fn, x0, y, eps)
        tol_frac, bound = 1e-3, eps
    elif attack == "pgd_l2":
        eps = 2.0
        z = torch.randn(8, 3 * 224 * 224 + 1, device=cuda)
        dirn = (z / z.norm(dim=1, keepdim=True))[:, :-1].reshape(x0.shape)
        got = A.pgd_l2(x0, y, fmodel, eps, 3 / 40, 10, start_direction=dirn)
        want = OA.pgd_l2(ref_fn, x0, y, eps, 3 / 40, 10, start_direction=dirn)
        assert (got - want).abs().max().item() < 1e-4
        assert OA._l2(got - x0).max().item() <= eps * (1 + 1e-5)
        return
    else:
        eps = 8 / 255
        got = A.mim_linf(x0, y, net, eps, 10, 0.002, 1.0, start_uniform=u)
        want = OA.mim_linf(net, x0, y, eps, 10, 0.002, 1.0, start_u=u)
        tol_frac, bound = 5e-3, eps
    assert (got - x0).abs().max().item() <= bound + 1e-6
    assert got.min().item() >= 0 and got.max().item() <= 1
    # a sign flip of a ~0 gradient moves a coordinate by 2*alpha: allow a tiny fraction of those
    mism = ((got - want).abs() > 1e-6).float().mean().item()
    assert mism < tol_frac, mism
    # attack quality agrees: same predictions on the adversarials
    assert torch.equal(ref_fn(got).argmax(1), ref_fn(want).argmax(1))


def test_addnoise_adv_surface(cuda):
    """The plugin call the solvers make (benchmark_eval_adv.py:198-209,231)."""
    from RobustART.noise import AddNoise
    from robustart_b200.attacks import PyTorchModel
    net = _mlp(cuda)
    x0, _, y = _data(cuda, n=4, seed=5)
    fmodel = PyTorchModel(net, bounds=(0, 1), preprocessing=dict(mean=MEAN, std=STD, axis=-3))
    gen = AddNoise("pgd_linf")
    gen.set_config(f_model=fmodel, eps=4 / 255, steps=5)
    adv = gen.add_noise(x0, y)
    assert adv.shape == x0.shape and adv.is_cuda and adv.dtype == torch.float32
    assert (adv - x0).abs().max().item() <= 4 / 255 + 1e-6
    with pytest.raises(AssertionError):
        gen.set_config(bogus=1)
    # defaults are not shared between instances (the reference aliases the module-level dict)
    assert AddNoise("pgd_linf").config["steps"] == 20


def test_pgd_l1_step_and_loop(cuda):
    """PGD-L1 (attack.py:44-49 -> ART ProjectedGradientDescentPyTorch(norm=1); ART is not vendored: PARITY UNPINNED, the update
    rules are stated in oracle.attacks.pgd_l1_step): the fused step kernel against that statement, then the loop through the
    plugin -- same start, same adversarials as the stated loop on the autograd model; ||delta||_1 <= eps; inside the box."""
    from oracle import attacks as OA
    from RobustART.noise import AddNoise
    from robustart_b200 import attacks, ops
    x0, grad, y = _data(cuda, n=5, seed=9)
    x = (x0 + 0.01 * torch.randn_like(x0)).clamp(0, 1).contiguous()
    for eps_step, eps in ((120.0, 1600.0), (500.0, 300.0)):
        want = OA.pgd_l1_step(x.cpu().double(), grad.cpu().double(), x0.cpu().double(), eps_step, eps)
        got = ops.pgd_step_l1_(x.clone(), grad, x0, eps_step, eps)
        assert (got.cpu().double() - want).abs().max().item() < 2e-6
        assert (got - x0).reshape(5, -1).abs().sum(1).max().item() <= eps * (1 + 1e-5)
    net = _mlp(cuda)
    gen = torch.Generator().manual_seed(3)
    start = (torch.randn(5, 3, 224, 224, generator=gen) * 0.004).to(cuda)
    adv = attacks.pgd_l1(x0, y, net, 1600.0, 224, 120.0, 6, 16, start=start)
    # the stated loop with autograd gradients of the mean CE loss (what ART's PyTorchClassifier.loss_gradient returns)
    xs = (x0 + start).clamp(0, 1)
    mean = torch.tensor(MEAN, device=cuda).view(1, 3, 1, 1)
    std = torch.tensor(STD, device=cuda).view(1, 3, 1, 1)
    for _ in range(6):
        xs = xs.detach().requires_grad_(True)
        loss = torch.nn.functional.cross_entropy(net((xs - mean) / std), y)
        (g,) = torch.autograd.grad(loss, xs)
        xs = OA.pgd_l1_step(xs.detach(), g, x0, 120.0, 1600.0)
    assert (adv - xs).abs().max().item() < 1e-5
    assert (adv - x0).reshape(5, -1).abs().sum(1).max().item() <= 1600.0 * (1 + 1e-5) and adv.min().item() >= 0 and adv.max().item() <= 1
    plug = AddNoise("pgd_l1")
    plug.set_config(model=net, eps=800.0, max_iter=3)
    out = plug.add_noise(x0, y)
    assert out.shape == x0.shape and (out - x0).reshape(5, -1).abs().sum(1).max().item() <= 800.0 * (1 + 1e-5)
    assert (out != x0).float().mean().item() > 0.5


def test_pgd_graph_replay_equals_eager(cuda):
    """A NativeModel's attack step is captured into a CUDA graph (forward_saved + CE gradient + input_grad + step kernel) and
    replayed; the adversarials must be bit-identical to the eager launch sequence, also when the cached graph is reused on new data."""
    from robustart_b200 import attacks, nets
    net = nets.build_model("resnet18", device=cuda, passes=3)
    g = torch.Generator().manual_seed(12)
    graphed, eager = attacks.NativeModel(net, use_graphs=True), attacks.NativeModel(net, use_graphs=False)
    for _ in range(2):
        x = torch.rand(3, 3, 224, 224, generator=g).to(cuda)
        y = torch.randint(0, 1000, (3,), generator=g).to(cuda)
        u = torch.rand(3, 3, 224, 224, generator=g).to(cuda)
        a = attacks.pgd_linf(x, y, graphed, 4 / 255, 3 / 40, 4, start_uniform=u)
        b = attacks.pgd_linf(x, y, eager, 4 / 255, 3 / 40, 4, start_uniform=u)
        assert torch.equal(a, b)
        assert (a - x).abs().max().item() <= 4 / 255 + 1e-6
    assert len(graphed._step_graphs) == 1


def test_random_start_l2_uniform_in_ball(cuda):
    """b200r_random_start_l2 (foolbox uniform_l2_n_balls): points uniform in the eps-ball -- the radius of a uniform point in a
    D-ball concentrates at eps * (1 - 1/D), the direction is isotropic, seeds / image offsets give reproducible streams."""
    from robustart_b200 import ops
    n, d, eps = 64, 3 * 32 * 32, 2.5
    x0 = torch.full((n, 3, 32, 32), 0.5, device=cuda)            # clipping to [0, 1] cannot touch |delta| <= 2.5 / sqrt(3072) per pixel scale
    x = ops.random_start_l2(x0, eps, seed=7)
    delta = (x - x0).view(n, -1).double()
    r = delta.norm(dim=1)
    assert (r <= eps * (1 + 1e-5)).all()
    # radius of a uniform point in the unit D-ball: U^(1/D); its mean is D / (D + 1)
    assert abs(r.mean().item() / eps - d / (d + 1.0)) < 2e-4
    # isotropy: coordinates have mean 0 and variance eps^2 / (D + 2)
    assert abs(delta.mean().item()) < 4 * eps / (d ** 0.5) / (n * d) ** 0.5 * 1.5
    assert abs(delta.var().item() * (d + 2) / eps ** 2 - 1.0) < 0.02
    # distinct samples are uncorrelated; the same (seed, offset) reproduces, a re-batched call continues the streams
    c = (delta[0] * delta[1]).sum().item() / (r[0] * r[1]).item()
    assert abs(c) < 0.1
    assert torch.equal(x, ops.random_start_l2(x0, eps, seed=7))
    assert not torch.equal(x, ops.random_start_l2(x0, eps, seed=8))
    assert torch.equal(x[16:], ops.random_start_l2(x0[16:], eps, seed=7, image_offset=16))
    # clipping
    y = ops.random_start_l2(torch.zeros(4, 3, 8, 8, device=cuda), 50.0, seed=1)
    assert y.min().item() == 0.0 and y.max().item() <= 1.0
    # normality of the direction: kurtosis of the coordinates of one long sample ~ 3
    z = ops.random_start_l2(torch.full((1, 3, 224, 224), 0.5, device=cuda), 1.0, seed=3).view(-1).double() - 0.5
    k = ((z - z.mean()) ** 4).mean() / z.var() ** 2
    assert abs(k.item() - 3.0) < 0.1


@pytest.mark.parametrize("n,c,h,w", [(5, 3, 224, 224), (3, 3, 100, 100), (2, 1, 8, 8), (2, 3, 256, 256)])
def test_cluster_step_kernels_match_the_multi_kernel_form(cuda, n, c, h, w, monkeypatch):
    """PGD-L2 and MI-FGSM steps as ONE cluster launch per step (8 CTAs per image, the image in registers, the per-image norm reduced
    through distributed shared memory in fixed order) against the multi-kernel form with atomics (B200R_L2_STEP=3k / B200R_MIM_STEP=2k;
    also what an image too large for a cluster's registers takes: 3 x 256 x 256): same updates, and bit-identical from run to run."""
    from robustart_b200 import ops
    from oracle import attacks as OA
    torch.manual_seed(n + h)
    x0 = torch.rand(n, c, h, w, device=cuda)
    g = torch.randn(n, c, h, w, device=cuda) * torch.rand(n, 1, 1, 1, device=cuda)
    eps, alpha = 2.0, 0.4
    fits = c * h * w // 4 <= 8 * 512 * 10               # else both settings run the multi-kernel form (atomics: not bit-reproducible)
    outs = {}
    for mode in ("cluster", "cluster", "3k"):
        monkeypatch.setenv("B200R_L2_STEP", mode)
        x = x0.clone()
        for _ in range(3):
            ops.pgd_step_l2_(x, g, x0, alpha, eps)
        outs.setdefault(mode, []).append(x)
    assert not fits or torch.equal(outs["cluster"][0], outs["cluster"][1])
    assert (outs["cluster"][0] - outs["3k"][0]).abs().max().item() < 2e-6
    assert OA._l2(outs["cluster"][0] - x0).max().item() <= eps * (1 + 1e-5)
    step, decay = 0.002, 1.0
    outs = {}
    for mode in ("cluster", "cluster", "2k"):
        monkeypatch.setenv("B200R_MIM_STEP", mode)
        x, m = x0.clone(), torch.zeros_like(x0)
        for it in range(3):
            ops.mim_step_linf_(x, m, g * (it + 1), x0, step, 8 / 255, decay)
        outs.setdefault(mode, []).append((x, m))
    assert not fits or (torch.equal(outs["cluster"][0][0], outs["cluster"][1][0]) and torch.equal(outs["cluster"][0][1], outs["cluster"][1][1]))
    ma, mb = outs["cluster"][0][1], outs["2k"][0][1]
    assert (ma - mb).abs().max().item() < 1e-5 * mb.abs().max().item()
    assert (outs["cluster"][0][0] != outs["2k"][0][0]).float().mean().item() < 1e-4
