"""AutoAttack (Linf, version 'standard') on B200 -- same call surface as the reference's wrapper
`autoattack_linf(input, label, model, norm, eps, version, verbose)` (RobustART/noise/utils/adv/attack.py:35-38)
and the same control flow as the vendored fra31/auto-attack code it calls:

  AutoAttack.run_standard_evaluation  .../Attacks/autoattack/autoattack.py:90-211   (robust_flags, attack order)
  APGDAttack / APGDAttack_targeted    .../autopgd_base.py:208-448, 450-529, 610-690 (checkpoints, step halving)
  SquareAttack (Linf)                 .../square.py:221-294, 192-219, 68-86
  FABAttack (targeted, Linf)          .../fab_base.py:84-270, fab_pt.py:102-117, fab_projections.py:7-59

The model is the caller's nn.Module taking NORMALISED input (NormalizeModel, autoattack.py:17-23); its forward /
input-gradient run through torch.autograd.  Device kernels of libb200robust do the elementwise hot spots: the
APGD double-projection update (b200r_apgd_step_linf), CE / DLR losses with their logit gradients
(b200r_ce_loss_grad, b200r_dlr_loss_grad), the Square proposal + masked accept, and (de)normalisation.
FAB's hyperplane projection (sort + cumsum + bisection over 150 528 coordinates) uses torch.sort / cumsum for now.
Differences, stated: seeds come from a counter instead of time.time() (reference: attack.py:36 seed=None ->
non-reproducible); L2/L1 norms and the 'plus'/'rand' versions are not implemented.
"""
from __future__ import annotations

import ctypes as C
import itertools
import math

import torch

from . import _lib, ops
from .attacks import PyTorchModel, as_f_model, forward_vjp  # noqa: F401

_seed_counter = itertools.count(1)


def _stream():
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------------------
# thin wrappers over the C-ABI
# ------------------------------------------------------------------------------------------------
def _apgd_step_(x_adv, x_old, grad, x0, step, eps, a):
    _lib.check(_lib.load().b200r_apgd_step_linf(x_adv.data_ptr(), x_old.data_ptr(), grad.data_ptr(), x0.data_ptr(), step.data_ptr(),
                                                x_adv.shape[0], x_adv[0].numel(), eps, a, _stream()))


def _dlr(logits, y, target=None, want_grad=True):
    n, k = logits.shape
    loss = torch.empty(n, dtype=torch.float32, device=logits.device)
    d = torch.empty_like(logits) if want_grad else None
    _lib.check(_lib.load().b200r_dlr_loss_grad(logits.data_ptr(), y.data_ptr(), None if target is None else target.data_ptr(),
                                               loss.data_ptr(), None if d is None else d.data_ptr(), n, k, _stream()))
    return loss, d


def _masked_rows_(dst, src, mask):
    m = mask.to(torch.uint8).contiguous()
    _lib.check(_lib.load().b200r_masked_rows_copy(dst.data_ptr(), src.data_ptr(), m.data_ptr(), dst.shape[0], dst[0].numel(), _stream()))


class _Model:
    """NormalizeModel (autoattack.py:17-23): inputs in [0,1], normalisation on our kernel, logits float32."""

    def __init__(self, model):
        self.f = as_f_model(model)
        self.forwards = 0
        self.backwards = 0

    def logits(self, x):
        self.forwards += x.shape[0]
        with torch.no_grad():
            return self.f(x.contiguous()).float().contiguous()

    def loss_and_grad(self, x, y, loss_kind, target=None):
        """returns (logits, per-sample loss, d(sum loss)/dx)"""
        self.forwards += x.shape[0]
        self.backwards += x.shape[0]
        lg, vjp = forward_vjp(self.f, x.detach().contiguous())
        if loss_kind == "ce":
            loss, d = ops.ce_loss_grad(lg, y)
        elif loss_kind == "dlr":
            loss, d = _dlr(lg, y)
        elif loss_kind == "dlr-targeted":
            loss, d = _dlr(lg, y, target)
        elif loss_kind == "fab-diff":          # -(z_y - z_t): fab_pt.py:102-117
            u = torch.arange(lg.shape[0], device=lg.device)
            loss = -(lg[u, y] - lg[u, target])
            d = torch.zeros_like(lg)
            d[u, y] = -1.0
            d[u, target] = 1.0
        else:
            raise ValueError(loss_kind)
        return lg, loss, vjp(d)


# ------------------------------------------------------------------------------------------------
# APGD (Linf)
# ------------------------------------------------------------------------------------------------
class APGD:
    def __init__(self, model: _Model, eps, n_iter=100, n_restarts=1, rho=0.75, loss="ce", n_target_classes=9):
        self.m, self.eps, self.n_iter, self.n_restarts, self.rho = model, float(eps), n_iter, n_restarts, rho
        self.loss, self.n_target_classes = loss, n_target_classes
        self.n_iter_2 = max(int(0.22 * n_iter), 1)      # autopgd_base.py:163-165
        self.n_iter_min = max(int(0.06 * n_iter), 1)
        self.size_decr = max(int(0.03 * n_iter), 1)

    def _check_oscillation(self, loss_steps, j, k, k3):
        t = torch.zeros(loss_steps.shape[1], device=loss_steps.device)
        for c in range(k):
            t += (loss_steps[j - c] > loss_steps[j - c - 1]).float()
        return (t <= k * k3).float()

    def single_run(self, x, y, target=None):
        """autopgd_base.py:208-448 (Linf branch). Returns (x_best, acc, loss_best, x_best_adv)."""
        eps, n = self.eps, x.shape[0]
        t = 2 * torch.rand_like(x) - 1
        x_adv = x + eps * t / (t.abs().flatten(1).max(1)[0].view(-1, 1, 1, 1) + 1e-12)
        x_adv = x_adv.clamp(0.0, 1.0).contiguous()
        x_best, x_best_adv = x_adv.clone(), x_adv.clone()
        loss_steps = torch.zeros(self.n_iter, n, device=x.device)
        logits, loss_indiv, grad = self.m.loss_and_grad(x_adv, y, self.loss, target)
        grad_best = grad.clone()
        acc = logits.max(1)[1] == y
        loss_best = loss_indiv.clone()
        step_size = torch.full((n,), 2.0 * eps, device=x.device)
        x_adv_old = x_adv.clone()
        k, counter3 = self.n_iter_2, 0
        loss_best_last_check = loss_best.clone()
        reduced_last_check = torch.ones_like(loss_best)
        for i in range(self.n_iter):
            _apgd_step_(x_adv, x_adv_old, grad, x, step_size, eps, 0.75 if i > 0 else 1.0)   # fused update
            logits, loss_indiv, grad = self.m.loss_and_grad(x_adv, y, self.loss, target)
            pred = logits.max(1)[1] == y
            acc = torch.min(acc, pred)
            _masked_rows_(x_best_adv, x_adv, ~pred)
            improved = loss_indiv > loss_best
            _masked_rows_(x_best, x_adv, improved)
            _masked_rows_(grad_best, grad, improved)
            loss_best = torch.where(improved, loss_indiv, loss_best)
            loss_steps[i] = loss_indiv
            counter3 += 1
            if counter3 == k:
                fl_osc = self._check_oscillation(loss_steps, i, k, self.rho)
                fl_no_impr = (1.0 - reduced_last_check) * (loss_best_last_check >= loss_best).float()
                fl_osc = torch.max(fl_osc, fl_no_impr)
                reduced_last_check = fl_osc.clone()
                loss_best_last_check = loss_best.clone()
                red = fl_osc > 0
                step_size = torch.where(red, step_size / 2.0, step_size)
                _masked_rows_(x_adv, x_best, red)            # restart from the best point
                _masked_rows_(grad, grad_best, red)
                k = max(k - self.size_decr, self.n_iter_min)
                counter3 = 0
        return x_best, acc, loss_best, x_best_adv

    def perturb(self, x, y, seed):
        """autopgd_base.py:450-529 (untargeted) / :610-690 (targeted over the 2nd..10th most likely classes)."""
        x = x.detach().float().contiguous()
        y = y.detach().long().contiguous()
        torch.manual_seed(seed)
        torch.cuda.manual_seed(seed)
        adv = x.clone()
        acc = self.m.logits(x).max(1)[1] == y
        targets = [None] if self.loss != "dlr-targeted" else list(range(2, self.n_target_classes + 2))
        for tc in targets:
            for _ in range(self.n_restarts):
                idx = acc.nonzero().flatten()
                if idx.numel() == 0:
                    continue
                xs, ys = x[idx].contiguous(), y[idx].contiguous()
                tgt = None
                if tc is not None:
                    tgt = self.m.logits(xs).sort(dim=1)[1][:, -tc].contiguous()
                _, acc_curr, _, adv_curr = self.single_run(xs, ys, tgt)
                fooled = ~acc_curr
                acc[idx[fooled]] = False
                adv[idx[fooled]] = adv_curr[fooled]
        return adv


# ------------------------------------------------------------------------------------------------
# Square (Linf)
# ------------------------------------------------------------------------------------------------
class Square:
    def __init__(self, model: _Model, eps, n_queries=5000, p_init=0.8):
        self.m, self.eps, self.n_queries, self.p_init = model, float(eps), n_queries, p_init

    def _p(self, it):  # square.py:192-219 (resc_schedule False)
        for hi, div in ((10, 1), (50, 2), (200, 4), (500, 8), (1000, 16), (2000, 32), (4000, 64), (6000, 128), (8000, 256)):
            if it <= hi:
                return self.p_init / div
        return self.p_init / 512

    def _margin_loss(self, x, y):  # square.py:68-86 (untargeted, loss='margin' as AutoAttack configures it)
        logits = self.m.logits(x)
        u = torch.arange(x.shape[0], device=x.device)
        y_corr = logits[u, y].clone()
        logits[u, y] = -float("inf")
        margin = y_corr - logits.max(dim=-1)[0]
        return margin, margin

    def perturb(self, x, y, seed):
        x = x.detach().float().contiguous()
        y = y.detach().long().contiguous()
        torch.manual_seed(seed)
        torch.cuda.manual_seed(seed)
        adv = x.clone()
        acc = self.m.logits(x).max(1)[1] == y
        idx = acc.nonzero().flatten()
        if idx.numel() == 0:
            return adv
        xs, ys = x[idx].contiguous(), y[idx].contiguous()
        n, c, h, w = xs.shape
        eps = self.eps
        # vertical stripes initialisation (square.py:229-231)
        x_best = (xs + eps * torch.sign(2 * torch.rand(n, c, 1, w, device=x.device) - 1)).clamp(0.0, 1.0).contiguous()
        margin_min, loss_min = self._margin_loss(x_best, ys)
        lib = _lib.load()
        for it in range(self.n_queries):
            act = (margin_min > 0.0).nonzero().flatten()
            if act.numel() == 0:
                break
            xc, xb, yc = xs[act].contiguous(), x_best[act].contiguous(), ys[act]
            s = max(int(round(math.sqrt(self._p(it) * c * h * w / c))), 1)
            vh = int((torch.rand(1).item()) * (h - s))           # random_int(0, h - s): low + (high-low)*rand, truncated
            vw = int((torch.rand(1).item()) * (w - s))
            signs = torch.sign(2 * torch.rand(c) - 1).tolist()
            x_new = torch.empty_like(xb)
            _lib.check(lib.b200r_square_propose_linf(xb.data_ptr(), xc.data_ptr(), x_new.data_ptr(), xb.shape[0], c, h, w, vh, vw, s,
                                                     (C.c_float * 3)(*signs), eps, _stream()))
            margin, loss = self._margin_loss(x_new, yc)
            improved = loss < loss_min[act]
            loss_min[act] = torch.where(improved, loss, loss_min[act])
            improved = improved | (margin <= 0.0)
            margin_min[act] = torch.where(improved, margin, margin_min[act])
            _masked_rows_(xb, x_new, improved)
            x_best[act] = xb
        acc_curr = self.m.logits(x_best).max(1)[1] == ys
        fooled = ~acc_curr
        adv[idx[fooled]] = x_best[fooled]
        return adv


# ------------------------------------------------------------------------------------------------
# FAB-T (Linf)
# ------------------------------------------------------------------------------------------------
def projection_linf(t, w, b):
    """fab_projections.py:7-59: project the rows of t onto {x: <w,x> = b} intersected with the box [0,1]^d, minimising the Linf
    norm of the step.  One sort-free kernel launch (csrc/attack_proj.cu: per-row Newton on the threshold); the reference's
    argsort / cumsum / index-bisection statement lives in oracle/autoattack.py as the checker."""
    return ops.fab_projection_linf(t.contiguous(), w.contiguous(), b.contiguous())


class FABT:
    def __init__(self, model: _Model, eps, n_iter=100, n_target_classes=9, alpha_max=0.1, eta=1.05, beta=0.9):
        self.m, self.eps, self.n_iter, self.ntc = model, float(eps), n_iter, n_target_classes
        self.alpha_max, self.eta, self.beta = alpha_max, eta, beta

    def single_run(self, x, y, target_class):
        """fab_base.py:84-270 with is_targeted=True, use_rand_start=False, Linf."""
        y_pred = self.m.logits(x).max(1)[1]
        pred = (y_pred == y).nonzero().flatten()
        if pred.numel() == 0:
            return x
        la_target = self.m.logits(x).sort(dim=-1)[1][:, -target_class]
        im2, la2, lt2 = x[pred].contiguous(), y[pred].contiguous(), la_target[pred].contiguous()
        bs = im2.shape[0]
        adv = im2.clone()
        adv_c = x.clone()
        res2 = torch.full((bs,), 1e10, device=x.device)
        x1 = im2.clone()
        x0 = im2.reshape(bs, -1)
        for _ in range(self.n_iter):
            _, df, dg = self.m.loss_and_grad(x1, la2, "fab-diff", lt2)       # df [bs], dg [bs, c, h, w]
            w = dg.reshape(bs, -1)
            b = -df + (w * x1.reshape(bs, -1)).sum(dim=-1)
            # fab_base.py:194-232: both projections (current iterate and original point onto the linearised boundary) in one
            # launch, then step-size rule + convex combination + clamp in a second one
            d3, a0 = ops.fab_projection_linf(torch.cat((x1.reshape(bs, -1), x0), 0), torch.cat((w, w), 0), torch.cat((b, b), 0),
                                             want_dmax=True)
            ops.fab_combine_linf_(x1, d3[:bs], im2, d3[bs:], a0[:bs], a0[bs:], self.eta, self.alpha_max)
            is_adv = self.m.logits(x1).max(1)[1] != la2
            if is_adv.any():
                ia = is_adv.nonzero().flatten()
                t = (x1[ia] - im2[ia]).reshape(ia.shape[0], -1).abs().max(dim=1)[0]
                better = t < res2[ia]
                adv[ia[better]] = x1[ia[better]]
                res2[ia] = torch.where(better, t, res2[ia])
                x1[ia] = im2[ia] + (x1[ia] - im2[ia]) * self.beta
        succ = (res2 < 1e10).nonzero().flatten()
        adv_c[pred[succ]] = adv[succ]
        return adv_c

    def perturb(self, x, y, seed):
        """fab_base.py:274-334 (targeted branch, n_restarts = 1)."""
        x = x.detach().float().contiguous()
        y = y.detach().long().contiguous()
        torch.manual_seed(seed)
        torch.cuda.manual_seed(seed)
        adv = x.clone()
        acc = self.m.logits(x).max(1)[1] == y
        for tc in range(2, self.ntc + 2):
            idx = acc.nonzero().flatten()
            if idx.numel() == 0:
                break
            xs, ys = x[idx].contiguous(), y[idx].contiguous()
            adv_curr = self.single_run(xs, ys, tc)
            acc_curr = self.m.logits(adv_curr).max(1)[1] == ys
            res = (xs - adv_curr).abs().reshape(xs.shape[0], -1).max(1)[0]
            acc_curr = acc_curr | (res > self.eps)
            fooled = ~acc_curr
            acc[idx[fooled]] = False
            adv[idx[fooled]] = adv_curr[fooled]
        return adv


# ------------------------------------------------------------------------------------------------
# driver
# ------------------------------------------------------------------------------------------------
class AutoAttack:
    def __init__(self, model, norm="Linf", eps=0.3, seed=None, verbose=True, version="standard", attacks_to_run=None,
                 n_iter=100, n_queries=5000, n_target_classes=9):
        if norm != "Linf":
            raise NotImplementedError("AutoAttack norm %r: only Linf is on the B200 path (SURVEY 8f N4)" % norm)
        if version != "standard":
            raise NotImplementedError("AutoAttack version %r: only 'standard' is implemented" % version)
        self.m = _Model(model)
        self.eps, self.seed, self.verbose = float(eps), seed, verbose
        self.attacks_to_run = list(attacks_to_run) if attacks_to_run else ["apgd-ce", "apgd-t", "fab-t", "square"]
        self.apgd = APGD(self.m, eps, n_iter=n_iter, n_restarts=1, loss="ce")
        self.apgd_t = APGD(self.m, eps, n_iter=n_iter, n_restarts=1, loss="dlr-targeted", n_target_classes=n_target_classes)
        self.fab = FABT(self.m, eps, n_iter=n_iter, n_target_classes=n_target_classes)
        self.square = Square(self.m, eps, n_queries=n_queries)
        self.history = []

    def _seed(self):
        return next(_seed_counter) if self.seed is None else self.seed

    def run_standard_evaluation(self, x_orig, y_orig, bs=250):
        """autoattack.py:90-211."""
        x_orig = x_orig.detach().float().contiguous()
        y_orig = y_orig.detach().long().view(-1).contiguous()
        n = x_orig.shape[0]
        robust = torch.zeros(n, dtype=torch.bool, device=x_orig.device)
        for s0 in range(0, n, bs):
            robust[s0:s0 + bs] = self.m.logits(x_orig[s0:s0 + bs]).max(1)[1] == y_orig[s0:s0 + bs]
        self.history = [("clean", robust.float().mean().item())]
        x_adv = x_orig.clone()
        for attack in self.attacks_to_run:
            idcs = robust.nonzero().flatten()
            if idcs.numel() == 0:
                break
            for s0 in range(0, idcs.numel(), bs):
                bi = idcs[s0:s0 + bs]
                x, y = x_orig[bi].contiguous(), y_orig[bi].contiguous()
                if attack == "apgd-ce":
                    adv = self.apgd.perturb(x, y, self._seed())
                elif attack == "apgd-t":
                    adv = self.apgd_t.perturb(x, y, self._seed())
                elif attack == "fab-t":
                    adv = self.fab.perturb(x, y, self._seed())
                elif attack == "square":
                    adv = self.square.perturb(x, y, self._seed())
                else:
                    raise ValueError("Attack not supported")
                false_batch = self.m.logits(adv).max(1)[1] != y
                nr = bi[false_batch]
                robust[nr] = False
                x_adv[nr] = adv[false_batch]
            self.history.append((attack, robust.float().mean().item()))
            if self.verbose:
                print("robust accuracy after %s: %.2f%%" % (attack.upper(), 100 * self.history[-1][1]))
        return x_adv


def autoattack_linf(input, label, model, norm, eps, version, verbose):
    """Same signature and semantics as attack.py:35-38 (bs = the whole batch)."""
    if not input.is_cuda:
        raise TypeError("adversarial noise needs CUDA tensors")
    aa = AutoAttack(model, norm=norm, eps=eps, version=version, verbose=verbose)
    return aa.run_standard_evaluation(input, label, bs=input.shape[0])
