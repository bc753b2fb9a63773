"""Golden vectors from the REFERENCE's own vendored attack code (this container only, CPU).

    python tests/golden/make_golden_attacks.py

Imports RobustART/noise/utils/adv/Attacks/{imfgsm_attack.py, autoattack/} from /root/reference by path.  The vendored code
hard-codes `.cuda()`; on this GPU-less box `torch.Tensor.cuda` / `torch.cuda.FloatTensor` are patched to CPU no-ops for the
duration of the run -- the arithmetic is untouched.  Pinned pieces:
  * `_mim_whitebox`                 imfgsm_attack.py:62-93     (tiny seeded CNN, random start from torch.manual_seed)
  * `APGDAttack.dlr_loss`, `dlr_loss_targeted`   autopgd_base.py:198-204,599-604
  * `projection_linf`               fab_projections.py:7-59
  * `SquareAttack.p_selection`      square.py:192-219
  * `pgd_linf_attack`               prototype/prototype/solver/adv_cls_solver_train_pgd_new.py:67-105 (the reference's own
    PGD-Linf loop: the second anchor of the foolbox restatement)
  * `APGDAttack.attack_single_run` (CE, DLR) and `APGDAttack_targeted.attack_single_run` (DLR-targeted), Linf, 20 iterations on
    the tiny CNN with device='cpu' (autopgd_base.py:208-448): inputs, the seed of the random start, and the four outputs
    (x_best, acc, loss_best, x_best_adv) -- what the PRODUCT's APGD control flow is checked against on CPU.
  * `FABAttack_PT.attack_single_run` targeted (classes 2 and 3), 15 iterations, no random start (fab_base.py:84-270).
  * `SquareAttack.perturb` (Linf, 300 queries, seed 0) on the same inputs (square.py:221-294).
  * `AutoAttack.run_standard_evaluation` with apgd-ce, apgd-t, fab-t (10 iterations each) and square (200 queries), seed 0: final
    adversarials and the robust accuracy after every stage (autoattack.py:90-211).
Output: tests/golden/attack_pieces.npz."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ATT = "/root/reference/RobustART/noise/utils/adv/Attacks"


def tiny_model():
    torch.manual_seed(123)
    m = nn.Sequential(nn.Conv2d(3, 8, 3, 2, 1), nn.ReLU(), nn.Conv2d(8, 16, 3, 2, 1), nn.ReLU(), nn.AdaptiveAvgPool2d(1), nn.Flatten(),
                      nn.Linear(16, 10))
    return m.eval()


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    spec = importlib.util.spec_from_file_location("ref_imfgsm", os.path.join(ATT, "imfgsm_attack.py"))
    mim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mim)
    sys.path.insert(0, ATT)
    import autoattack.autopgd_base as apgd      # noqa: E402  (the vendored package)
    import autoattack.fab_projections as fabp   # noqa: E402
    import autoattack.square as sq              # noqa: E402
    out = {}
    # ---- MI-FGSM ----
    model = tiny_model()
    g = torch.Generator().manual_seed(5)
    X = torch.rand(3, 3, 32, 32, generator=g)
    y = torch.tensor([1, 7, 3])
    eps, steps, step_size, decay = 8 / 255, 6, 0.002, 1.0
    torch.manual_seed(77)
    adv = mim._mim_whitebox(model, X, y, eps, steps, step_size, decay).detach()
    torch.manual_seed(77)
    u = torch.FloatTensor(*X.shape).uniform_(0, 1)          # the same generator draws the reference's uniform_(-eps, eps) consumed
    out.update(mim_X=X.numpy(), mim_y=y.numpy(), mim_u=u.numpy(), mim_adv=adv.numpy(),
               mim_cfg=np.array([eps, steps, step_size, decay], np.float64))
    # ---- DLR losses ----
    g = torch.Generator().manual_seed(9)
    z = torch.randn(64, 20, generator=g) * 3
    yy = torch.randint(0, 20, (64,), generator=g)
    yy[:20] = z[:20].argmax(1)
    tgt = z.sort(1)[1][:, -3].contiguous()
    self_ = types.SimpleNamespace(y_target=tgt)
    out.update(dlr_z=z.numpy(), dlr_y=yy.numpy(), dlr_t=tgt.numpy(),
               dlr=apgd.APGDAttack.dlr_loss(self_, z, yy).numpy(), dlr_targeted=apgd.APGDAttack_targeted.dlr_loss_targeted(self_, z, yy).numpy())
    # ---- FAB projection ----
    g = torch.Generator().manual_seed(11)
    t = torch.rand(48, 60, generator=g)
    w = torch.randn(48, 60, generator=g)
    w[:, ::7] = 0
    b = torch.randn(48, generator=g) * 2
    out.update(proj_t=t.numpy(), proj_w=w.numpy(), proj_b=b.numpy(), proj=fabp.projection_linf(t, w, b).numpy())
    # ---- Square schedule ----
    s = types.SimpleNamespace(rescale_schedule=False, n_queries=5000, p_init=0.8)
    its = np.array([0, 1, 10, 11, 50, 51, 200, 201, 500, 501, 1000, 1001, 2000, 2001, 4000, 4001, 6000, 6001, 8000, 8001, 9999])
    out.update(sq_it=its, sq_p=np.array([sq.SquareAttack.p_selection(s, int(i)) for i in its]))
    # ---- the reference's in-repo PGD-Linf loop (adv_cls_solver_train_pgd_new.py:67-105), cut out by AST position ----
    import ast
    import textwrap
    import torch.nn.functional as F
    src = open("/root/reference/prototype/prototype/solver/adv_cls_solver_train_pgd_new.py").read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "pgd_linf_attack")
    ns = {"torch": torch, "F": F}
    exec(textwrap.dedent("\n".join(src.split("\n")[fn.lineno - 1:fn.end_lineno])), ns)
    g = torch.Generator().manual_seed(31)
    xp = torch.rand(4, 3, 32, 32, generator=g)
    yp = torch.tensor([0, 4, 2, 9])
    eps_p, rel, steps_p = 8 / 255, 3 / 40, 10
    torch.manual_seed(99)
    advp = ns["pgd_linf_attack"](model, xp, yp, eps=eps_p, alpha=rel * eps_p, steps=steps_p, random_start=True)
    torch.manual_seed(99)
    up = torch.empty_like(xp).uniform_(0, 1)
    out.update(pgd_x=xp.numpy(), pgd_y=yp.numpy(), pgd_u=up.numpy(), pgd_adv=advp.numpy(), pgd_cfg=np.array([eps_p, rel, steps_p], np.float64))
    # ---- APGD single runs (Linf) ----
    for p in model.parameters():
        p.requires_grad_(False)
    g = torch.Generator().manual_seed(21)
    xa = torch.rand(8, 3, 32, 32, generator=g)
    xa = (xa * 0.3 + torch.rand(8, 3, 1, 1, generator=g) * 0.7).clamp(0, 1)      # differing global colour: differing predictions
    with torch.no_grad():
        ya = model(xa).argmax(1)
    ya[7] = (ya[7] + 1) % 10                                # one sample starts misclassified; some of the others stay robust
    eps_a = 8 / 255
    out.update(apgd_x=xa.numpy(), apgd_y=ya.numpy(), apgd_cfg=np.array([eps_a, 20, 42], np.float64))
    for loss in ("ce", "dlr"):
        a = apgd.APGDAttack(model, n_restarts=1, n_iter=20, verbose=False, eps=eps_a, norm='Linf', eot_iter=1, rho=.75, seed=0,
                            device='cpu', loss=loss)
        a.init_hyperparam(xa)
        torch.manual_seed(42)
        xb, acc, lb, xba = a.attack_single_run(xa, ya)
        out.update({"apgd_%s_x_best" % loss: xb.detach().numpy(), "apgd_%s_acc" % loss: acc.numpy(), "apgd_%s_loss" % loss: lb.detach().numpy(),
                    "apgd_%s_x_best_adv" % loss: xba.detach().numpy()})
    at = apgd.APGDAttack_targeted(model, n_restarts=1, n_iter=20, verbose=False, eps=eps_a, norm='Linf', eot_iter=1, rho=.75, seed=0,
                                  device='cpu')
    at.init_hyperparam(xa)
    with torch.no_grad():
        at.y_target = model(xa).sort(dim=1)[1][:, -2]
    torch.manual_seed(42)
    xb, acc, lb, xba = at.attack_single_run(xa, ya)
    out.update(apgd_t_target=at.y_target.numpy(), apgd_t_x_best=xb.detach().numpy(), apgd_t_acc=acc.numpy(), apgd_t_loss=lb.detach().numpy(),
               apgd_t_x_best_adv=xba.detach().numpy())
    # ---- FAB-T single runs (Linf, no random start) ----
    import autoattack.fab_pt as fab             # noqa: E402
    f = fab.FABAttack_PT(model, n_restarts=1, n_iter=15, eps=eps_a, norm='Linf', targeted=True, device='cpu', verbose=False)
    for tc in (2, 3):
        f.target_class = tc
        out["fab_t%d" % tc] = f.attack_single_run(xa.clone(), ya.clone(), use_rand_start=False, is_targeted=True).detach().numpy()
    # ---- Square (Linf): 300 queries at 16/255, seed 0 (square.py:221-294 + perturb) ----
    import autoattack.square as refsq           # noqa: E402
    sq_att = refsq.SquareAttack(model, p_init=.8, n_queries=300, eps=16 / 255, norm='Linf', n_restarts=1, seed=0, verbose=False,
                                device='cpu', resc_schedule=False)
    out.update(square_adv=sq_att.perturb(xa.clone(), ya.clone()).detach().numpy(), square_cfg=np.array([16 / 255, 300, 0], np.float64))
    # ---- the AutoAttack driver: apgd-ce -> apgd-t -> fab-t on the shrinking robust set (autoattack.py:90-211) ----
    # The reference wraps the
    # model in NormalizeModel, so the tiny CNN sees ImageNet-normalised input here.
    import autoattack.autoattack as refaa       # noqa: E402
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    g = torch.Generator().manual_seed(21)
    xd = torch.rand(10, 3, 32, 32, generator=g)
    xd = (xd * 0.3 + torch.rand(10, 3, 1, 1, generator=g) * 0.7).clamp(0, 1)
    with torch.no_grad():
        yd = model((xd - mean) / std).argmax(1)
    yd[9] = (yd[9] + 1) % 10
    out.update(aa_x=xd.numpy(), aa_y=yd.numpy(), aa_cfg=np.array([eps_a, 10, 0], np.float64))
    stages = ['apgd-ce', 'apgd-t', 'fab-t', 'square']
    hist = []
    for k in range(1, 5):
        aa = refaa.AutoAttack(model, norm='Linf', eps=eps_a, seed=0, verbose=False, version='custom', attacks_to_run=stages[:k], device='cpu')
        aa.apgd.n_restarts, aa.apgd.n_iter = 1, 10
        aa.apgd_targeted.n_iter, aa.apgd_targeted.n_target_classes = 10, 9
        aa.fab.n_restarts, aa.fab.n_iter, aa.fab.n_target_classes = 1, 10, 9
        aa.square.n_queries = 200
        adv = aa.run_standard_evaluation(xd.clone(), yd.clone(), bs=10)
        with torch.no_grad():
            hist.append(float((model((adv - mean) / std).argmax(1) == yd).float().mean()))
    out.update(aa_adv=adv.detach().numpy(), aa_robust_after=np.array(hist))
    np.savez_compressed(os.path.join(HERE, "attack_pieces.npz"), **out)
    print("wrote attack_pieces.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
