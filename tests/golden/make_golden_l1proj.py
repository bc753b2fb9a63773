"""Golden vectors for K-L1proj from the REFERENCE's own vendored L1_projection (autopgd_base.py:19-83), this container only.

    python tests/golden/make_golden_l1proj.py        -> tests/golden/l1_projection.npz

The function is cut out of the reference source by AST position and executed (the module itself imports foolbox-free code but
sits in a package whose __init__ pulls absent dependencies)."""
import ast
import os
import textwrap

import numpy as np
import torch

SRC = "/root/reference/RobustART/noise/utils/adv/Attacks/autoattack/autopgd_base.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_fn():
    src = open(SRC).read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "L1_projection")
    ns = {"torch": torch}
    exec(textwrap.dedent("\n".join(src.split("\n")[fn.lineno - 1:fn.end_lineno])), ns)
    return ns["L1_projection"]


def main():
    f = reference_fn()
    g = torch.Generator().manual_seed(5)
    out = {}
    for k, (rows, dim, eps, scale) in enumerate([(6, 48, 3.0, 0.4), (5, 300, 12.0, 0.3), (4, 1000, 5.0, 0.05), (3, 64, 100.0, 0.5)]):
        x = torch.rand(rows, dim, generator=g)
        y = torch.randn(rows, dim, generator=g) * scale
        y[:, ::5] = 0
        out["x%d" % k], out["y%d" % k], out["eps%d" % k] = x.numpy(), y.numpy(), np.float64(eps)
        out["d%d" % k] = f(x, y, eps).numpy()
        z = y + torch.from_numpy(out["d%d" % k])
        print(k, "||y||1 max %.2f -> ||y+d||1 max %.3f (eps %.1f), box [%.3f, %.3f]" % (y.abs().sum(1).max(), z.abs().sum(1).max(), eps, (x + z).min(), (x + z).max()))
    np.savez(os.path.join(HERE, "l1_projection.npz"), **out)


if __name__ == "__main__":
    main()
