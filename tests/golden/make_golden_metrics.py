"""Golden vectors for the eval reduction from the REFERENCE's own code (this container only).

    python tests/golden/make_golden_metrics.py

`DistributedSampler` (prototype/prototype/data/sampler.py:8-52) and `accuracy` (prototype/prototype/utils/misc.py:441-455) are cut
out of the reference sources by AST position and executed as they stand (their modules import linklink / easydict and cannot be
imported).  Output: tests/golden/metrics_reference.json -- per-rank index lists for several (N, world) pairs with round_up False
(what imagenet_dataloader.py:271 passes) and top-1/top-5 precision of seeded logits that include exact ties."""
import ast
import json
import math
import os
import textwrap

import numpy as np
import torch
from torch.utils.data.sampler import Sampler

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/prototype/prototype"


def cut(path, name):
    src = open(path).read()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name == name:
            return textwrap.dedent("\n".join(src.split("\n")[node.lineno - 1:node.end_lineno]))
    raise KeyError(name)


def main():
    ns = {"torch": torch, "Sampler": Sampler, "math": math, "np": np, "link": None}
    exec(cut(os.path.join(REF, "data/sampler.py"), "DistributedSampler"), ns)
    exec(cut(os.path.join(REF, "utils/misc.py"), "accuracy"), ns)
    out = {"sampler": {}, "accuracy": {}}
    for n, w in [(10, 4), (7, 1), (1001, 8), (50, 3), (64, 8), (50000, 8)]:
        ranks = [[int(i) for i in ns["DistributedSampler"](range(n), world_size=w, rank=r, round_up=False)] for r in range(w)]
        if n > 2000:          # the full validation set: lengths + SHA-256 of the int32 index bytes per rank
            import hashlib
            ranks = [{"len": len(r), "sha256": hashlib.sha256(np.asarray(r, np.int32).tobytes()).hexdigest()} for r in ranks]
        out["sampler"]["%d/%d" % (n, w)] = ranks
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(40, 12, generator=g).round(decimals=1)       # one decimal: plenty of exact ties
    target = torch.randint(0, 12, (40,), generator=g)
    top1, top5 = ns["accuracy"](logits, target, topk=(1, 5))
    out["accuracy"] = {"logits": logits.tolist(), "target": target.tolist(), "top1": float(top1), "top5": float(top5)}
    json.dump(out, open(os.path.join(HERE, "metrics_reference.json"), "w"))
    print("wrote metrics_reference.json", {k: len(v) for k, v in out["sampler"].items()}, float(top1), float(top5))


if __name__ == "__main__":
    main()
