"""Golden result-file text from the REFERENCE's own dump / merge / eval code (this container only).

    python tests/golden/make_golden_results.py

The reference modules cannot be imported (easydict, ffmpeg, DALI ...), so the three functions are cut out of their
source files by AST position and executed as they stand:
  ImageNetDataset.dump   prototype/prototype/data/datasets/imagenet_dataset.py:250-277
  BaseDataset.merge      prototype/prototype/data/datasets/base_dataset.py:116-133   (link.get_world_size() stubbed)
  ImageNetEvaluator.load_res / eval   prototype/prototype/data/metrics/imagenet_evaluator.py:24-67
Inputs: softmax scores of seeded random logits plus hand-placed rounding traps (exact ties at the 9th decimal, values
around 1e-4 where repr switches to exponent form, denormals, 1.0).  Output: tests/golden/result_lines.json.
"""
import ast
import io
import json
import os
import sys
import tempfile
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/prototype/prototype"


def cut(path, cls, fn):
    src = open(path).read()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name == fn:
                    return textwrap.dedent("\n".join(src.split("\n")[f.lineno - 1:f.end_lineno]))
    raise KeyError((cls, fn))


def make_scores():
    g = torch.Generator().manual_seed(11)
    s = torch.softmax(torch.randn(4, 1000, generator=g) * 3, 1).numpy().astype(np.float32)
    traps = np.array([0.5, 1.0, 0.0, 1e-4, 9.9999e-5, 1.00005e-4, 1.5e-8, 0.5e-8, 2.5e-8, 4.9999999e-9, 5.0000001e-9,
                      0.123456785, 0.123456775, 0.999999995, 0.99999999, 1e-45, 3.3e-39, 0.000012345678, 0.1, 0.25000001,
                      0.00000001, 0.000000015, 0.33333334, 0.6666667, 7.0064923e-45], dtype=np.float32)
    s[3, :traps.size] = traps
    return s


def main():
    ns = {"json": json}
    exec(cut(os.path.join(REF, "data/datasets/imagenet_dataset.py"), "ImageNetDataset", "dump"), ns)
    mns = {"os": os, "link": types.SimpleNamespace(get_world_size=lambda: 2)}
    exec(cut(os.path.join(REF, "data/datasets/base_dataset.py"), "BaseDataset", "merge"), mns)
    ens = {"json": json, "torch": torch, "np": np, "ClsMetric": lambda res: types.SimpleNamespace(metric=res, set_cmp_key=lambda k: None)}
    exec(cut(os.path.join(REF, "data/metrics/imagenet_evaluator.py"), "ImageNetEvaluator", "load_res"), ens)
    exec(cut(os.path.join(REF, "data/metrics/imagenet_evaluator.py"), "ImageNetEvaluator", "eval"), ens)

    scores = make_scores()
    label = np.array([int(scores[0].argmax()), 3, int(np.argsort(-scores[2])[3]), 0])
    pred = scores.argmax(1)
    me = types.SimpleNamespace(tensor2numpy=lambda x: x)
    texts = {}
    for kind in ("pytorch", "dali"):
        w = io.StringIO()
        out = {"prediction": pred, "label": label, "score": scores}
        if kind == "pytorch":
            out.update(filename=["val/a_%d.JPEG" % i for i in range(4)], image_id=[7, 8, 9, 10])
        ns["dump"](me, w, out)
        texts[kind] = w.getvalue()
    with tempfile.TemporaryDirectory() as d:
        lines = texts["pytorch"].splitlines(True)
        open(os.path.join(d, "results.txt.rank0"), "w").write("".join(lines[:3]))
        open(os.path.join(d, "results.txt.rank1"), "w").write("".join(lines[3:]))
        merged = mns["merge"](None, os.path.join(d, "results.txt.rank"))
        assert os.path.basename(merged) == "results.txt.all"
        merged_text = open(merged).read()
        ev = types.SimpleNamespace(topk=[1, 5], load_res=lambda p: ens["load_res"](None, p))
        metric = ens["eval"](ev, merged).metric
    json.dump({"scores_f32_bytes_hex": scores.tobytes().hex(), "shape": list(scores.shape), "label": label.tolist(),
               "prediction": pred.tolist(), "text_pytorch": texts["pytorch"], "text_dali": texts["dali"], "merged": merged_text,
               "metric": metric}, open(os.path.join(HERE, "result_lines.json"), "w"))
    print("wrote result_lines.json", len(texts["pytorch"]), metric)


if __name__ == "__main__":
    sys.exit(main())
