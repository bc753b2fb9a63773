"""robust.json from the REFERENCE's own merge_eval_res (this container only).

    python tests/golden/make_golden_robust_json.py

ImageNetCDataset.merge_eval_res (prototype/prototype/data/datasets/imagnetc.py:166-218) is cut out of the reference source by
AST position and run on a directory of seeded `{group}-{type}-{sev}-metric` files.  Output: tests/golden/robust_json.json =
{"metrics": {file name: {"top1", "top5"}}, "robust": the reference's robust.json}."""
import ast
import json
import os
import tempfile
import textwrap
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    src = open("/root/reference/prototype/prototype/data/datasets/imagnetc.py").read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "merge_eval_res")
    ns = {"json": json, "np": np}
    exec(textwrap.dedent("\n".join(src.split("\n")[fn.lineno - 1:fn.end_lineno])), ns)
    groups = {"noise": ["gaussian_noise", "shot_noise", "impulse_noise"], "blur": ["defocus_blur", "glass_blur", "motion_blur", "zoom_blur"],
              "weather": ["snow", "frost", "fog", "brightness"], "digital": ["contrast", "elastic_transform", "pixelate", "jpeg_compression"],
              "extra": ["speckle_noise", "spatter", "gaussian_blur", "saturate"]}
    rs = np.random.RandomState(4)
    metrics = {}
    with tempfile.TemporaryDirectory() as d:
        for g, types_ in groups.items():
            for t in types_:
                for s in range(1, 6):
                    m = {"top1": float(np.float32(rs.uniform(5, 80))), "top5": float(np.float32(rs.uniform(80, 99)))}
                    metrics["%s-%s-%d-metric" % (g, t, s)] = m
                    json.dump(m, open(os.path.join(d, "%s-%s-%d-metric" % (g, t, s)), "w"))
        me = types.SimpleNamespace(get_table=lambda res: [], logger=types.SimpleNamespace(info=lambda *a: None))
        ns["merge_eval_res"](me, d)
        robust = json.load(open(os.path.join(d, "robust.json")))
    json.dump({"metrics": metrics, "robust": robust}, open(os.path.join(HERE, "robust_json.json"), "w"))
    print("wrote robust_json.json", robust["all"])


if __name__ == "__main__":
    main()
