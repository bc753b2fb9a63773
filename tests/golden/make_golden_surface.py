"""The plugin surface of the reference, read from its sources without importing them (this container only).

    python tests/golden/make_golden_surface.py

  * noise_list, default_config, the keys of function_dict   RobustART/noise/utils/add_noise_utils.py:7-18,41-50
  * corruption_tuple names and order                         RobustART/noise/utils/imagenet_c/__init__.py:5-8
  * model_name_dict keys and types                           prototype/prototype/utils/model_config.py (the benchmark's name table)
  * ImageTransfer.get_params on 48 image shapes with random.seed(2024)   RobustART/noise/utils/imagenet_s_gen.py:199-239
The literals are evaluated with ast.literal_eval / name extraction; output: tests/golden/plugin_surface.json."""
import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def assigned(tree, name):
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id == name for t in node.targets):
            return node.value
    raise KeyError(name)


def value(node):
    """literal_eval that also folds the arithmetic the reference writes into its defaults (8 / 255, 3 / 40); no names, no calls."""
    for n in ast.walk(node):
        assert isinstance(n, (ast.Dict, ast.List, ast.Tuple, ast.Constant, ast.BinOp, ast.UnaryOp, ast.Div, ast.Mult, ast.Add, ast.Sub,
                              ast.USub, ast.Load)), type(n)
    return eval(compile(ast.Expression(node), "<surface>", "eval"), {"__builtins__": {}})


def main():
    t = ast.parse(open(os.path.join(REF, "RobustART/noise/utils/add_noise_utils.py")).read())
    out = {"noise_list": ast.literal_eval(assigned(t, "noise_list")), "default_config": value(assigned(t, "default_config")),
           "function_dict_keys": [ast.literal_eval(k) for k in assigned(t, "function_dict").keys]}
    t = ast.parse(open(os.path.join(REF, "RobustART/noise/utils/imagenet_c/__init__.py")).read())
    out["corruption_tuple"] = [e.id for e in assigned(t, "corruption_tuple").elts]
    t = ast.parse(open(os.path.join(REF, "prototype/prototype/utils/model_config.py")).read())
    md = assigned(t, "model_name_dict")
    names = {}
    for k, v in zip(md.keys, md.values):
        try:
            names[ast.literal_eval(k)] = ast.literal_eval(v).get("type")
        except Exception:
            names[ast.literal_eval(k)] = None
    out["model_name_dict_types"] = names
    # ImageTransfer.get_params (imagenet_s_gen.py:199-239), executed from source with Python's `random` seeded
    import math
    import random
    import textwrap
    import types
    src = open(os.path.join(REF, "RobustART/noise/utils/imagenet_s_gen.py")).read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "get_params")
    ns = {"math": math, "random": random}
    exec(textwrap.dedent("\n".join(src.split("\n")[fn.lineno - 1:fn.end_lineno])), ns)
    random.seed(2024)
    shapes = [(375, 500), (500, 375), (224, 224), (30, 600), (600, 30), (64, 48), (1, 1), (17, 400)] * 6
    out["get_params"] = {"seed": 2024, "shapes": shapes,
                         "params": [list(ns["get_params"](None, types.SimpleNamespace(shape=(h, w, 3)))) for h, w in shapes]}
    json.dump(out, open(os.path.join(HERE, "plugin_surface.json"), "w"), indent=1)
    print("wrote plugin_surface.json", out["noise_list"], len(out["corruption_tuple"]), len(names))


if __name__ == "__main__":
    main()
