"""Result-file compatibility mode: the reference's per-image result lines, rank files, merged `.all` file and
the file-based evaluator, for downstream scripts that parse them (RobustART/metrics/AR_evaluator.py:23-39,
WCAR_evaluator.py:23-44, imagnetc.py:166-218).  The fast path of this repo never writes them (device counters + one
all-reduce); this module is the opt-in `data.test.dump_results: true` / `--dump-results` path.

Formats reproduced byte for byte:
  * line            ImageNetDataset.dump, imagenet_dataset.py:250-277:
                    json.dumps({"filename", "image_id", "prediction", "label", "score": [float("%.8f" % s) ...]}) + "\\n"
                    ("filename"/"image_id" omitted for the DALI-type output, :268-276)
  * rank files      `<prefix>/results.txt.rank{r}`, merged by concatenation in rank order into `results.txt.all`
                    (base_dataset.py:116-133; benchmark_eval_adv.py:211-213, cls_solver.py:385-396)
  * evaluation      ImageNetEvaluator.eval, imagenet_evaluator.py:49-67: top-k over the ROUNDED scores with torch.topk

`float("%.8f" % s)` is vectorised exactly: q = round-half-even(s * 1e8) as an integer (ties and near-ties re-done with
Python's own formatting), and q / 1e8 is the correctly rounded double of the decimal string -- the same double
`float("0.dddddddd")` yields -- so json.dumps prints the same shortest repr.
"""
from __future__ import annotations

import json
import os
from typing import Iterable, Optional, Sequence

import numpy as np


def round8(scores: np.ndarray) -> np.ndarray:
    """float64 array equal elementwise to float("%.8f" % s) for float32/float64 input scores."""
    s = np.asarray(scores, dtype=np.float64)
    shape = s.shape
    s = s.reshape(-1)
    x = s * 1e8
    q = np.rint(x)
    out = q / 1e8
    # the product carries <= 1 ulp of rounding error: re-do everything within 1e-4 of a tie with the exact decimal formatting
    frac = np.abs(x - np.floor(x) - 0.5)
    for i in np.nonzero((frac < 1e-4) | ~np.isfinite(x))[0]:
        out[i] = float("%.8f" % s[i])
    return out.reshape(shape)


def format_lines(prediction, label, score, filename: Optional[Sequence[str]] = None, image_id: Optional[Sequence[int]] = None) -> str:
    """The text ImageNetDataset.dump writes for one batch (imagenet_dataset.py:250-277)."""
    prediction, label = np.asarray(prediction), np.asarray(label)
    vals = round8(np.asarray(score))
    lines = []
    for i in range(prediction.shape[0]):
        if filename is not None:
            res = {"filename": filename[i], "image_id": int(image_id[i]), "prediction": int(prediction[i]), "label": int(label[i]),
                   "score": vals[i].tolist()}
        else:
            res = {"prediction": int(prediction[i]), "label": int(label[i]), "score": vals[i].tolist()}
        lines.append(json.dumps(res, ensure_ascii=False) + "\n")
    return "".join(lines)


class ResultWriter:
    """`<dir>/<stem>.rank{r}` writer; `write_batch` takes what the solver loop has after softmax/topk (cls_solver.py:417-424)."""

    def __init__(self, result_dir: str, rank: int, stem: str = "results.txt"):
        os.makedirs(result_dir, exist_ok=True)
        self.path = os.path.join(result_dir, "%s.rank%d" % (stem, rank))
        self._f = open(self.path, "w")

    def write_batch(self, prediction, label, score, filename=None, image_id=None):
        self._f.write(format_lines(prediction, label, score, filename, image_id))
        self._f.flush()

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def merge(prefix: str, world_size: int) -> str:
    """Concatenate `<prefix>{rank}` for rank in order into `<prefix minus last extension>.all` (base_dataset.py:116-133)."""
    merged = prefix.rsplit(".", 1)[0] + ".all"
    with open(merged, "w") as out:
        for rank in range(world_size):
            part = prefix + str(rank)
            assert os.path.exists(part), "No such file or directory: %s" % part
            with open(part, "r") as fin:
                for line in fin:
                    out.write(line)
    return merged


def load_res(res_file: str) -> dict:
    """ImageNetEvaluator.load_res (imagenet_evaluator.py:24-47): column-wise lists, malformed lines skipped."""
    res = {}
    with open(res_file) as f:
        for line in f:
            try:
                info = json.loads(line)
            except json.JSONDecodeError:
                continue
            for k, v in info.items():
                res.setdefault(k, []).append(v)
    return res


def evaluate(res_file: str, topk: Iterable[int] = (1, 5)) -> dict:
    """ImageNetEvaluator.eval (imagenet_evaluator.py:49-67) on a merged result file: {"top1": %, "top5": %}."""
    import torch
    topk = tuple(topk)
    res = load_res(res_file)
    pred = torch.from_numpy(np.array(res["score"]))
    label = torch.from_numpy(np.array(res["label"]))
    num = pred.size(0)
    _, idx = pred.topk(max(topk), 1, True, True)
    correct = idx.t().eq(label.reshape(1, -1).expand_as(idx.t()))
    return {"top%d" % k: correct[:k].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / num).item() for k in topk}


IMAGENET_C_GROUPS = {
    "noise": ["gaussian_noise", "shot_noise", "impulse_noise"],
    "blur": ["defocus_blur", "glass_blur", "motion_blur", "zoom_blur"],
    "weather": ["snow", "frost", "fog", "brightness"],
    "digital": ["contrast", "elastic_transform", "pixelate", "jpeg_compression"],
    "extra": ["speckle_noise", "spatter", "gaussian_blur", "saturate"],
}


def merge_imagenet_c_metrics(result_path: str, severities=(1, 2, 3, 4, 5)) -> dict:
    """ImageNetCDataset.merge_eval_res (datasets/imagnetc.py:166-218) on a directory of `{group}-{type}-{sev}-metric` files
    (what ImageNetCEvaluator.eval and EvalSolver.evaluate_imagenet_c write): mean top-1 ERROR over the severities per
    corruption type, `all_with_extra` / `all_without_extra` over the types, dumped to `robust.json`.  Same numbers as the
    reference (np.average of the same lists in the same order); plain floats instead of numpy scalars in the JSON."""
    import numpy as np
    all_data = {"all": {"all_with_extra": {}, "all_without_extra": {}}}
    avg, avg_wo_extra = [], []
    for group, types in IMAGENET_C_GROUPS.items():
        all_data[group] = {}
        for noise_type in types:
            lst = []
            for i in severities:
                with open(os.path.join(result_path, "%s-%s-%d-metric" % (group, noise_type, i))) as f:
                    lst.append(100 - float(json.load(f)["top1"]))
            all_data[group][noise_type] = float(np.average(lst))
            if group != "extra":
                avg_wo_extra.append(np.average(lst))
            avg.append(np.average(lst))
    all_data["all"]["all_with_extra"] = float(np.average(avg))
    all_data["all"]["all_without_extra"] = float(np.average(avg_wo_extra))
    with open(os.path.join(result_path, "robust.json"), "w") as f:
        json.dump(all_data, f)
    return all_data
