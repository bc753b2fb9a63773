"""ImageNet-C file evaluator (reference RobustART/metrics/imagenetc_evaluator.py:28-77): top-k over the rounded scores of a
merged result file, metric written next to it with "results.txt.all" replaced by "metric".  Same class name, `eval(res_file)`
signature and output file; the arithmetic is robustart_b200.resultfile.evaluate (== ImageNetEvaluator.eval)."""
import json

from robustart_b200 import resultfile


class ClsMetric:
    def __init__(self, metric_dict=None):
        self.metric = dict(metric_dict or {})
        self.cmp_key, self.v = None, None

    def set_cmp_key(self, key):
        self.cmp_key = key
        self.v = self.metric[key]

    def __repr__(self):
        return f'metric={self.metric} key={self.cmp_key}'

    __str__ = __repr__


class ImageNetCEvaluator:
    def __init__(self, topk=(1, 5)):
        self.topk = list(topk)

    def load_res(self, res_file):
        return resultfile.load_res(res_file)

    def eval(self, res_file):
        metric = ClsMetric(resultfile.evaluate(res_file, self.topk))
        metric.set_cmp_key(f'top{self.topk[0]}')
        with open(res_file.replace("results.txt.all", "metric"), "w") as f:
            json.dump(metric.metric, f)
        return metric
