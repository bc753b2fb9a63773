"""Worst-Case Adversarial Robustness: of the images classified correctly before the attacks, the share still correct after
EVERY attack of a list (reference RobustART/metrics/WCAR_evaluator.py:23-44).  Same class name, signature, print-out and
return value; parsing as in AR_evaluator.py."""
from .AR_evaluator import _pred_label


class WorstCaseAdvRobustEvaluator:
    def eval(self, clean_path, multi_adv_result_paths):
        with open(clean_path) as f:
            lines_clean = f.readlines()
        list_lines_att = []
        for p in multi_adv_result_paths:
            with open(p) as f:
                list_lines_att.append(f.readlines())
        cnt_before_att = cnt_after_att = 0
        for ind, clean in enumerate(lines_clean):
            p, l = _pred_label(clean)
            if p == l:
                cnt_before_att += 1
                if all(pa == la for pa, la in (_pred_label(lines[ind]) for lines in list_lines_att)):
                    cnt_after_att += 1
        WCAR = cnt_after_att / max(cnt_before_att, 1) * 100
        print('Worst-Case Adversarial Robustness: {}'.format(WCAR))
        return WCAR
