"""Adversarial Robustness: of the images classified correctly before the attack, the share still correct after it
(reference RobustART/metrics/AR_evaluator.py:23-39).

Same class name, `eval(clean_path, adv_path)` signature, print-out and return value (AR in percent).  Two reference quirks
are not reproduced: `parse_line` is declared without `self` there (calling it raises TypeError), and it takes the first two
`key: value` fields of a line as (prediction, label), which holds only for DALI-type lines; here every line is parsed as
JSON and the `prediction` / `label` fields are used whatever precedes them.  The reference hard-codes 50000 lines; here it
is the number of lines of the clean file (the adversarial file must be aligned with it, image for image)."""
import json


def _pred_label(line):
    info = json.loads(line)
    return int(info["prediction"]), int(info["label"])


class AdvRobustEvaluator:
    def eval(self, clean_path, adv_path):
        with open(clean_path) as f:
            lines_clean = f.readlines()
        with open(adv_path) as f:
            lines_att = f.readlines()
        assert len(lines_att) >= len(lines_clean), "adversarial result file shorter than the clean one"
        cnt_before_att = cnt_after_att = 0
        for clean, att in zip(lines_clean, lines_att):
            p, l = _pred_label(clean)
            if p == l:
                cnt_before_att += 1
                pa, la = _pred_label(att)
                if pa == la:
                    cnt_after_att += 1
        AR = cnt_after_att / max(cnt_before_att, 1) * 100
        print('Clean Acc: {}, Adversarial Robustness: {}'.format(cnt_before_att / max(len(lines_clean), 1) * 100, AR))
        return AR
