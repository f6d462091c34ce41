"""The criterion ``build_model`` returns (tuber_b200/models/criterion.py) against the reference's own ``SetCriterionAVA`` /
``SetCriterion`` + Hungarian matchers: tests/golden/criterion.npz was written by oracle/make_golden_criterion.py from the
unmodified reference classes (models/criterion.py:11-410, models/detr/matcher.py, matcher_ucf.py)."""
import os

import numpy as np
import pytest
import torch

import tuber_b200
from tuber_b200.models.criterion import HungarianMatcher, SetCriterion, SetCriterionAVA

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "criterion.npz")


def _case(g, prefix, target_keys, int_keys=()):
    L = g[f"{prefix}.pred_logits"].shape[0]
    layers = [{k: torch.from_numpy(g[f"{prefix}.{k}"][l]) for k in ("pred_logits", "pred_boxes", "pred_logits_b")} for l in range(L)]
    outputs = dict(layers[-1], aux_outputs=layers[:-1])
    counts = g[f"{prefix}.n_targets"]
    starts = np.concatenate([[0], np.cumsum(counts)])
    targets = []
    for b, n in enumerate(counts):
        t = {}
        for k in target_keys:
            a = g[f"{prefix}.t.{k}"]
            if k in ("boxes", "labels"):
                v = torch.from_numpy(a[starts[b]:starts[b] + n])
                t[k] = v.reshape(-1).long() if k in int_keys else v
            else:
                t[k] = torch.from_numpy(a[b]).reshape(-1)[0] if k == "key_pos" else torch.from_numpy(a[b]).reshape(-1)
        targets.append(t)
    want = {k.split(".loss.")[1]: float(g[k]) for k in g.files if k.startswith(prefix + ".loss.")}
    return outputs, targets, want


def _check(got, want):
    assert set(got) == set(want)
    for k, v in want.items():
        assert abs(float(got[k]) - v) <= 2e-6 * max(1.0, abs(v)), (k, float(got[k]), v)


@pytest.mark.parametrize("variant,evaluation", [("ava_eval", True), ("ava_train", False)])
def test_ava_criterion_matches_reference(variant, evaluation):
    g = np.load(GOLD)
    outputs, targets, want = _case(g, variant, ("boxes", "labels"))
    cfg = tuber_b200.load_cfg("TubeR_CSN152_AVA21.yaml", ["CONFIG.EVAL_ONLY", evaluation])
    _, crit, _ = tuber_b200.build_model(cfg)
    assert isinstance(crit, SetCriterionAVA) and crit.evaluation is evaluation
    _check(crit(outputs, targets), want)
    assert len(want) == 5 + 4 * 5                    # 4 losses + class_error on the last layer, 4 per auxiliary layer
    # the loop's weighted sum (utils/video_action_recognition.py:357-361) only reads keys the weight dictionary has
    assert all(k in crit.weight_dict for k in want if k != "class_error")


def test_jhmdb_criterion_matches_reference():
    g = np.load(GOLD)
    outputs, targets, want = _case(g, "jhmdb", ("boxes", "labels", "vis", "key_pos"), int_keys=("labels",))
    for t in targets:
        t["vis"] = t["vis"].long()
    cfg = tuber_b200.load_cfg("Tuber_CSN152_JHMDB.yaml", ["CONFIG.MODEL.TEMP_LEN", 8, "CONFIG.MODEL.DS_RATE", 8])
    _, crit, _ = tuber_b200.build_model(cfg)
    assert isinstance(crit, SetCriterion) and crit.matcher.cost_class == 1
    _check(crit(outputs, targets), want)


def test_matcher_is_optimal_and_one_to_one():
    torch.manual_seed(3)
    out = {"pred_boxes": torch.rand(2, 6, 4) * 0.4 + 0.3, "pred_logits_b": torch.randn(2, 6, 3), "pred_logits": torch.randn(2, 6, 80)}
    targets = [{"boxes": torch.cat((torch.zeros(3, 1), torch.rand(3, 4) * 0.4 + 0.3), 1)}, {"boxes": torch.zeros(0, 5)}]
    pairs = HungarianMatcher(12, 5, 2, "ava")(out, targets)
    assert len(pairs[0][0]) == 3 and len(set(pairs[0][0].tolist())) == 3 and sorted(pairs[0][1].tolist()) == [0, 1, 2]
    assert len(pairs[1][0]) == 0
    with pytest.raises(ValueError):
        HungarianMatcher(0, 0, 0)
