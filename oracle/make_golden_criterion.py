"""Generate tests/golden/criterion.npz from the UNMODIFIED reference criterion + matchers (build container only).

TEST INFRASTRUCTURE.  Imports /root/reference/models/criterion.py, models/detr/matcher.py and matcher_ucf.py as-is, builds
``SetCriterionAVA`` / ``SetCriterion`` with the constructor arguments the reference's ``build_model`` passes
(models/tuber_ava.py:184-216) and evaluates them on seeded synthetic outputs (last layer + 5 auxiliary layers) and targets
in the layout the evaluation loop hands over (utils/video_action_recognition.py:282-305; datasets/ava_frame.py:76-127).
Stored: the inputs and every entry of the returned loss dictionary, for three variants:

  ava_eval   EVAL_ONLY True   (un-weighted BCE)            ava_train  EVAL_ONLY False (matched queries weighted by LOSS_COFS.WEIGHT)
  jhmdb      SetCriterion + matcher_ucf: softmax classes with a no-object class, queries of the key frame selected by key_pos

    python oracle/make_golden_criterion.py
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("TUBER_REFERENCE", "/root/reference")


def ava_case(g, B=4, Q=15, C=80, L=6, counts=(2, 1, 3, 0)):
    outs = []
    for _ in range(L):
        outs.append({"pred_logits": torch.randn(B, Q, C, generator=g),
                     "pred_boxes": torch.cat((torch.rand(B, Q, 2, generator=g) * 0.6 + 0.2, torch.rand(B, Q, 2, generator=g) * 0.3 + 0.05), -1),
                     "pred_logits_b": torch.randn(B, Q, 3, generator=g) * 2})
    targets = []
    for b, n in enumerate(counts):
        boxes = torch.cat((torch.full((n, 1), 16.0), torch.rand(n, 2, generator=g) * 0.6 + 0.2, torch.rand(n, 2, generator=g) * 0.3 + 0.05), -1)
        labels = (torch.rand(n, C, generator=g) < 0.04).float()
        for i in range(n):
            labels[i, int(torch.randint(0, C, (1,), generator=g))] = 1.0       # at least one action per person
        targets.append({"boxes": boxes, "labels": labels})
    return outs, targets


def jhmdb_case(g, B=3, NQ=10, T=8, C=21, L=6):
    Q = NQ * T
    outs = []
    for _ in range(L):
        outs.append({"pred_logits": torch.randn(B, Q, C + 1, generator=g),
                     "pred_boxes": torch.cat((torch.rand(B, Q, 2, generator=g) * 0.6 + 0.2, torch.rand(B, Q, 2, generator=g) * 0.3 + 0.05), -1),
                     "pred_logits_b": torch.randn(B, 2, generator=g)})
    targets = []
    for b in range(B):
        n = 1 + b % 2
        boxes = torch.cat((torch.full((n, 1), 4.0), torch.rand(n, 2, generator=g) * 0.6 + 0.2, torch.rand(n, 2, generator=g) * 0.3 + 0.05), -1)
        targets.append({"boxes": boxes, "labels": torch.randint(0, C, (n,), generator=g), "vis": torch.tensor([1]),
                        "key_pos": torch.tensor(int(torch.randint(0, T, (1,), generator=g)))})
    return outs, targets


def pack(prefix, outs, targets, losses, store):
    for k in outs[0]:
        store[f"{prefix}.{k}"] = torch.stack([o[k] for o in outs]).numpy()          # (L, ...): index L-1 is the last layer
    store[f"{prefix}.n_targets"] = np.array([len(t["boxes"]) for t in targets])
    for k in targets[0]:
        per_box = k in ("boxes", "labels")
        width = max(int(t[k].numel() // max(1, len(t["boxes"]))) for t in targets) if per_box else 1
        store[f"{prefix}.t.{k}"] = torch.cat([t[k].reshape(len(t["boxes"]), width) if per_box else t[k].reshape(1, 1)
                                              for t in targets]).numpy()
    for k, v in losses.items():
        store[f"{prefix}.loss.{k}"] = np.asarray(float(v), dtype=np.float64)


def main():
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        from models.criterion import SetCriterion, SetCriterionAVA
        from models.detr.matcher import HungarianMatcher as MatcherAVA
        from models.detr.matcher_ucf import HungarianMatcher as MatcherUCF
    store = {}
    g = torch.Generator().manual_seed(77)

    def weight_dict(ce):
        w = {"loss_ce": ce, "loss_bbox": 5, "loss_giou": 2, "loss_ce_b": 1}
        for i in range(5):
            w.update({f"{k}_{i}": v for k, v in list(w.items())[:4]})
        return w

    outs, targets = ava_case(g)
    model_out = dict(outs[-1], aux_outputs=outs[:-1])
    for name, evaluation in (("ava_eval", True), ("ava_train", False)):
        crit = SetCriterionAVA(10, 80, num_queries=15, matcher=MatcherAVA(12, 5, 2, "ava", True, False), weight_dict=weight_dict(12),
                               eos_coef=0.1, losses=["labels", "boxes"], data_file="ava", evaluation=evaluation)
        pack(name, outs, targets, crit(model_out, targets), store)

    outs, targets = jhmdb_case(g)
    model_out = dict(outs[-1], aux_outputs=outs[:-1])
    crit = SetCriterion(10, 21, num_queries=10, matcher=MatcherUCF(1, 5, 2, "jhmdb", False, False), weight_dict=weight_dict(1),
                        eos_coef=0.1, losses=["labels", "boxes"], data_file="jhmdb", evaluation=True)
    pack("jhmdb", outs, targets, crit(model_out, targets), store)

    out = os.path.join(ROOT, "tests", "golden", "criterion.npz")
    np.savez_compressed(out, **store)
    print("wrote", out, os.path.getsize(out), "bytes;", sum(1 for k in store if ".loss." in k), "loss entries")


if __name__ == "__main__":
    main()
