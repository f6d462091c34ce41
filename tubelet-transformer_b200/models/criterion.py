"""Evaluation-time criterion of TubeR: the object ``build_model`` returns as its second value.

The reference's evaluation loop calls ``criterion(outputs, targets)`` on every batch and reads
``loss_bbox / loss_giou / loss_ce / class_error / loss_ce_b`` and ``criterion.weight_dict`` from the result
(utils/video_action_recognition.py:305-307,352-374), so a drop-in model needs a criterion that answers.  This module
restates, for the loop's use, what the reference computes there:

* matching  -- models/detr/matcher.py:37-80 (AVA: cost = COST_BBOX * L1 + COST_CLASS * (-p_actor) + COST_GIOU * (-GIoU), one linear
  assignment per clip) and models/detr/matcher_ucf.py (JHMDB / UCF: the class term is -softmax(logits)[label]);
* losses    -- models/criterion.py:41-84,100-121,171-209 (``SetCriterionAVA``) and :246-271,287-322,370-410 (``SetCriterion``).

It is host-side tensor plumbing over a handful of (B, Q, .) outputs -- no hot-path arithmetic; gradients flow through it
like through any torch code, but training (optimiser, schedules, data augmentation) stays with the reference.
``tests/test_criterion_cpu.py`` holds every entry of the loss dictionary to the reference's own classes
(``oracle/make_golden_criterion.py`` -> ``tests/golden/criterion.npz``).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn


def _corners(b: Tensor) -> Tensor:
    """(cx, cy, w, h) -> (x0, y0, x1, y1)  (models/transformer/util/box_ops.py:9-13)."""
    c, s = b[..., :2], b[..., 2:]
    return torch.cat((c - 0.5 * s, c + 0.5 * s), dim=-1)


def _giou(a: Tensor, b: Tensor) -> Tensor:
    """Generalised IoU of corner boxes, broadcasting a (..., 4) against b (..., 4)  (box_ops.py:24-60)."""
    area_a = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    area_b = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    iw = (torch.min(a[..., 2], b[..., 2]) - torch.max(a[..., 0], b[..., 0])).clamp(min=0)
    ih = (torch.min(a[..., 3], b[..., 3]) - torch.max(a[..., 1], b[..., 1])).clamp(min=0)
    inter = iw * ih
    union = area_a + area_b - inter
    hull = ((torch.max(a[..., 2], b[..., 2]) - torch.min(a[..., 0], b[..., 0])).clamp(min=0)
            * (torch.max(a[..., 3], b[..., 3]) - torch.min(a[..., 1], b[..., 1])).clamp(min=0))
    return inter / union - (hull - union) / hull


class HungarianMatcher(nn.Module):
    """One optimal query <-> ground-truth assignment per clip (models/detr/matcher.py, matcher_ucf.py)."""

    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1, data_file: str = "ava",
                 binary_loss: bool = False, before: bool = False):
        super().__init__()
        if cost_class == 0 and cost_bbox == 0 and cost_giou == 0:
            raise ValueError("all matching costs are zero")
        self.cost_class, self.cost_bbox, self.cost_giou = cost_class, cost_bbox, cost_giou
        self.data_file, self.binary_loss, self.before = data_file, binary_loss, before

    @torch.no_grad()
    def forward(self, outputs: Dict[str, Tensor], targets: Sequence[Dict[str, Tensor]]) -> List[Tuple[Tensor, Tensor]]:
        from scipy.optimize import linear_sum_assignment
        boxes = outputs["pred_boxes"]                                  # (B, Q, 4)
        B, Q = boxes.shape[:2]
        if self.data_file == "ava":
            p_actor = outputs["pred_logits_b"].softmax(-1)[..., 1]     # (B, Q)
        else:
            prob = outputs["pred_logits"].softmax(-1)                  # (B, Q, C+1)
        pairs = []
        for b, t in enumerate(targets):
            gt = t["boxes"][:, 1:]                                     # column 0 is the frame position
            if gt.shape[0] == 0:
                e = torch.empty(0, dtype=torch.int64)
                pairs.append((e, e.clone()))
                continue
            l1 = torch.cdist(boxes[b], gt, p=1)
            g = _giou(_corners(boxes[b])[:, None, :], _corners(gt)[None, :, :])
            if self.data_file == "ava":
                cls = -p_actor[b][:, None].expand(Q, gt.shape[0])
            else:
                cls = -prob[b][:, t["labels"]]
            cost = self.cost_bbox * l1 + self.cost_class * cls - self.cost_giou * g
            rows, cols = linear_sum_assignment(cost.cpu())
            pairs.append((torch.as_tensor(rows, dtype=torch.int64), torch.as_tensor(cols, dtype=torch.int64)))
        return pairs


class _SetCriterionBase(nn.Module):
    """Shared driver: match the last layer, evaluate the losses, repeat for every auxiliary layer with a ``_{i}`` suffix
    (models/criterion.py:171-209)."""

    def __init__(self, weight, num_classes, num_queries, matcher, weight_dict, eos_coef, losses, data_file, evaluation=False):
        super().__init__()
        self.weight, self.num_classes, self.num_queries = weight, num_classes, num_queries
        self.matcher, self.weight_dict, self.eos_coef = matcher, weight_dict, eos_coef
        self.losses, self.data_file, self.evaluation = list(losses), data_file, evaluation

    # -- hooks ------------------------------------------------------------------------------------
    def _select(self, layer_out: Dict[str, Tensor], targets) -> Dict[str, Tensor]:
        return {k: v for k, v in layer_out.items() if k != "aux_outputs"}

    def _labels(self, out, targets, batch_idx, query_idx, matched, log: bool) -> Dict[str, Tensor]:
        raise NotImplementedError

    # -- shared -----------------------------------------------------------------------------------
    def _boxes(self, out, targets, pairs, batch_idx, query_idx, num_boxes: Tensor) -> Dict[str, Tensor]:
        src = out["pred_boxes"][batch_idx, query_idx]
        tgt = torch.cat([t["boxes"][j] for t, (_, j) in zip(targets, pairs)], dim=0)[:, 1:]
        l1 = (src - tgt).abs().sum() / num_boxes
        gi = (1 - _giou(_corners(src), _corners(tgt))).sum() / num_boxes
        return {"loss_bbox": l1, "loss_giou": gi}

    def _layer(self, out, targets, num_boxes, log: bool) -> Dict[str, Tensor]:
        pairs = self.matcher(out, targets)
        dev = out["pred_logits"].device
        batch_idx = torch.cat([torch.full_like(i, b) for b, (i, _) in enumerate(pairs)]).to(dev)
        query_idx = torch.cat([i for i, _ in pairs]).to(dev)
        res: Dict[str, Tensor] = {}
        for name in self.losses:
            if name == "labels":
                matched = torch.cat([t["labels"][j.to(t["labels"].device)] for t, (_, j) in zip(targets, pairs)])
                res.update(self._labels(out, targets, batch_idx, query_idx, matched, log))
            elif name == "boxes":
                res.update(self._boxes(out, targets, pairs, batch_idx, query_idx, num_boxes))
            else:
                raise ValueError(f"unsupported loss '{name}' (the reference's mask losses need segmentation heads TubeR does not have)")
        return res

    def forward(self, outputs: Dict[str, Tensor], targets: Sequence[Dict[str, Tensor]]) -> Dict[str, Tensor]:
        dev = outputs["pred_logits"].device
        num_boxes = torch.as_tensor([float(sum(len(t["labels"]) for t in targets))], dtype=torch.float, device=dev)
        losses = self._layer(self._select(outputs, targets), targets, num_boxes, log=True)
        for i, aux in enumerate(outputs.get("aux_outputs", ())):
            for k, v in self._layer(self._select(aux, targets), targets, num_boxes, log=False).items():
                losses[f"{k}_{i}"] = v
        return losses


class SetCriterionAVA(_SetCriterionBase):
    """AVA: multi-label sigmoid classification on matched queries + 3-way actor-ness (models/criterion.py:11-209)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        w = torch.ones(3)
        w[-1] = self.eos_coef
        self.register_buffer("empty_weight", w)

    def _labels(self, out, targets, batch_idx, query_idx, matched, log):
        logits, logits_b = out["pred_logits"], out["pred_logits_b"]
        actor = torch.full(logits_b.shape[:2], 2, dtype=torch.int64, device=logits.device)
        actor[batch_idx, query_idx] = 1
        loss_b = F.cross_entropy(logits_b.transpose(1, 2), actor, self.empty_weight.to(logits.device))
        want = torch.zeros_like(logits, dtype=torch.float32)
        want[batch_idx, query_idx] = matched.to(want)
        if self.evaluation:
            loss = F.binary_cross_entropy(logits.sigmoid(), want)
        else:
            w = torch.ones(logits.shape[:2], dtype=torch.float32, device=logits.device)
            w[batch_idx, query_idx] = self.weight
            loss = F.binary_cross_entropy(logits.sigmoid(), want, weight=w[:, :, None])
        res = {"loss_ce": loss, "loss_ce_b": loss_b}
        if log:
            res["class_error"] = 100 - _exact_set_accuracy(logits[batch_idx, query_idx], matched)
        return res


class SetCriterion(_SetCriterionBase):
    """JHMDB / UCF: single-label softmax classification (with a no-object class) on the key frame's queries + the per-clip
    2-way visibility head (models/criterion.py:212-410)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        w = torch.ones(self.num_classes + 1)
        w[-1] = self.eos_coef
        self.register_buffer("empty_weight", w)

    def _select(self, layer_out, targets):
        # the queries of the key frame: num_queries * key_pos + (0 .. num_queries-1)   (criterion.py:377-379)
        nq = self.num_queries
        dev = layer_out["pred_logits"].device
        first = torch.stack([t["key_pos"].reshape(()).to(dev).long() * nq for t in targets])
        pick = first[:, None] + torch.arange(nq, device=dev)[None, :]
        sel = {}
        for k, v in layer_out.items():
            if k == "aux_outputs":
                continue
            sel[k] = v.gather(1, pick[:, :, None].expand(-1, -1, v.shape[-1])) if k in ("pred_boxes", "pred_logits") else v
        return sel

    def _labels(self, out, targets, batch_idx, query_idx, matched, log):
        logits = out["pred_logits"]
        loss_b = F.cross_entropy(out["pred_logits_b"], torch.cat([t["vis"] for t in targets]).view(-1))
        want = torch.full(logits.shape[:2], self.num_classes, dtype=torch.int64, device=logits.device)
        want[batch_idx, query_idx] = matched
        res = {"loss_ce": F.cross_entropy(logits.transpose(1, 2), want, self.empty_weight.to(logits.device)), "loss_ce_b": loss_b}
        if log:
            hit = logits[batch_idx, query_idx]
            if matched.numel() == 0:
                res["class_error"] = 100 - torch.zeros([], device=logits.device)
            else:
                res["class_error"] = 100 - (hit.argmax(-1) == matched).float().sum() * (100.0 / matched.numel())
        return res

    def _boxes(self, out, targets, pairs, batch_idx, query_idx, num_boxes):
        if float(num_boxes) > 0:
            return super()._boxes(out, targets, pairs, batch_idx, query_idx, num_boxes)
        zero = torch.zeros(1, device=out["pred_boxes"].device)          # criterion.py:318-321: |1 - 1|
        return {"loss_bbox": zero, "loss_giou": zero.clone()}


def _exact_set_accuracy(logits: Tensor, labels: Tensor):
    """Percentage of matched queries whose top-k classes (k = number of positive labels) are exactly the positive set
    (utils/misc.py:498-519).  Returns a python float, or a zero tensor when nothing was matched, like the reference."""
    if labels.numel() == 0:
        return torch.zeros([], device=logits.device)
    hits = 0
    for row, lab in zip(logits, labels):
        pos = lab.nonzero().flatten()
        top = row.topk(len(pos), 0, True, True)[1]
        hits += int(set(pos.tolist()) == set(top.tolist()))
    return hits * (100.0 / labels.shape[0])


def build_criterion(cfg) -> nn.Module:
    """The matcher + criterion + weight dictionary of the reference's ``build_model`` (models/tuber_ava.py:184-216)."""
    c = cfg.CONFIG
    ava = c.DATA.DATASET_NAME == "ava"
    matcher = HungarianMatcher(cost_class=c.MATCHER.COST_CLASS, cost_bbox=c.MATCHER.COST_BBOX, cost_giou=c.MATCHER.COST_GIOU,
                               data_file=c.DATA.DATASET_NAME, binary_loss=c.MATCHER.BNY_LOSS, before=c.MATCHER.BEFORE)
    weights = {"loss_ce": c.LOSS_COFS.DICE_COF, "loss_bbox": c.LOSS_COFS.BBOX_COF, "loss_giou": c.LOSS_COFS.GIOU_COF, "loss_ce_b": 1}
    if c.TRAIN.AUX_LOSS:
        for i in range(c.MODEL.DEC_LAYERS - 1):
            weights.update({f"{k}_{i}": v for k, v in list(weights.items())[:4]})
    cls = SetCriterionAVA if ava else SetCriterion
    return cls(c.LOSS_COFS.WEIGHT, c.DATA.NUM_CLASSES, num_queries=c.MODEL.QUERY_NUM, matcher=matcher, weight_dict=weights,
               eos_coef=c.LOSS_COFS.EOS_COF, losses=["labels", "boxes"], data_file=c.DATA.DATASET_NAME,
               evaluation=bool(c.EVAL_ONLY))
