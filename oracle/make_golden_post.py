"""Generate tests/golden/postprocess.npz from the UNMODIFIED reference post-processors (build container only).

TEST INFRASTRUCTURE.  Imports /root/reference/models/criterion.py as-is and runs ``PostProcessAVA`` / ``PostProcess``
(criterion.py:413-482) on seeded random model outputs, then formats the AVA rows exactly as the reference's evaluation loop
writes them (utils/video_action_recognition.py:411-415: ``"{} {}\\n".format(id, np.concatenate([box, scores, p]).tolist())``)
and parses them back the way evaluates/evaluate_ava.py:101-130 does.  Stored: inputs, the three outputs of each
post-processor, the text lines and the parsed numbers.

    python oracle/make_golden_post.py
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


def main():
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        from models.criterion import PostProcess, PostProcessAVA     # the reference's own classes
    g = torch.Generator().manual_seed(123)
    B, Q, C = 3, 15, 80
    logits = torch.randn(B, Q, C, generator=g) * 2
    boxes = torch.rand(B, Q, 4, generator=g) * 0.5 + 0.2
    logits_b = torch.randn(B, Q, 3, generator=g) * 3           # spread so that the 0.8 gate opens for some queries
    sizes = torch.tensor([[240.0, 320.0], [256.0, 341.0], [360.0, 480.0]])
    s_ava, b_ava, p_ava = PostProcessAVA()({"pred_logits": logits, "pred_boxes": boxes, "pred_logits_b": logits_b}, sizes)
    # the loop's file format (video_action_recognition.py:411-415) and the evaluator's parser (evaluate_ava.py:108-112)
    ids = ["vid%02d_%04d" % (b, 900 + b) for b in range(B)]
    lines, parsed = [], []
    for b in range(B):
        for q in range(Q):
            data = np.concatenate([b_ava[b, q], s_ava[b, q], p_ava[b, q]])
            lines.append("{} {}\n".format(ids[b], data.tolist()))
    for line in lines:
        vals = line.split(' [')[1].split(']')[0].split(',')
        parsed.append([float(x) for x in vals])
    # JHMDB-style post-processor: softmax scores over 22 classes, per-clip 2-way logits
    Cj, Qj = 22, 40
    logits_j = torch.randn(B, Qj, Cj, generator=g)
    boxes_j = torch.rand(B, Qj, 4, generator=g) * 0.5 + 0.2
    logits_bj = torch.randn(B, 2, generator=g)
    s_j, b_j, p_j = PostProcess()({"pred_logits": logits_j, "pred_boxes": boxes_j, "pred_logits_b": logits_bj}, sizes)
    out = os.path.join(ROOT, "tests", "golden", "postprocess.npz")
    np.savez_compressed(out, logits=logits.numpy(), boxes=boxes.numpy(), logits_b=logits_b.numpy(), sizes=sizes.numpy(),
                        scores_ava=s_ava, boxes_ava=b_ava, p_ava=p_ava, ids=np.array(ids), lines=np.array(lines),
                        parsed=np.array(parsed, dtype=np.float64),
                        logits_j=logits_j.numpy(), boxes_j=boxes_j.numpy(), logits_bj=logits_bj.numpy(), scores_j=s_j, boxes_jo=b_j,
                        p_j=p_j)
    print("wrote", out, os.path.getsize(out), "bytes;", int((p_ava > 0.8).sum()), "of", p_ava.size, "queries pass the gate")


if __name__ == "__main__":
    main()
