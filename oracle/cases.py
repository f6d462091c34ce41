"""Named parity cases shared by oracle/make_golden.py and tests/.  TEST INFRASTRUCTURE.

Each case = (experiment YAML from configs/, key overrides, input shape(s), seeds, BN mode).
Sizes are chosen so the CPU oracle finishes in seconds.  Letters follow SURVEY.md section 8:
A = BASELINE configs[0] (CSN-50, 2+2 layers, 4 queries, one 8x128x128 clip), B = CSN50/decode,
C = CSN152/avg, D = CSN152/decode (the released part of config D), E = JHMDB (centre slice,
320 queries, 2-way head on pooled features).
"""
from __future__ import annotations

import os
import sys
from typing import Optional, Tuple

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import tuber_oracle as O  # noqa: E402

_A = ["CONFIG.MODEL.ENC_LAYERS", 2, "CONFIG.MODEL.DEC_LAYERS", 2, "CONFIG.MODEL.QUERY_NUM", 4,
      "CONFIG.MODEL.TEMP_LEN", 8]

CASES = {
    # name: yaml, overrides, list of (T,H,W) per clip, weight seed, clip seed, bn mode
    "A_csn50": dict(yaml="TubeR_CSN50_AVA21.yaml", over=_A, clips=[(8, 128, 128)], wseed=0, cseed=2, bn="identity"),
    "A_csn50_avg_bnrand": dict(yaml="TubeR_CSN50_AVA21.yaml", over=_A + ["CONFIG.MODEL.TEMPORAL_DS_STRATEGY", "avg"],
                               clips=[(8, 128, 128)], wseed=1, cseed=3, bn="random"),
    "B_small": dict(yaml="TubeR_CSN50_AVA21.yaml", over=[], clips=[(32, 64, 64)] * 2, wseed=0, cseed=2, bn="random"),
    "C_small": dict(yaml="TubeR_CSN152_AVA21.yaml", over=[], clips=[(32, 64, 96)] * 2, wseed=0, cseed=2, bn="random"),
    "C_ragged": dict(yaml="TubeR_CSN152_AVA21.yaml", over=[], clips=[(32, 96, 128), (32, 96, 80), (32, 64, 128)],
                     wseed=4, cseed=5, bn="random"),
    "C_mid": dict(yaml="TubeR_CSN152_AVA21.yaml", over=[], clips=[(32, 128, 128)], wseed=0, cseed=2, bn="identity"),
    "D_small": dict(yaml="TubeR_CSN152_AVA22.yaml", over=[], clips=[(32, 64, 64)], wseed=6, cseed=7, bn="random"),
    "E_small": dict(yaml="Tuber_CSN152_JHMDB.yaml", over=[], clips=[(16, 64, 64)] * 2, wseed=8, cseed=9, bn="random"),
    # the sizes the reference's evaluation transform really produces are not multiples of anything: Resize_Custom keeps the aspect
    # ratio (datasets/video_transforms.py:213-228: 256 x int(256 * w / h) = 256x341 for 4:3, 256x455 for 16:9; JHMDB 224x298),
    # and a batch mixes videos of different widths.  Scaled down by 4 here; the full sizes run against the live oracle on the GPU box
    "B_odd": dict(yaml="TubeR_CSN50_AVA21.yaml", over=[], clips=[(32, 64, 85), (32, 64, 113)], wseed=12, cseed=13, bn="random"),
    "C_odd": dict(yaml="TubeR_CSN152_AVA21.yaml", over=[], clips=[(32, 64, 113), (32, 75, 100)], wseed=14, cseed=15, bn="random"),
    "E_odd": dict(yaml="Tuber_CSN152_JHMDB.yaml", over=[], clips=[(16, 56, 75)], wseed=16, cseed=17, bn="random"),
    # conv rows wider than 128 outputs (W > 256): the stem takes its single-CTA kernel + separate max pool
    "B_wide": dict(yaml="TubeR_CSN50_AVA21.yaml", over=[], clips=[(8, 40, 341)], wseed=18, cseed=19, bn="random"),
    "C_long": dict(yaml="TubeR_CSN152_AVA21.yaml", over=[], clips=[(64, 64, 64)], wseed=10, cseed=11, bn="random"),
}


# Configuration branches no shipped YAML selects (max pool, LAST_STRIDE, SINGLE_FRAME: False): the oracle is pinned on them too
# (tests/test_oracle_golden.py); they are not part of the GPU parity parametrisation above.
CPU_ONLY_CASES = {
    "X_max": dict(yaml="TubeR_CSN50_AVA21.yaml", over=_A + ["CONFIG.MODEL.TEMPORAL_DS_STRATEGY", "max", "CONFIG.MODEL.TEMP_LEN", 32],
                  clips=[(32, 64, 64)], wseed=20, cseed=21, bn="random"),
    "X_last_stride": dict(yaml="TubeR_CSN50_AVA21.yaml", over=_A + ["CONFIG.MODEL.LAST_STRIDE", True], clips=[(8, 128, 96)],
                          wseed=22, cseed=23, bn="random"),
    "X_all_frames": dict(yaml="TubeR_CSN50_AVA21.yaml", over=_A + ["CONFIG.MODEL.SINGLE_FRAME", False], clips=[(16, 64, 64)] * 2,
                         wseed=24, cseed=25, bn="random"),
}
CASES_ALL = dict(CASES, **CPU_ONLY_CASES)


def load_case_cfg(name: str):
    import tuber_b200  # config loader only (plain YAML plumbing)
    c = CASES_ALL[name]
    return tuber_b200.load_cfg(c["yaml"], c["over"])


def build_case(name: str) -> Tuple[object, dict, torch.Tensor, Optional[torch.Tensor]]:
    """-> (cfg, state_dict, clips (B,3,T,H,W), mask (B,H,W) or None)."""
    c = CASES_ALL[name]
    cfg = load_case_cfg(name)
    sd = O.make_state_dict(cfg, seed=c["wseed"], bn=c["bn"])
    shapes = c["clips"]
    if all(s == shapes[0] for s in shapes):
        t, h, w = shapes[0]
        return cfg, sd, O.make_clips(len(shapes), t, h, w, seed=c["cseed"]), None
    clips = [O.make_clips(1, t, h, w, seed=c["cseed"] + i)[0] for i, (t, h, w) in enumerate(shapes)]
    batch, mask = O.pad_clips(clips)
    return cfg, sd, batch, mask
