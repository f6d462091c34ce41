"""yacs-free config surface for the TubeR forward path.

The reference reads its settings from a yacs ``CfgNode`` built by
``get_cfg_defaults()`` (pipelines/video_action_recognition_config.py:220-222)
and ``cfg.merge_from_file(yaml)`` (eval_tuber_ava.py:55-56).  yacs is not a
dependency of this package; ``CfgNode`` below gives the same attribute access
(``cfg.CONFIG.MODEL.D_MODEL``), the same ``merge_from_file`` /
``merge_from_list`` / ``clone`` / ``dump`` calls, and the same string handling
(yacs literal-evals YAML strings, so ``LR: 1e-4`` -- a *string* for PyYAML --
becomes a float).  ``build_model`` accepts either this node or a real yacs one.

Only the keys the forward path reads get defaults here:
DATA.{DATASET_NAME,NUM_CLASSES,IMG_SIZE}, MODEL.{BACKBONE_NAME,
TEMPORAL_DS_STRATEGY,SINGLE_FRAME,LAST_STRIDE,TEMP_LEN,DS_RATE,D_MODEL,NHEAD,
ENC_LAYERS,DEC_LAYERS,DIM_FEEDFORWARD,DROPOUT,NORMALIZE_BEFORE,QUERY_NUM,
GENERATE_LFB,PRETRAINED,PRETRAIN_BACKBONE_DIR}, TRAIN.{AUX_LOSS,LR_BACKBONE},
EVAL_ONLY; everything else in a YAML is carried through untouched.
"""
from __future__ import annotations

import ast
import copy
import os
from typing import Any, Iterable

import yaml


class CfgNode(dict):
    """Attribute-style nested dict with the slice of the yacs API callers use."""

    def __init__(self, init: dict | None = None):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name: str) -> Any:
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name: str, value: Any) -> None:
        self[name] = value

    # -- yacs-compatible surface ------------------------------------------------
    def clone(self) -> "CfgNode":
        return copy.deepcopy(self)

    def merge_from_other_cfg(self, other: dict) -> None:
        _merge(self, other)

    def merge_from_file(self, path: str) -> None:
        with open(path, "r") as f:
            loaded = yaml.safe_load(f) or {}
        _merge(self, loaded)

    def merge_from_list(self, kv: Iterable[Any]) -> None:
        kv = list(kv)
        if len(kv) % 2:
            raise ValueError("merge_from_list expects KEY VALUE pairs")
        for key, value in zip(kv[0::2], kv[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                if p not in node:
                    node[p] = CfgNode()
                node = node[p]
            node[parts[-1]] = _coerce(value)

    def dump(self, **kw) -> str:
        return yaml.safe_dump(_plain(self), **kw)

    def freeze(self) -> None:  # yacs no-op stand-ins
        pass

    def defrost(self) -> None:
        pass


def _plain(node: Any) -> Any:
    if isinstance(node, dict):
        return {k: _plain(v) for k, v in node.items()}
    if isinstance(node, tuple):
        return list(node)
    return node


def _coerce(value: Any) -> Any:
    """yacs semantics: strings that parse as Python literals become literals."""
    if isinstance(value, str):
        try:
            return ast.literal_eval(value)
        except (ValueError, SyntaxError):
            return value
    return value


def _merge(dst: CfgNode, src: dict) -> None:
    for k, v in src.items():
        if isinstance(v, dict):
            if k not in dst or not isinstance(dst[k], dict):
                dst[k] = CfgNode()
            _merge(dst[k], v)
        else:
            dst[k] = _coerce(v)


_DEFAULTS = {
    "DDP_CONFIG": {
        "WORLD_SIZE": 1, "WORLD_RANK": 0, "GPU_WORLD_SIZE": 8, "GPU_WORLD_RANK": 0,
        "DIST_URL": "tcp://127.0.0.1:10001", "WOLRD_URLS": ["127.0.0.1"],
        "AUTO_RANK_MATCH": True, "DIST_BACKEND": "nccl", "GPU": 0, "DISTRIBUTED": True,
    },
    "CONFIG": {
        "EVAL_ONLY": False, "USE_LFB": False, "USE_LOCATION": False, "TWO_STREAM": False,
        "TRAIN": {"AUX_LOSS": True, "LR_BACKBONE": 1e-5, "BATCH_SIZE": 2},
        "VAL": {"BATCH_SIZE": 1, "FREQ": 1},
        "DATA": {"DATASET_NAME": "ava", "NUM_CLASSES": 80, "IMG_SIZE": 256, "TEMP_LEN": 32, "LABEL_PATH": "", "ANNO_PATH": "", "DATA_PATH": ""},
        "MODEL": {
            "SINGLE_FRAME": True, "BACKBONE_NAME": "CSN-152", "TEMPORAL_DS_STRATEGY": "avg",
            "LAST_STRIDE": False, "GENERATE_LFB": False,
            "ENC_LAYERS": 6, "DEC_LAYERS": 6, "D_MODEL": 256, "NHEAD": 8,
            "DIM_FEEDFORWARD": 2048, "QUERY_NUM": 15, "NORMALIZE_BEFORE": False,
            "DROPOUT": 0.1, "DS_RATE": 8, "TEMP_LEN": 32,
            "PRETRAINED": False, "PRETRAIN_BACKBONE_DIR": "", "PRETRAIN_TRANSFORMER_DIR": "",
            "PRETRAINED_PATH": "", "LOAD": False, "LOAD_FC": True,
        },
        "MATCHER": {"COST_CLASS": 12, "COST_BBOX": 5, "COST_GIOU": 2, "BNY_LOSS": True, "BEFORE": False},
        "LOSS_COFS": {"MASK_COF": 1, "DICE_COF": 12, "BBOX_COF": 5, "GIOU_COF": 2,
                      "EOS_COF": 0.1, "WEIGHT": 10, "WEIGHT_CHANGE": 1000,
                      "LOSS_CHANGE_COF": 2, "CLIPS_MAX_NORM": 0.1},
        "LOG": {"BASE_PATH": "", "RES_DIR": "tmp"},
    },
}

CONFIG_DIR = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "configs"))


def get_cfg_defaults() -> CfgNode:
    """Same role as pipelines/video_action_recognition_config.py:220-222."""
    return CfgNode(copy.deepcopy(_DEFAULTS))


def load_cfg(yaml_path: str | None = None, overrides: Iterable[Any] | None = None) -> CfgNode:
    """defaults <- YAML file <- ``["CONFIG.MODEL.QUERY_NUM", 4, ...]`` overrides.

    ``yaml_path`` may be a bare name (``TubeR_CSN152_AVA21.yaml``), resolved
    against this repo's ``configs/`` directory.
    """
    cfg = get_cfg_defaults()
    if yaml_path:
        if not os.path.exists(yaml_path):
            cand = os.path.join(CONFIG_DIR, yaml_path)
            if os.path.exists(cand):
                yaml_path = cand
        cfg.merge_from_file(yaml_path)
    if overrides:
        cfg.merge_from_list(overrides)
    return cfg
