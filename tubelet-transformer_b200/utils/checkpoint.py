"""Weight ingest for the TubeR forward path: host-side plumbing only.

Three sources, each mirroring the reference's loader so that released files work unchanged:
  * load_csn_mat      Caffe2 ir-CSN ``.mat`` backbone weights  (models/backbones/ir_CSN_152.py:213-318, ir_CSN_50.py:213-320)
  * load_detr_weights DETR-COCO ``detr.pth`` partial initialisation (utils/model_utils.py:10-36)
  * load_model        TubeR ``.pth`` checkpoints, intersection by parameter name (utils/model_utils.py:66-95)
All three only fill the module's ``state_dict`` (names / shapes are the reference's); the device plan is rebuilt from it on
the next forward (``DETR.load_state_dict`` drops the packed copy).
"""
from __future__ import annotations

from typing import Dict, Iterable, Tuple

import numpy as np
import torch

_BN = (("weight", "_s"), ("bias", "_b"), ("running_mean", "_rm"), ("running_var", "_riv"))   # ir_CSN_152.py:229-241 (use_affine=False)


def _csn_mat_names(blocks: Iterable[int]) -> Iterable[Tuple[str, str]]:
    """(state_dict name under backbone.body, .mat array name) for every tensor the reference copies (load_fc=False)."""
    yield "conv1.weight", "conv1_w"
    for ours, theirs in _BN:
        yield f"bn1.{ours}", "conv1_spatbn_relu" + theirs
    count = 0                                         # blocks are numbered consecutively: start_count = cumulative block counts
    for li, nb in enumerate(blocks):
        for bi in range(nb):
            p = f"layer{li + 1}.{bi}"
            for conv in (1, 3, 4):
                yield f"{p}.conv{conv}.weight", f"comp_{count}_conv_{conv}_w"
                for ours, theirs in _BN:
                    yield f"{p}.bn{conv}.{ours}", f"comp_{count}_spatbn_{conv}{theirs}"
            if bi == 0:                               # down_sample = (conv, bn) on the first block of every stage
                yield f"{p}.down_sample.0.weight", f"shortcut_projection_{count}_w"
                for ours, theirs in _BN:
                    yield f"{p}.down_sample.1.{ours}", f"shortcut_projection_{count}_spatbn{theirs}"
            count += 1


def csn_mat_to_state_dict(mat: Dict[str, np.ndarray], blocks: Iterable[int], shapes: Dict[str, torch.Size],
                          prefix: str = "backbone.body.") -> Dict[str, torch.Tensor]:
    """Arrays of a loaded ``.mat`` -> ``{state_dict name: tensor}`` (conv weights keep their shape, BN vectors are flattened)."""
    out = {}
    for ours, theirs in _csn_mat_names(blocks):
        if theirs not in mat:
            raise KeyError(f"'{theirs}' is missing from the .mat file")
        name = prefix + ours
        t = torch.from_numpy(np.ascontiguousarray(mat[theirs]))
        if ".bn" in ours or "down_sample.1" in ours or ours.startswith("bn1"):
            t = t.reshape(-1)
        if name not in shapes:
            raise KeyError(f"the model has no parameter '{name}'")
        if tuple(t.shape) != tuple(shapes[name]):
            raise ValueError(f"{theirs}: shape {tuple(t.shape)} does not match {name} {tuple(shapes[name])}")
        out[name] = t.to(torch.float32)
    return out


def load_csn_mat(model: torch.nn.Module, path: str) -> None:
    """Fill ``backbone.body.*`` from a Caffe2 ir-CSN ``.mat`` file (the reference's ``load_weights(..., load_fc=False)``)."""
    import scipy.io as sio
    sd = model.state_dict()
    blocks = [sum(1 for k in sd if k.startswith(f"backbone.body.layer{i}.") and k.endswith(".conv1.weight")) for i in (1, 2, 3, 4)]
    new = csn_mat_to_state_dict(sio.loadmat(path), blocks, {k: v.shape for k, v in sd.items()})
    sd.update(new)
    model.load_state_dict(sd)


def _strip_first(k: str) -> str:
    return k.split(".", 1)[1] if "." in k else k


def load_detr_weights(model: torch.nn.Module, checkpoint, cfg) -> Dict[str, list]:
    """DETR-COCO initialisation (utils/model_utils.py:10-36): copy ``checkpoint['model']`` entries whose SECOND path component is
    ``transformer`` / ``bbox_embed`` / ``query_embed`` (the file carries a wrapper prefix such as ``module.``), truncating
    ``query_embed`` to this model's query count, when the name exists in the model.  Works with and without the same wrapper
    prefix on the model's own names.  Returns the used / unused checkpoint keys."""
    if isinstance(checkpoint, str):
        checkpoint = torch.load(checkpoint, map_location="cpu")
    sd = model.state_dict()
    m = cfg.CONFIG.MODEL
    query_size = m.QUERY_NUM if m.SINGLE_FRAME else m.QUERY_NUM * (m.TEMP_LEN // m.DS_RATE)
    picked = {}
    for k, v in checkpoint["model"].items():
        parts = k.split(".")
        if len(parts) < 2:
            continue
        if parts[1] in ("transformer", "bbox_embed"):
            picked[k] = v
        elif parts[1] == "query_embed":
            picked[k] = v[:query_size]
    used, unused = [], []
    for k, v in picked.items():
        name = k if k in sd else _strip_first(k)
        if name in sd:
            sd[name] = v
            used.append(k)
        else:
            unused.append(k)
    model.load_state_dict(sd)
    return {"used": used, "unused": unused}


def load_model(model: torch.nn.Module, checkpoint) -> Dict[str, list]:
    """TubeR checkpoint (utils/model_utils.py:66-95): intersection of ``checkpoint['model']`` with the model's names; a leading
    ``module.`` (DistributedDataParallel) on either side is tolerated.  Returns used / unused / not-found names."""
    if isinstance(checkpoint, str):
        checkpoint = torch.load(checkpoint, map_location="cpu")
    sd = model.state_dict()
    used, unused = [], []
    for k, v in checkpoint["model"].items():
        name = k if k in sd else (k[len("module."):] if k.startswith("module.") else "module." + k)
        if name in sd:
            sd[name] = v
            used.append(name)
        else:
            unused.append(k)
    model.load_state_dict(sd)
    return {"used": used, "unused": unused, "not_found": [k for k in sd if k not in used]}
