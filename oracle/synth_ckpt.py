"""Synthetic checkpoint files for the weight-ingest tests.  TEST INFRASTRUCTURE.

Seeded (numpy PCG64) stand-ins for the files the reference downloads: a Caffe2 ir-CSN ``.mat`` with the array names / shapes
``load_weights`` reads (models/backbones/ir_CSN_152.py:242-318) and a DETR-COCO style ``detr.pth`` (utils/model_utils.py:10-36).
"""
from __future__ import annotations

import numpy as np
import torch


def csn_mat_arrays(blocks, seed: int = 0):
    rng = np.random.Generator(np.random.PCG64(seed))
    planes = [64, 128, 256, 512]
    out = {}

    def arr(name, *shape):
        out[name] = rng.standard_normal(shape, dtype=np.float32)

    def bn(name, c):
        for suf in ("_s", "_b", "_rm"):
            arr(name + suf, 1, c)                   # scipy stores 1-D arrays as (1, C)
        out[name + "_riv"] = (rng.random((1, c), dtype=np.float32) + 0.5)

    arr("conv1_w", 64, 3, 3, 7, 7)
    bn("conv1_spatbn_relu", 64)
    count, cin = 0, 64
    for li, nb in enumerate(blocks):
        p, cout = planes[li], planes[li] * 4
        for bi in range(nb):
            arr(f"comp_{count}_conv_1_w", p, cin if bi == 0 else cout, 1, 1, 1)
            arr(f"comp_{count}_conv_3_w", p, 1, 3, 3, 3)
            arr(f"comp_{count}_conv_4_w", cout, p, 1, 1, 1)
            for k, c in ((1, p), (3, p), (4, cout)):
                bn(f"comp_{count}_spatbn_{k}", c)
            if bi == 0:
                arr(f"shortcut_projection_{count}_w", cout, cin, 1, 1, 1)
                bn(f"shortcut_projection_{count}_spatbn", cout)
            count += 1
        cin = cout
    arr("last_out_L400_w", 400, 2048)               # present in the released files, not loaded (load_fc=False)
    arr("last_out_L400_b", 1, 400)
    return out


def detr_checkpoint(model_state_dict, seed: int = 1, queries: int = 100, prefix: str = "module."):
    """A ``detr.pth``-like dict: every transformer / bbox_embed tensor of the model plus a 100-query ``query_embed`` and a few
    tensors the loader must ignore, all under the wrapper prefix the reference's key test expects (``k.split('.')[1]``)."""
    g = torch.Generator().manual_seed(seed)
    ck = {}
    for k, v in model_state_dict.items():
        if k.startswith("transformer.") or k.startswith("bbox_embed."):
            ck[prefix + k] = torch.randn(v.shape, generator=g)
    ck[prefix + "query_embed.weight"] = torch.randn(queries, 256, generator=g)
    ck[prefix + "class_embed.weight"] = torch.randn(92, 256, generator=g)          # COCO head: not part of TubeR
    ck[prefix + "backbone.0.body.conv1.weight"] = torch.randn(64, 3, 7, 7, generator=g)
    ck[prefix + "transformer.decoder.layers.9.linear1.weight"] = torch.randn(8, 8, generator=g)   # name absent from the model
    return {"model": ck}
