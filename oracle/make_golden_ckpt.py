"""Generate tests/golden/ckpt_hashes.json with the UNMODIFIED reference weight loaders (build container only).

TEST INFRASTRUCTURE.  Builds the reference model, runs the reference's own ``load_weights`` (Caffe2 ``.mat`` backbone,
models/backbones/ir_CSN_{50,152}.py) and ``load_detr_weights`` (utils/model_utils.py:10-36, on a ``DataParallel``-wrapped model as
``deploy_model`` does) on the seeded synthetic files of oracle/synth_ckpt.py, and stores a SHA-1 of every resulting tensor.

    python oracle/make_golden_ckpt.py
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import json
import os
import sys
import tempfile

import scipy.io as sio
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("TUBER_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import synth_ckpt                  # noqa: E402
from oracle.cases import load_case_cfg         # noqa: E402


def sha(t: torch.Tensor) -> str:
    return hashlib.sha1(t.detach().to(torch.float32).contiguous().numpy().tobytes()).hexdigest()


def main():
    sys.path.insert(0, REF)
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        from models.tuber_ava import build_model as ref_build_model
        from utils.model_utils import load_detr_weights as ref_load_detr
        import models.backbones.ir_CSN_50 as csn50
        import models.backbones.ir_CSN_152 as csn152
    for case, mod, blocks in (("A_csn50", csn50, (3, 4, 6, 3)), ("C_small", csn152, (3, 8, 36, 3))):
        cfg = load_case_cfg(case)
        with contextlib.redirect_stdout(io.StringIO()):
            model, _, _ = ref_build_model(cfg)
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "csn.mat")
            sio.savemat(path, synth_ckpt.csn_mat_arrays(blocks, seed=7))
            with contextlib.redirect_stdout(io.StringIO()):
                mod.load_weights(model.backbone.body, pretrain_path=path, load_fc=False, use_affine=False, tune_point=4)
        out[case + "/mat"] = {k: sha(v) for k, v in model.state_dict().items() if k.startswith("backbone.body.") and "out_fc" not in k
                              and "num_batches_tracked" not in k}
        if case == "A_csn50":
            wrapped = torch.nn.DataParallel(model)                       # 'module.' names, as after deploy_model
            with tempfile.TemporaryDirectory() as d:
                path = os.path.join(d, "detr.pth")
                torch.save(synth_ckpt.detr_checkpoint(model.state_dict(), seed=11), path)
                with contextlib.redirect_stdout(io.StringIO()):
                    ref_load_detr(wrapped, path, cfg)
            out[case + "/detr"] = {k: sha(v) for k, v in model.state_dict().items()
                                   if k.startswith(("transformer.", "bbox_embed.", "query_embed."))}
    dst = os.path.join(ROOT, "tests", "golden", "ckpt_hashes.json")
    json.dump(out, open(dst, "w"), indent=0, sort_keys=True)
    print("wrote", dst, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
