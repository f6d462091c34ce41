"""CPU: weight ingest against the reference's own loaders.

tests/golden/ckpt_hashes.json holds a SHA-1 of every tensor after the UNMODIFIED reference ``load_weights`` (Caffe2 ``.mat``) and
``load_detr_weights`` ran on the seeded synthetic files of oracle/synth_ckpt.py (generator: oracle/make_golden_ckpt.py); here the
same files go through tuber_b200.utils.checkpoint into this package's model and must give bit-identical tensors."""
import hashlib
import json
import os

import pytest
import scipy.io as sio
import torch

import tuber_b200
from oracle import synth_ckpt
from oracle.cases import load_case_cfg
from tuber_b200.utils import checkpoint as CK

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ckpt_hashes.json")))


def _sha(t):
    return hashlib.sha1(t.detach().to(torch.float32).contiguous().numpy().tobytes()).hexdigest()


@pytest.mark.parametrize("case,blocks", [("A_csn50", (3, 4, 6, 3)), ("C_small", (3, 8, 36, 3))])
def test_csn_mat_ingest_matches_reference_loader(tmp_path, case, blocks):
    model, _, _ = tuber_b200.build_model(load_case_cfg(case))
    path = str(tmp_path / "csn.mat")
    sio.savemat(path, synth_ckpt.csn_mat_arrays(blocks, seed=7))
    CK.load_csn_mat(model, path)
    sd = model.state_dict()
    gold = GOLD[case + "/mat"]
    assert set(gold) <= set(sd)
    bad = [k for k, h in gold.items() if _sha(sd[k]) != h]
    assert not bad, bad[:5]


def test_detr_partial_init_matches_reference_loader(tmp_path):
    cfg = load_case_cfg("A_csn50")
    model, _, _ = tuber_b200.build_model(cfg)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    path = str(tmp_path / "detr.pth")
    torch.save(synth_ckpt.detr_checkpoint(model.state_dict(), seed=11), path)
    info = CK.load_detr_weights(model, path, cfg)
    sd = model.state_dict()
    gold = GOLD["A_csn50/detr"]
    bad = [k for k, h in gold.items() if _sha(sd[k]) != h]
    assert not bad, bad[:5]
    assert tuple(sd["query_embed.weight"].shape) == (cfg.CONFIG.MODEL.QUERY_NUM, 256)            # truncated from 100 queries
    assert any("layers.9" in k for k in info["unused"])                                          # names absent from the model are skipped
    untouched = [k for k in sd if not k.startswith(("transformer.", "bbox_embed.", "query_embed."))]
    assert all(torch.equal(sd[k], before[k]) for k in untouched)


def test_tuber_checkpoint_intersection(tmp_path):
    cfg = load_case_cfg("A_csn50")
    src, _, _ = tuber_b200.build_model(cfg)
    dst, _, _ = tuber_b200.build_model(cfg)
    ck = {"model": {"module." + k: torch.randn_like(v) if v.is_floating_point() else v for k, v in src.state_dict().items()},
          "epoch": 3}
    ck["model"]["module.fc.weight"] = torch.zeros(3)                                             # a name the model does not have
    path = str(tmp_path / "tuber.pth")
    torch.save(ck, path)
    info = CK.load_model(dst, path)
    assert info["unused"] == ["module.fc.weight"] and not info["not_found"]
    for k, v in dst.state_dict().items():
        assert torch.equal(v, ck["model"]["module." + k]), k
