"""CPU: the oracle restatement vs. the committed outputs of the unmodified reference.

tests/golden/<case>.npz were produced by oracle/make_golden.py, which runs
/root/reference's own build_model(cfg)/forward on the same seeded weights and clips.
"""
import os

import numpy as np
import pytest
import torch

from oracle import tuber_oracle as O
from oracle.cases import CASES_ALL, build_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-5   # fp32 CPU vs fp32 CPU, different op grouping only


@pytest.mark.parametrize("name", list(CASES_ALL))
def test_oracle_matches_reference_golden(name):
    cfg, sd, clips, mask = build_case(name)
    g = np.load(os.path.join(GOLD, name + ".npz"))
    taps = {}
    out = O.forward(cfg, sd, clips, mask, taps)
    for key in ("pred_logits", "pred_boxes", "pred_logits_b"):
        ref = torch.from_numpy(g[key])
        assert tuple(out[key].shape) == tuple(ref.shape), key
        emax, el2 = O.rel_err(out[key], ref)
        assert emax <= TOL and el2 <= TOL, (name, key, emax, el2)
    xt = taps["xt"]
    assert tuple(xt.shape) == tuple(g["xt_shape"])
    probe = xt.flatten()[torch.from_numpy(g["xt_probe_idx"])]
    assert O.rel_err(probe, torch.from_numpy(g["xt_probe"]))[0] <= TOL
    assert O.rel_err(taps["hs"][-1], torch.from_numpy(g["hs_last"]))[1] <= TOL
    assert O.rel_err(taps["mem_c"][:, :16], torch.from_numpy(g["mem_c_first"]))[1] <= TOL


def test_param_spec_counts():
    """560 tensors for CSN50/decode (SURVEY.md section 8b probe of the live reference)."""
    from oracle.cases import load_case_cfg
    cfg = load_case_cfg("B_small")
    assert len(O.param_spec(cfg)) == 560


def test_reference_dict_shape():
    cfg, sd, clips, mask = build_case("A_csn50")
    out = O.as_reference_dict(O.forward(cfg, sd, clips, mask))
    assert out["pred_logits"].shape == (1, 4, 80)
    assert out["pred_boxes"].shape == (1, 4, 4)
    assert out["pred_logits_b"].shape == (1, 4, 3)
    assert len(out["aux_outputs"]) == 1


def test_postprocess_ava_gate():
    logits = torch.zeros(1, 2, 80)
    boxes = torch.tensor([[[0.5, 0.5, 0.2, 0.4], [0.25, 0.25, 0.5, 0.5]]])
    logits_b = torch.tensor([[[0.0, 5.0, 0.0], [0.0, 0.0, 0.0]]])
    scores, xyxy, pb = O.postprocess_ava(logits, boxes, logits_b, torch.tensor([[100.0, 200.0]]))
    assert scores[0, 1].abs().max() == 0            # p_actor = 1/3 < 0.8 -> gated to zero
    assert torch.allclose(scores[0, 0], 0.5 * pb[0, 0])
    assert torch.allclose(xyxy[0, 0], torch.tensor([80.0, 30.0, 120.0, 70.0]))


def test_postprocess_oracle_matches_reference_golden():
    """oracle.postprocess_ava against the reference's own PostProcessAVA (tests/golden/postprocess.npz, oracle/make_golden_post.py)."""
    g = np.load(os.path.join(GOLD, "postprocess.npz"))
    scores, xyxy, pb = O.postprocess_ava(torch.from_numpy(g["logits"]), torch.from_numpy(g["boxes"]), torch.from_numpy(g["logits_b"]),
                                         torch.from_numpy(g["sizes"]))
    assert np.abs(scores.numpy() - g["scores_ava"]).max() < 1e-6
    assert np.abs(xyxy.numpy() - g["boxes_ava"]).max() < 1e-4
    assert np.abs(pb.numpy() - g["p_ava"]).max() < 1e-6


def test_detection_line_format_is_the_reference_file_format():
    """format_detection_lines reproduces the loop's text lines (video_action_recognition.py:411-415) byte for byte and they
    parse back the way evaluates/evaluate_ava.py:108-112 reads them."""
    import tuber_b200
    g = np.load(os.path.join(GOLD, "postprocess.npz"))
    rows = np.concatenate([g["boxes_ava"], g["scores_ava"], g["p_ava"]], axis=-1)          # (B, Q, 4 + C + 1)
    lines = tuber_b200.format_detection_lines([str(i) for i in g["ids"]], rows)
    assert lines == [str(l) for l in g["lines"]]
    parsed = [[float(x) for x in line.split(' [')[1].split(']')[0].split(',')] for line in lines]
    assert np.array_equal(np.array(parsed), g["parsed"])


# ---- uint8 input transform ---------------------------------------------------------
def test_input_transform_oracle_and_value_table_match_reference_golden():
    """oracle.frames_to_clips and the library's host-side value table (tuber_input_lut, no device needed) against the
    reference's own ToTensor + Normalize + stack + permute (tests/golden/input_u8.npz, oracle/make_golden_input.py): bit-exact."""
    from tuber_b200 import _lib
    g = np.load(os.path.join(GOLD, "input_u8.npz"))
    for mk, sk, ck in (("mean", "std", "clips"), ("mean2", "std2", "clips2")):
        mean, std = [float(v) for v in g[mk]], [float(v) for v in g[sk]]
        lut = np.asarray(_lib.input_lut(mean, std), dtype=np.float32).reshape(3, 256)
        for name in ("a", "b", "table"):
            fr, ref = g[f"frames_{name}"], g[f"{ck}_{name}"]
            got = O.frames_to_clips(torch.from_numpy(fr), mean, std).numpy()
            assert got.shape == ref.shape
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (ck, name)
            via_lut = np.stack([lut[c][fr[..., c]] for c in range(3)], axis=1)            # (B,3,T,H,W)
            assert np.array_equal(via_lut.view(np.uint32), ref.view(np.uint32)), (ck, name, "lut")
    with pytest.raises(_lib.TuberError):
        _lib.input_lut((0.5, 0.5, 0.5), (0.2, 0.0, 0.2))                                   # zero std is refused
