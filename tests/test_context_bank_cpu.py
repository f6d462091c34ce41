"""CPU: host logic of the long-term context path -- window bookkeeping, the optional ltc_* parameters, and the
oracle's own definition of the layer (which has NO reference counterpart: the reference never released it)."""
import pytest
import torch

import tuber_b200
from oracle import tuber_oracle as O

_SMALL = ["CONFIG.MODEL.ENC_LAYERS", 2, "CONFIG.MODEL.DEC_LAYERS", 2, "CONFIG.MODEL.QUERY_NUM", 4, "CONFIG.MODEL.TEMP_LEN", 8]


def test_window_bookkeeping():
    bank = tuber_b200.ContextBank(window=4)
    ent = torch.arange(10, dtype=torch.float32).view(10, 1, 1).expand(10, 2, 3).contiguous()
    bank.append("v", ent[:6])
    bank.append("v", ent[6:])
    bank.append("w", ent[:3])
    assert len(bank) == 13 and bank.num_clips("v") == 10 and bank.num_clips("w") == 3
    assert [bank.span("v", i) for i in (0, 1, 2, 3, 5, 8, 9)] == [(0, 4), (0, 4), (0, 4), (1, 5), (3, 7), (6, 10), (6, 10)]
    assert bank.span("w", 1) == (0, 3)                       # a video shorter than the window: all of it
    w = bank.window_for("v", 5)
    assert w.shape == (1, 8, 3) and w[0, ::2, 0].tolist() == [3.0, 4.0, 5.0, 6.0]
    ws = bank.windows_for("v", [0, 9])
    assert ws.shape == (2, 8, 3) and ws[1, ::2, 0].tolist() == [6.0, 7.0, 8.0, 9.0]
    with pytest.raises(IndexError):
        bank.span("v", 10)
    with pytest.raises(ValueError):
        tuber_b200.ContextBank(0)


def test_ltc_parameters_are_optional_and_named():
    plain, _, _ = tuber_b200.build_model(tuber_b200.load_cfg("TubeR_CSN50_AVA21.yaml", _SMALL))
    cfg = tuber_b200.load_cfg("TubeR_CSN50_AVA21.yaml", _SMALL + ["CONFIG.USE_LFB", True])
    ltc, _, _ = tuber_b200.build_model(cfg)
    extra = sorted(set(ltc.state_dict()) - set(plain.state_dict()))
    assert extra == sorted(n for n, _, _ in O.ltc_param_spec(cfg))
    sd = O.make_state_dict(cfg, seed=0)
    sd.update(O.make_ltc_state_dict(cfg))
    ltc.load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError):
        plain.forward(torch.zeros(1, 3, 8, 64, 64), lfb_features=torch.zeros(1, 4, 256))   # built without USE_LFB


def test_oracle_context_layer_definition():
    cfg = tuber_b200.load_cfg("TubeR_CSN50_AVA21.yaml", _SMALL + ["CONFIG.USE_LFB", True])
    sd = O.make_state_dict(cfg, seed=0, bn="random")
    sd.update(O.make_ltc_state_dict(cfg))
    clips = O.make_clips(2, 8, 64, 64, seed=4)
    taps = {}
    base = O.forward(cfg, sd, clips, None, taps)
    ent = taps["bank_new"]
    assert ent.shape == (2, 16, 256)                         # H'W' = 4 x 4 tokens per clip
    bank = ent.reshape(1, 32, 256)
    taps2 = {}
    out = O.forward(cfg, sd, clips, None, taps2, bank=bank)
    # only the class branch sees the bank: boxes and actor-ness logits are unchanged, class logits move
    assert torch.equal(out["pred_boxes"], base["pred_boxes"]) and torch.equal(out["pred_logits_b"], base["pred_logits_b"])
    assert (out["pred_logits"] - base["pred_logits"]).abs().max() > 1e-3
    # a window shared by the batch == the same window given per clip
    per_clip = O.forward(cfg, sd, clips, None, None, bank=bank.expand(2, -1, -1).contiguous())
    assert torch.equal(per_clip["pred_logits"], out["pred_logits"])
    # the layer itself, restated with torch.nn.MultiheadAttention
    mha = torch.nn.MultiheadAttention(256, 8, batch_first=True).eval()
    mha.load_state_dict({"in_proj_weight": sd["ltc_attn.in_proj_weight"], "in_proj_bias": sd["ltc_attn.in_proj_bias"],
                         "out_proj.weight": sd["ltc_attn.out_proj.weight"], "out_proj.bias": sd["ltc_attn.out_proj.bias"]})
    mem_c = taps2["mem_c"]
    with torch.no_grad():
        want = torch.nn.functional.layer_norm(mem_c + mha(mem_c, bank.expand(2, -1, -1), bank.expand(2, -1, -1))[0], (256,),
                                              sd["ltc_norm.weight"], sd["ltc_norm.bias"], 1e-5)
    assert (want - taps2["mem_ltc"]).abs().max() < 1e-5
