"""GPU: the long-term context path (tuber_forward_ltc; SURVEY 8f row 3, BASELINE.json configs[3]) against the oracle's definition of
the layer.  "Parity unpinned": the reference never released this feature, so the oracle here is this repository's own restatement
of the paper's description (oracle/tuber_oracle.py::forward(bank=...)), not a restatement of reference code."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-3
_SMALL = ["CONFIG.MODEL.ENC_LAYERS", 2, "CONFIG.MODEL.DEC_LAYERS", 2, "CONFIG.MODEL.QUERY_NUM", 4, "CONFIG.MODEL.TEMP_LEN", 8]


def _setup(yaml, over, seed=0):
    import tuber_b200
    from oracle import tuber_oracle as O
    cfg = tuber_b200.load_cfg(yaml, list(over) + ["CONFIG.USE_LFB", True])
    sd = O.make_state_dict(cfg, seed=seed, bn="random")
    sd.update(O.make_ltc_state_dict(cfg, seed=seed + 100))
    model, _, _ = tuber_b200.build_model(cfg)
    model.load_state_dict(sd, strict=True)
    return cfg, sd, model.cuda().eval()


def _rel(a, b):
    from oracle import tuber_oracle as O
    return O.rel_err(a.detach().float().cpu(), b.detach().float().cpu())


def _check(got, ref):
    for k in ("pred_logits", "pred_boxes", "pred_logits_b"):
        emax, el2 = _rel(got[k].permute(1, 0, 2, 3), ref[k])
        assert emax <= TOL and el2 <= TOL, (k, emax, el2)


@pytest.mark.parametrize("yaml,over,shape", [("TubeR_CSN50_AVA21.yaml", _SMALL, (8, 96, 128)),
                                             ("TubeR_CSN152_AVA22.yaml", [], (32, 64, 64))])
def test_context_layer_matches_oracle(yaml, over, shape):
    from oracle import tuber_oracle as O
    cfg, sd, model = _setup(yaml, over)
    T, H, W = shape
    # fill pass: three batches of two clips -> six bank entries
    entries, ref_entries, batches = [], [], []
    for i in range(3):
        clips = O.make_clips(2, T, H, W, seed=40 + i)
        batches.append(clips)
        e = torch.empty(model.bank_entry_shape(2, T, H, W), device="cuda")
        model.forward_raw(clips.cuda(), bank_out=e)
        taps = {}
        O.forward(cfg, sd, clips, None, taps)
        entries.append(e)
        ref_entries.append(taps["bank_new"])
        emax, el2 = _rel(e, taps["bank_new"])
        assert emax <= TOL and el2 <= TOL, ("bank entries", emax, el2)
    tokens = entries[0].shape[1]
    bank = torch.cat(entries).reshape(1, 6 * tokens, 256).contiguous()
    ref_bank = torch.cat(ref_entries).reshape(1, 6 * tokens, 256)
    clips = batches[1]
    taps = {}
    ref = O.forward(cfg, sd, clips, None, taps, bank=ref_bank)
    got = model.forward_raw(clips.cuda(), bank=bank)
    _check(got, ref)
    emax, el2 = _rel(model.debug_fetch("mem_ltc").view(2, -1, 256), taps["mem_ltc"])
    assert emax <= TOL and el2 <= TOL, ("mem_ltc", emax, el2)
    # the context layer changes the class logits and nothing else
    plain = model.forward_raw(clips.cuda())
    assert torch.equal(plain["pred_boxes"], got["pred_boxes"]) and torch.equal(plain["pred_logits_b"], got["pred_logits_b"])
    assert (plain["pred_logits"] - got["pred_logits"]).abs().max() > 1e-3
    # one window per clip (different windows) against the oracle; the same window twice == the shared one, bit for bit
    per = torch.cat([bank[:, : 4 * tokens], bank[:, 2 * tokens:]]).contiguous()
    ref_per = torch.cat([ref_bank[:, : 4 * tokens], ref_bank[:, 2 * tokens:]])
    _check(model.forward_raw(clips.cuda(), bank=per), O.forward(cfg, sd, clips, None, None, bank=ref_per))
    same = model.forward_raw(clips.cuda(), bank=bank.expand(2, -1, -1).contiguous())
    assert torch.equal(same["pred_logits"], got["pred_logits"])
    # CUDA graph replay with a bank, and entries + bank in one call
    model.use_cuda_graph(True)
    e2 = torch.empty_like(entries[1])
    for _ in range(3):
        again = model.forward_raw(clips.cuda(), bank=bank, bank_out=e2)
        torch.cuda.synchronize()
        assert torch.equal(again["pred_logits"], got["pred_logits"]) and torch.equal(e2, entries[1])


def test_reference_style_calls_and_errors():
    import tuber_b200
    from oracle import tuber_oracle as O
    from tuber_b200 import _lib
    cfg, sd, model = _setup("TubeR_CSN50_AVA21.yaml", _SMALL)
    clips = O.make_clips(2, 8, 64, 64, seed=7)
    # MODEL.GENERATE_LFB: model(samples) returns the bank entries (tuber_jhmdb.py:111-112)
    gcfg = tuber_b200.load_cfg("TubeR_CSN50_AVA21.yaml", _SMALL + ["CONFIG.USE_LFB", True, "CONFIG.MODEL.GENERATE_LFB", True])
    gen, _, _ = tuber_b200.build_model(gcfg)
    gen.load_state_dict(sd, strict=True)
    gen = gen.cuda().eval()
    entries = gen(tuber_b200.NestedTensor(clips.cuda(), torch.zeros(2, 64, 64, dtype=torch.bool, device="cuda")))
    assert entries.shape == (2, 16, 256)
    bank = tuber_b200.ContextBank(window=64)
    bank.append("video0", entries)
    window = bank.window_for("video0", 1)
    # CONFIG.USE_LFB: model(samples, lfb_features) (utils/video_action_recognition.py:133-137)
    out = model(clips.cuda(), window)
    ref = O.as_reference_dict(O.forward(cfg, sd, clips, None, None, bank=window.cpu()))
    for k in ("pred_logits", "pred_boxes", "pred_logits_b"):
        emax, el2 = _rel(out[k], ref[k])
        assert emax <= TOL and el2 <= TOL, (k, emax, el2)
    assert len(out["aux_outputs"]) == 1
    with pytest.raises(_lib.TuberError):                      # 3 windows for 2 clips
        model.forward_raw(clips.cuda(), bank=window.expand(3, -1, -1).contiguous())
    with pytest.raises(ValueError):
        model.forward_raw(clips.cuda(), bank=window.cpu())
    with pytest.raises(ValueError):
        model.forward_raw(clips.cuda(), bank_out=torch.empty(2, 15, 256, device="cuda"))
    # a plan without the layer refuses a bank
    pcfg = tuber_b200.load_cfg("TubeR_CSN50_AVA21.yaml", _SMALL)
    plain, _, _ = tuber_b200.build_model(pcfg)
    plain.load_state_dict(O.make_state_dict(pcfg, seed=0, bn="random"), strict=True)
    plain = plain.cuda().eval()
    assert _lib.load().tuber_has_ltc(plain.plan()) == 0 and _lib.load().tuber_has_ltc(model.plan()) == 1
    with pytest.raises(_lib.TuberError):
        plain.forward_raw(clips.cuda(), bank=window)
    e = torch.empty(2, 16, 256, device="cuda")
    plain.forward_raw(clips.cuda(), bank_out=e)               # producing entries needs no extra weights
    assert torch.equal(e, entries)


def test_full_window_64_clips_at_full_size():
    """BASELINE.json configs[3]: a 64-clip window (64 x 256 = 16 384 bank tokens) at 32x256x256, one clip against the oracle."""
    from oracle import tuber_oracle as O
    cfg, sd, model = _setup("TubeR_CSN50_AVA21.yaml", [])
    g = torch.Generator().manual_seed(3)
    clips = O.make_clips(1, 32, 256, 256, seed=8)
    e = torch.empty(model.bank_entry_shape(1, 32, 256, 256), device="cuda")
    model.forward_raw(clips.cuda(), bank_out=e)
    assert e.shape == (1, 256, 256)
    # 63 synthetic neighbours with the statistics of a real entry + the clip's own entry
    others = torch.randn(63, 256, 256, generator=g) * float(e.std()) + float(e.mean())
    bank = torch.cat([others[:32].cuda(), e, others[32:].cuda()]).reshape(1, 64 * 256, 256).contiguous()
    torch.set_num_threads(os.cpu_count() or 1)
    ref = O.forward(cfg, sd, clips, None, None, bank=bank.cpu())
    model.use_cuda_graph(True)
    for _ in range(2):
        got = model.forward_raw(clips.cuda(), bank=bank)
        torch.cuda.synchronize()
        _check(got, ref)
