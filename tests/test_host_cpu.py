"""CPU: host logic -- the C-ABI library loads and exports what include/tuber_b200.h declares, the
model mirrors the reference's state_dict, the boundary containers and the clip sharding work.
No compute call is made here (there is no GPU, and the library has no CPU path)."""
import ctypes as C
import os
import re

import pytest
import torch

import tuber_b200
from oracle import tuber_oracle as O
from oracle.cases import CASES, load_case_cfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
YAMLS = ["TubeR_CSN50_AVA21.yaml", "TubeR_CSN152_AVA21.yaml", "TubeR_CSN152_AVA22.yaml", "Tuber_CSN152_JHMDB.yaml"]


def _header_functions():
    text = open(os.path.join(ROOT, "include", "tuber_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tuber_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from tuber_b200 import _lib
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tuber_b200.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert set(_lib.PROTOTYPES) == set(names)
    assert lib.tuber_abi_version() == _lib.TUBER_ABI_VERSION
    assert [lib.tuber_stage_name(i).decode() for i in range(_lib.NUM_STAGES)][0] == "stem"


def test_config_struct_matches_header():
    from tuber_b200 import _lib
    assert C.sizeof(_lib.TuberConfig) == 16 * 4
    assert C.sizeof(_lib.TuberShapeInfo) == 8 * 4 + 8


def test_plan_create_fails_loudly_without_gpu():
    from tuber_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = tuber_b200.load_cfg("TubeR_CSN50_AVA21.yaml")
    model, _, _ = tuber_b200.build_model(cfg)
    plan = C.c_void_p()
    st = _lib.load().tuber_plan_create(C.byref(model._tcfg), C.byref(plan))
    assert st == -4 and b"no CPU fallback" in _lib.load().tuber_last_error()
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 3, 8, 64, 64))


def test_bad_config_rejected():
    from tuber_b200 import _lib
    cfg = tuber_b200.load_cfg("TubeR_CSN50_AVA21.yaml")
    model, _, _ = tuber_b200.build_model(cfg)
    bad = _lib.TuberConfig.from_buffer_copy(model._tcfg)
    bad.abi_version = 99
    plan = C.c_void_p()
    assert _lib.load().tuber_plan_create(C.byref(bad), C.byref(plan)) == -1


@pytest.mark.parametrize("yaml", YAMLS)
def test_state_dict_mirrors_reference(yaml):
    """names/shapes == oracle.param_spec, which make_golden.py proved equal to the live reference's."""
    cfg = tuber_b200.load_cfg(yaml)
    model, criterion, post = tuber_b200.build_model(cfg)
    spec = {n: s for n, s, _ in O.param_spec(cfg)}
    sd = model.state_dict()
    assert set(sd) == set(spec)
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(spec[k]), k
    model.load_state_dict(O.make_state_dict(cfg, 0), strict=True)
    assert set(post) == {"bbox"}
    # the criterion is live (tests/test_criterion_cpu.py) and carries the weights the evaluation loop multiplies with
    assert {"loss_ce", "loss_bbox", "loss_giou", "loss_ce_b", "loss_ce_4"} <= set(criterion.weight_dict)


def test_case_configs_build():
    for name in CASES:
        cfg = load_case_cfg(name)
        tuber_b200.build_model(cfg)


def test_nested_tensor_padding_matches_oracle():
    clips = [torch.randn(3, 4, 12, 16), torch.randn(3, 4, 12, 10), torch.randn(3, 4, 8, 16)]
    nt = tuber_b200.nested_tensor_from_tensor_list(clips)
    ref, mask = O.pad_clips(clips)
    assert torch.equal(nt.tensors, ref) and torch.equal(nt.mask, mask)
    t, m = nt.decompose()
    assert t is nt.tensors and m is nt.mask


def test_postprocess_matches_oracle():
    torch.manual_seed(0)
    out = {"pred_logits": torch.randn(2, 5, 80), "pred_boxes": torch.rand(2, 5, 4), "pred_logits_b": torch.randn(2, 5, 3) * 4}
    sizes = torch.tensor([[240.0, 320.0], [256.0, 256.0]])
    scores, boxes, pb = tuber_b200.PostProcessAVA()(out, sizes)
    rs, rb, rp = O.postprocess_ava(out["pred_logits"], out["pred_boxes"], out["pred_logits_b"], sizes)
    assert torch.allclose(torch.from_numpy(scores), rs) and torch.allclose(torch.from_numpy(boxes), rb)
    assert torch.allclose(torch.from_numpy(pb), rp)


def test_shard_range_partitions():
    for n in (1, 7, 16, 17):
        for world in (1, 2, 4, 8):
            spans = [tuber_b200.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    out = {"pred_logits": torch.randn(3, 15, 80), "pred_boxes": torch.rand(3, 15, 4), "pred_logits_b": torch.randn(3, 15, 3)}
    packed = tuber_b200.pack_detections(out)
    assert packed.shape == (3, 15 * 87)
    back = tuber_b200.unpack_detections(packed, 15, 80, ava=True)
    for k in out:
        assert torch.equal(back[k], out[k])
    j = {"pred_logits": torch.randn(2, 320, 22), "pred_boxes": torch.rand(2, 320, 4), "pred_logits_b": torch.randn(2, 2)}
    back = tuber_b200.unpack_detections(tuber_b200.pack_detections(j), 320, 22, ava=False)
    for k in j:
        assert torch.equal(back[k], j[k])
