"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE.  Imports /root/reference as-is (read-only; nothing is copied), builds
the reference ``DETR`` through its own ``models.tuber_ava.build_model(cfg)``, loads the
oracle's seeded synthetic ``state_dict`` with ``strict=True`` (which also proves that
``oracle.tuber_oracle.param_spec`` names/shapes equal the reference's), runs
``model.eval()(NestedTensor(clips, mask))`` on CPU and stores the three output tensors for
every decoder layer plus a few intermediate statistics.  Inputs and weights are NOT stored:
they are regenerated from (config, seed) by ``make_state_dict`` / ``make_clips`` (numpy
PCG64, platform-stable), which keeps every fixture a few hundred KB at most.

Usage (from the repo root, in the container that has /root/reference):
    python oracle/make_golden.py            # regenerate all cases
    python oracle/make_golden.py A_csn50    # one case

The GPU box has no /root/reference; nothing under tests/ or bench.py imports this file.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("TUBER_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import tuber_oracle as O          # noqa: E402
from oracle.cases import CASES, CASES_ALL, build_case   # noqa: E402


def _reference_model(cfg):
    """reference build_model(cfg) -> DETR, with stdout noise swallowed."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        from models.tuber_ava import build_model as ref_build_model   # reference module
        model, _, _ = ref_build_model(cfg)
    return model.eval()


def run_case(name: str) -> None:
    case = CASES_ALL[name]
    cfg, sd, clips, mask = build_case(name)
    model = _reference_model(cfg)
    ref_sd = model.state_dict()
    missing = set(ref_sd) - set(sd)
    extra = set(sd) - set(ref_sd)
    assert not missing and not extra, f"state_dict mismatch: missing {sorted(missing)[:5]} extra {sorted(extra)[:5]}"
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(sd[k].shape), (k, v.shape, sd[k].shape)
    model.load_state_dict(sd, strict=True)

    from utils.misc import NestedTensor      # reference container (utils/misc.py:405)
    taps = {}
    hooks = [
        model.backbone.body.register_forward_hook(lambda m, i, o: taps.__setitem__("xt", o[0])),
        model.transformer.register_forward_hook(lambda m, i, o: taps.__setitem__("hs", o[0])),
        model.encoder.register_forward_hook(lambda m, i, o: taps.__setitem__("mem_c", o[0])),
    ]
    m = mask if mask is not None else torch.zeros((clips.shape[0],) + tuple(clips.shape[3:]), dtype=torch.bool)
    t0 = time.time()
    with torch.no_grad():
        out = model(NestedTensor(clips, m))
    dt = time.time() - t0
    for h in hooks:
        h.remove()

    nl = cfg.CONFIG.MODEL.DEC_LAYERS
    assert len(out["aux_outputs"]) == nl - 1
    stack = lambda key: torch.stack([a[key] for a in out["aux_outputs"]] + [out[key]]).numpy()
    xt = taps["xt"]
    # fixed pseudo-random probe of backbone features so stage-level drift is visible
    rng = np.random.default_rng(1234)
    probe_idx = rng.integers(0, xt.numel(), size=256)
    mem_c = taps["mem_c"]            # (T'H'W', L*B, d) replicated L times in the reference
    bsz = clips.shape[0]
    fixture = {
        "pred_logits": stack("pred_logits"),
        "pred_boxes": stack("pred_boxes"),
        "pred_logits_b": stack("pred_logits_b"),
        "xt_shape": np.array(xt.shape),
        "xt_probe_idx": probe_idx,
        "xt_probe": xt.flatten()[torch.from_numpy(probe_idx)].numpy(),
        "xt_absmean": np.array(float(xt.abs().mean())),
        "hs_last": taps["hs"][-1].numpy(),                           # (B,Q,d)
        "mem_c_first": mem_c[:, :bsz].permute(1, 0, 2)[:, :16].contiguous().numpy(),   # (B,16,d)
        "torch_version": np.array(torch.__version__),
        "ref_seconds": np.array(dt),
    }
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
    np.savez_compressed(path, **fixture)
    print(f"{name}: reference forward {dt:.2f}s -> {path} "
          f"({os.path.getsize(path) / 1024:.0f} KB) | mean|logits| {np.abs(fixture['pred_logits']).mean():.3f} "
          f"mean|boxes| {np.abs(fixture['pred_boxes']).mean():.3f} mean|logits_b| {np.abs(fixture['pred_logits_b']).mean():.3f}")
    del sys.modules["models.tuber_ava"]      # build_model prints / caches nothing else; allow re-import


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES_ALL)
    torch.set_num_threads(os.cpu_count() or 1)
    for n in names:
        run_case(n)
