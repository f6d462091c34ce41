"""The reference's OWN evaluation loops driving this repo's model, unchanged (BASELINE north_star: "drops into eval_tuber_ava.py";
second test: eval_tuber_jhmdb.py's loop, validate_tuber_ucf_detection, on the Tuber_CSN152_JHMDB configuration).

From the unmodified copy of the reference under baseline/_ref/ (tools/install_reference.py; git-ignored, travels with the snapshot):
``utils.model_utils.deploy_model`` (.cuda(gpu) + DistributedDataParallel(find_unused_parameters=True) + torch.load of
PRETRAIN_TRANSFORMER_DIR, model_utils.py:39-63), ``load_model`` (the TubeR .pth, intersected by name, :66-95) and
``utils.video_action_recognition.validate_tuber_detection`` (:222-453) are called as eval_tuber_ava.py calls them (:28-44), once with ``tuber_b200.build_model(cfg)`` and once with the reference's ``build_model(cfg)``
on the same weights and the same synthetic loader; the per-rank result files the loop writes ({rank}.txt, GT_{rank}.txt) and the
losses it prints must agree.  Only things outside the hot path are stubbed: the mAP evaluators (they open a hard-coded
/xxx/datasets csv, evaluates/evaluate_ava.py:36) and the loop's final 30 s sleep.
"""
import contextlib
import io
import os
import re
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


class _Writer:
    def add_scalar(self, *a, **kw):
        pass


class _NoEvaluator:
    def __init__(self, *a, **kw):
        pass

    def load_GT_from_path(self, paths):
        self.gt = list(paths)

    def load_detection_from_path(self, paths):
        self.det = list(paths)

    def evaluate(self):
        return [0.0], {}


def _loader(NestedTensor, batches, B, T, H, W, C):
    """What the reference's collate hands the loop (utils/misc.py collate_fn; datasets/ava_frame.py:76-127): (NestedTensor, targets)."""
    from oracle import tuber_oracle as O
    g = torch.Generator().manual_seed(5)
    key_pos = T // 2
    data, idx = [], 0
    for i in range(batches):
        clips = O.make_clips(B, T, H, W, seed=40 + i)
        mask = torch.zeros((B, H, W), dtype=torch.bool)
        targets = []
        for b in range(B):
            n = 1 + (i + b) % 3
            box = torch.cat((torch.rand(n, 2, generator=g) * 0.5 + 0.25, torch.rand(n, 2, generator=g) * 0.3 + 0.1), 1)
            labels = (torch.rand(n, C, generator=g) < 0.05).float()
            labels[:, (i * B + b) % C] = 1.0
            raw = torch.cat((torch.full((n, 1), float(idx)), torch.full((n, 1), float(key_pos)), box * torch.tensor([W, H, W, H])), 1)
            targets.append({"image_id": ["vid%02d_%04d" % (i, 902 + b), key_pos], "boxes": torch.cat((torch.full((n, 1), float(key_pos)), box), 1),
                            "raw_boxes": raw, "labels": labels, "size": torch.as_tensor([H, W]), "orig_size": torch.as_tensor([H, W])})
            idx += 1
        data.append((NestedTensor(clips, mask), targets))
    return data


def _fresh(data):
    # the loop deletes t["image_id"] in place (video_action_recognition.py:288-289): hand every run its own dictionaries
    return [(s, [dict(t) for t in ts]) for s, ts in data]


def _parse(path):
    ids, rows = [], []
    with open(path) as f:
        for line in f:
            ids.append(line.split(" [")[0])
            rows.append([float(x) for x in line.split(" [")[1].split("]")[0].split(",")])
    return ids, np.asarray(rows)


def _parse_scalar(path):
    """binary_{rank}.txt of the JHMDB loop: "<id> <list or number>" per line"""
    ids, rows = [], []
    with open(path) as f:
        for line in f:
            key, val = line.rstrip("\n").split(" ", 1)
            ids.append(key)
            rows.append(np.atleast_1d(np.asarray(eval(val, {"__builtins__": {}}), dtype=np.float64)))
    return ids, np.stack(rows) if rows else np.zeros((0, 1))


@pytest.mark.skipif(not os.path.isdir(REF), reason="baseline/_ref missing: run tools/install_reference.py where /root/reference is mounted")
def test_reference_eval_loop_runs_on_the_b200_model_unchanged(tmp_path, monkeypatch):
    import torch.distributed as dist
    import tuber_b200
    from oracle import tuber_oracle as O

    monkeypatch.syspath_prepend(REF)
    for name in [m for m in sys.modules if m.split(".")[0] in ("utils", "models", "evaluates", "datasets", "pipelines")]:
        monkeypatch.delitem(sys.modules, name)
    with contextlib.redirect_stdout(io.StringIO()):
        import utils.video_action_recognition as loop                  # the reference's modules, unmodified
        from models.tuber_ava import build_model as ref_build_model
        from utils.misc import NestedTensor
        from utils.model_utils import deploy_model, load_model
    monkeypatch.setattr(loop, "STDetectionEvaluater", _NoEvaluator)
    monkeypatch.setattr(loop, "STDetectionEvaluaterSinglePerson", _NoEvaluator)
    monkeypatch.setattr(loop.time, "sleep", lambda s: None)

    T, H, W, B = 32, 128, 160, 2
    cfg = tuber_b200.load_cfg("TubeR_CSN50_AVA21.yaml", ["CONFIG.EVAL_ONLY", True, "CONFIG.LOG.BASE_PATH", str(tmp_path), "DDP_CONFIG.GPU", 0,
                                                        "DDP_CONFIG.GPU_WORLD_RANK", 0, "DDP_CONFIG.GPU_WORLD_SIZE", 1,
                                                        "CONFIG.MODEL.PRETRAIN_TRANSFORMER_DIR", str(tmp_path / "detr.pth"),
                                                        "CONFIG.MODEL.PRETRAINED_PATH", str(tmp_path / "tuber.pth"), "CONFIG.MODEL.LOAD", True])
    sd = O.make_state_dict(cfg, seed=31, bn="random")
    sd["class_embed_b.bias"] = torch.tensor([0.0, 4.0, 0.0])             # opens PostProcessAVA's 0.8 actor gate for most queries
    # deploy_model unconditionally loads DETR-COCO weights into the wrapped model (model_utils.py:10-36,60): a synthetic detr.pth with
    # the wrapper's key prefix; it carries 100 COCO queries, of which the loader keeps QUERY_NUM
    detr = {"module." + k: v.clone() for k, v in sd.items() if k.split(".")[0] in ("transformer", "bbox_embed")}
    detr["module.query_embed.weight"] = torch.cat((sd["query_embed.weight"], torch.zeros(85, 256)))
    torch.save({"model": detr}, cfg.CONFIG.MODEL.PRETRAIN_TRANSFORMER_DIR)
    # the released TubeR checkpoint layout: {"model": {"module.<name>": tensor}, "epoch": n}; both arms start from their own random
    # initialisation and receive every weight through the wrapper, as in eval_tuber_ava.py:28-39
    torch.save({"model": {"module." + k: v for k, v in sd.items()}, "epoch": 0}, cfg.CONFIG.MODEL.PRETRAINED_PATH)
    data = _loader(NestedTensor, 4, B, T, H, W, 80)

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                               # the reference arm must be fp32 to be a reference
    results = {}
    try:
        for arm, builder in (("b200", tuber_b200.build_model), ("reference", ref_build_model)):
            cfg.CONFIG.LOG.RES_DIR = arm
            with contextlib.redirect_stdout(io.StringIO()):
                model, criterion, post = builder(cfg)
            log = io.StringIO()
            with contextlib.redirect_stdout(log):
                model = deploy_model(model, cfg, is_tuber=True)           # eval_tuber_ava.py:29
                assert isinstance(model, torch.nn.parallel.DistributedDataParallel)
                model, _ = load_model(model, cfg, load_fc=cfg.CONFIG.MODEL.LOAD_FC)     # eval_tuber_ava.py:39
                loop.validate_tuber_detection(cfg, model, criterion, post, _fresh(data), 0, _Writer())   # eval_tuber_ava.py:44
            assert "not found layers: dict_keys([])" in log.getvalue()
            ids, rows = _parse(tmp_path / arm / "0.txt")
            gt = open(tmp_path / arm / "GT_0.txt").read()
            last = [ln for ln in log.getvalue().splitlines() if ln.startswith("class_error:")][-1]
            results[arm] = (ids, rows, gt, [float(x) for x in re.findall(r": ([-0-9.eE+]+)", last)])
            del model
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
        if created:
            dist.destroy_process_group()

    (ids, rows, gt, losses), (rids, rrows, rgt, rlosses) = results["b200"], results["reference"]
    assert ids == rids and len(ids) == 4 * B * 15 and gt == rgt
    assert rows.shape == rrows.shape == (4 * B * 15, 4 + 80 + 1)
    p, rp = rows[:, -1], rrows[:, -1]
    assert np.abs(p - rp).max() < 1e-4
    assert (rp > 0.8).mean() > 0.3                                         # the gate is open for a good part of the queries
    assert np.abs(rows[:, :4] - rrows[:, :4]).max() < 1e-3 * max(H, W)     # boxes, pixels
    clear = np.abs(rp - 0.8) > 1e-3                                        # rows whose gate decision cannot flip inside the tolerance
    assert clear.mean() > 0.9
    assert np.abs(rows[clear, 4:-1] - rrows[clear, 4:-1]).max() < 1e-4     # class scores in [0, 1]
    assert len(losses) == len(rlosses) == 6
    for a, b in zip(losses, rlosses):                                      # class_error, loss, loss_bbox, loss_giou, loss_ce, loss_ce_b
        assert abs(a - b) <= 2e-3 * max(1.0, abs(b)), (losses, rlosses)


def _loader_jhmdb(NestedTensor, batches, B, T, H, W, C, temp_len):
    """(NestedTensor, targets) as datasets/jhmdb_frame.py hands them to validate_tuber_ucf_detection: integer labels, `vis`, `key_pos`
    (index of the key frame among the TEMP_LEN query groups), boxes as (key frame, cx, cy, w, h)."""
    from oracle import tuber_oracle as O
    g = torch.Generator().manual_seed(7)
    data, idx = [], 0
    for i in range(batches):
        clips = O.make_clips(B, T, H, W, seed=60 + i)
        mask = torch.zeros((B, H, W), dtype=torch.bool)
        targets = []
        for b in range(B):
            n = 1 + (i + b) % 2
            key_pos = int(torch.randint(0, temp_len, (1,), generator=g))
            box = torch.cat((torch.rand(n, 2, generator=g) * 0.5 + 0.25, torch.rand(n, 2, generator=g) * 0.3 + 0.1), 1)
            raw = torch.cat((torch.full((n, 1), float(idx)), torch.full((n, 1), float(key_pos)), box * torch.tensor([W, H, W, H])), 1)
            targets.append({"image_id": ["jhmdb%02d_%03d" % (i, b), key_pos], "boxes": torch.cat((torch.full((n, 1), float(key_pos)), box), 1),
                            "raw_boxes": raw, "labels": torch.randint(0, C, (n,), generator=g), "vis": torch.tensor([1]),
                            "key_pos": torch.tensor(key_pos), "size": torch.as_tensor([H, W]), "orig_size": torch.as_tensor([H, W])})
            idx += 1
        data.append((NestedTensor(clips, mask), targets))
    return data


@pytest.mark.skipif(not os.path.isdir(REF), reason="baseline/_ref missing: run tools/install_reference.py where /root/reference is mounted")
def test_reference_jhmdb_eval_loop_runs_on_the_b200_model_unchanged(tmp_path, monkeypatch):
    """eval_tuber_jhmdb.py:28-44 -- deploy_model, load_model, validate_tuber_ucf_detection (video_action_recognition.py:457-680) -- on
    Tuber_CSN152_JHMDB.yaml (320 tubelet queries, softmax classes with a no-object class, 2-way clip head on the pooled backbone
    feature, SetCriterion + matcher_ucf, PostProcess): result, binary and ground-truth files and the printed losses must agree
    between this package's model and the reference's."""
    import torch.distributed as dist
    import tuber_b200
    from oracle import tuber_oracle as O

    monkeypatch.syspath_prepend(REF)
    for name in [m for m in sys.modules if m.split(".")[0] in ("utils", "models", "evaluates", "datasets", "pipelines")]:
        monkeypatch.delitem(sys.modules, name)
    with contextlib.redirect_stdout(io.StringIO()):
        import utils.video_action_recognition as loop                  # the reference's modules, unmodified
        from models.tuber_ava import build_model as ref_build_model    # (eval_tuber_jhmdb.py:9 imports this module too)
        from utils.misc import NestedTensor
        from utils.model_utils import deploy_model, load_model
    monkeypatch.setattr(loop, "STDetectionEvaluaterUCF", _NoEvaluator)
    monkeypatch.setattr(loop.time, "sleep", lambda s: None)

    T, H, W, B = 16, 128, 160, 2
    cfg = tuber_b200.load_cfg("Tuber_CSN152_JHMDB.yaml", ["CONFIG.EVAL_ONLY", True, "CONFIG.LOG.BASE_PATH", str(tmp_path), "DDP_CONFIG.GPU", 0,
                                                         "DDP_CONFIG.GPU_WORLD_RANK", 0, "DDP_CONFIG.GPU_WORLD_SIZE", 1,
                                                         "CONFIG.MODEL.PRETRAIN_TRANSFORMER_DIR", str(tmp_path / "detr.pth"),
                                                         "CONFIG.MODEL.PRETRAINED_PATH", str(tmp_path / "tuber.pth"), "CONFIG.MODEL.LOAD", True])
    nq, temp_len, C = cfg.CONFIG.MODEL.QUERY_NUM, cfg.CONFIG.MODEL.TEMP_LEN, cfg.CONFIG.DATA.NUM_CLASSES
    sd = O.make_state_dict(cfg, seed=33, bn="random")
    # DETR-COCO weights for the transformer and the box head only: with SINGLE_FRAME the loader would cut query_embed to QUERY_NUM rows
    # (model_utils.py:19-24), which cannot fit the QUERY_NUM * TEMP_LEN rows of this configuration
    detr = {"module." + k: v.clone() for k, v in sd.items() if k.split(".")[0] in ("transformer", "bbox_embed")}
    torch.save({"model": detr}, cfg.CONFIG.MODEL.PRETRAIN_TRANSFORMER_DIR)
    torch.save({"model": {"module." + k: v for k, v in sd.items()}, "epoch": 0}, cfg.CONFIG.MODEL.PRETRAINED_PATH)
    data = _loader_jhmdb(NestedTensor, 3, B, T, H, W, C, temp_len)

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29534")
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    results = {}
    try:
        for arm, builder in (("b200", tuber_b200.build_model), ("reference", ref_build_model)):
            cfg.CONFIG.LOG.RES_DIR = arm
            with contextlib.redirect_stdout(io.StringIO()):
                model, criterion, post = builder(cfg)
            log = io.StringIO()
            with contextlib.redirect_stdout(log):
                model = deploy_model(model, cfg, is_tuber=True)
                criterion = criterion.cuda()                               # eval_tuber_jhmdb.py:41
                model, _ = load_model(model, cfg, load_fc=cfg.CONFIG.MODEL.LOAD_FC)
                loop.validate_tuber_ucf_detection(cfg, model, criterion, post, _fresh(data), 0, _Writer())
            assert "not found layers: dict_keys([])" in log.getvalue()
            ids, rows = _parse(tmp_path / arm / "0.txt")
            bids, brows = _parse_scalar(tmp_path / arm / "binary_0.txt")
            gt = open(tmp_path / arm / "GT_0.txt").read()
            last = [ln for ln in log.getvalue().splitlines() if ln.startswith("class_error:")][-1]
            results[arm] = (ids, rows, bids, brows, gt, [float(x) for x in re.findall(r": ([-0-9.eE+]+)", last)])
            del model
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
        if created:
            dist.destroy_process_group()

    (ids, rows, bids, brows, gt, losses), (rids, rrows, rbids, rbrows, rgt, rlosses) = results["b200"], results["reference"]
    assert ids == rids and bids == rbids and len(ids) == 3 * B * nq and gt == rgt
    assert rows.shape == rrows.shape == (3 * B * nq, 4 + C + 1)
    assert np.abs(rows[:, :4] - rrows[:, :4]).max() < 1e-3 * max(H, W)     # boxes of the key frame's queries, pixels
    assert np.abs(rows[:, 4:] - rrows[:, 4:]).max() < 1e-4                 # class scores in [0, 1]
    assert brows.shape == rbrows.shape and np.abs(brows - rbrows).max() < 1e-4
    assert len(losses) == len(rlosses) == 5
    for a, b in zip(losses, rlosses):                                      # class_error, loss, loss_bbox, loss_giou, loss_ce
        assert abs(a - b) <= 2e-3 * max(1.0, abs(b)), (losses, rlosses)
