"""tuber_b200 -- B200-native forward path of TubeR (CSN backbone + DETR encoder/decoder + heads).

Host side: Python/PyTorch for tensor plumbing only.  All arithmetic runs in the hand-written
sm_100a kernels of ``csrc/`` behind the C-ABI declared in ``include/tuber_b200.h``
(``libtuber_b200.so``, loaded with ctypes by ``_lib``); there is no CPU or PyTorch fallback.

    from tuber_b200 import load_cfg, build_model
    cfg = load_cfg("TubeR_CSN50_AVA21.yaml")
    model, criterion, postprocessors = build_model(cfg)       # reference signature (tuber_ava.py:160)
    out = model.cuda()(NestedTensor(clips, mask))              # reference signature (tuber_ava.py:97)
"""
from .config import CfgNode, get_cfg_defaults, load_cfg  # noqa: F401
from .models.tuber_ava import DETR, PostProcess, PostProcessAVA, build_model, format_detection_lines  # noqa: F401
from .utils.misc import NestedTensor, nested_tensor_from_tensor_list  # noqa: F401
from .utils.context_bank import ContextBank  # noqa: F401
from .distributed import gather_detections, pack_detections, shard_range, unpack_detections  # noqa: F401
from .launch import spawn_workers  # noqa: F401
from .frames import FrameDecoder, clip_size  # noqa: F401

__all__ = ["CfgNode", "get_cfg_defaults", "load_cfg", "DETR", "build_model", "PostProcess", "PostProcessAVA",
           "NestedTensor", "nested_tensor_from_tensor_list", "shard_range", "pack_detections",
           "unpack_detections", "gather_detections", "format_detection_lines", "ContextBank", "spawn_workers", "FrameDecoder", "clip_size"]
