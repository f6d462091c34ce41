"""tuber_b200 -- B200-native forward path of TubeR (CSN backbone + DETR encoder/decoder + heads).

Host side: Python/PyTorch for tensor plumbing only.  All arithmetic runs in the hand-written
sm_100a kernels of ``csrc/`` behind the C-ABI declared in ``include/tuber_b200.h``.
"""
from .config import CfgNode, get_cfg_defaults, load_cfg  # noqa: F401

__all__ = ["CfgNode", "get_cfg_defaults", "load_cfg"]
