"""Clip-sharded multi-GPU inference: one process per GPU, replicated weights, one all-gather.

Clips are independent in the forward (eval-mode BatchNorm, no cross-sample op), so rank r of N
runs clips ``shard_range(B, r, N)`` of a global batch and the per-clip detections are exchanged
with ONE ``all_gather`` of a packed fp32 tensor -- replacing the reference's per-rank text files
and two barriers (utils/video_action_recognition.py:411-423,452).  ``torch.distributed`` (NCCL over
NVLink on the GPU box, gloo in the CPU tests) is plumbing only.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def shard_range(n_clips: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of the clips rank ``rank`` of ``world`` processes; earlier ranks take the remainder."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    base, rem = divmod(n_clips, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_detections(out: Dict[str, Tensor]) -> Tensor:
    """Last-layer detections of each clip as one row: [Q*4 boxes | Q*C logits | actor-ness logits]."""
    boxes, logits, lb = out["pred_boxes"], out["pred_logits"], out["pred_logits_b"]
    b = boxes.shape[0]
    return torch.cat((boxes.reshape(b, -1), logits.reshape(b, -1), lb.reshape(b, -1)), dim=1).contiguous()


def unpack_detections(packed: Tensor, num_queries: int, num_classes: int, ava: bool = True) -> Dict[str, Tensor]:
    b = packed.shape[0]
    nb, nl = num_queries * 4, num_queries * num_classes
    lb = packed[:, nb + nl:]
    return {"pred_boxes": packed[:, :nb].reshape(b, num_queries, 4),
            "pred_logits": packed[:, nb:nb + nl].reshape(b, num_queries, num_classes),
            "pred_logits_b": lb.reshape(b, num_queries, 3) if ava else lb.reshape(b, 2)}


def gather_detections(packed_local: Tensor, n_clips: int, group: Optional[dist.ProcessGroup] = None) -> Tensor:
    """All ranks' packed detections in global clip order, (n_clips, row).  One collective; uneven shards
    are padded to the largest shard for the exchange and trimmed afterwards."""
    if not dist.is_available() or not dist.is_initialized():
        return packed_local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return packed_local
    sizes = [shard_range(n_clips, r, world) for r in range(world)]
    lo, hi = sizes[rank]
    if packed_local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {packed_local.shape[0]} clips, expected {hi - lo}")
    width = packed_local.shape[1]
    cap = max(e - s for s, e in sizes)
    send = packed_local
    if hi - lo < cap:
        send = torch.zeros((cap, width), dtype=packed_local.dtype, device=packed_local.device)
        send[: hi - lo] = packed_local
    recv = torch.empty((world * cap, width), dtype=packed_local.dtype, device=packed_local.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if all(e - s == cap for s, e in sizes):
        return recv
    return torch.cat([recv[r * cap: r * cap + (e - s)] for r, (s, e) in enumerate(sizes)], dim=0)
