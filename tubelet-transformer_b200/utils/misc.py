"""The boundary container of the forward path: ``NestedTensor`` and the clip-list collation.

Mirrors the slice of the reference's ``utils/misc.py`` that ``DETR.forward`` touches:
``NestedTensor`` (utils/misc.py:405-425) and ``nested_tensor_from_tensor_list`` (:367-402).
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import Tensor


class NestedTensor(object):
    """A batch of zero-padded clips ``tensors`` (B,3,T,H,W) and its ``mask`` (B,H,W), True on padding."""

    def __init__(self, tensors: Tensor, mask: Optional[Tensor]):
        self.tensors = tensors
        self.mask = mask

    def to(self, device) -> "NestedTensor":
        mask = self.mask.to(device) if self.mask is not None else None
        return NestedTensor(self.tensors.to(device), mask)

    def decompose(self):
        return self.tensors, self.mask

    def __repr__(self) -> str:
        return str(self.tensors)


def nested_tensor_from_tensor_list(tensor_list: List[Tensor]) -> NestedTensor:
    """Pad (3,T,H,W) clips to the batch maximum; the mask marks padded pixels of the H,W plane."""
    if tensor_list[0].ndim != 4:
        raise ValueError("expected a list of (C,T,H,W) clips")
    dims = [max(int(c.shape[i]) for c in tensor_list) for i in range(4)]
    first = tensor_list[0]
    batch = torch.zeros((len(tensor_list), *dims), dtype=first.dtype, device=first.device)
    mask = torch.ones((len(tensor_list), dims[2], dims[3]), dtype=torch.bool, device=first.device)
    for i, clip in enumerate(tensor_list):
        c, t, h, w = clip.shape
        batch[i, :c, :t, :h, :w].copy_(clip)
        mask[i, :h, :w] = False
    return NestedTensor(batch, mask)
