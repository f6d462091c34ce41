"""Long-term context bank: host-side bookkeeping of the per-clip entries the model produces with ``MODEL.GENERATE_LFB`` /
``forward_raw(bank_out=...)`` and consumes as ``model(samples, lfb_features)`` (BASELINE.json configs[3]:
a 64-clip window).  Tensor plumbing only: the entries stay on the device they were produced on, a window is one ``torch.cat``.

The reference announces the feature (README.md:16-18,86) and keeps its switches (CONFIG.USE_LFB, MODEL.GENERATE_LFB) but never
released the code; the window convention here follows the paper: the clips within +-window/2 of the current one, the current clip
included.
"""
from __future__ import annotations

from typing import Dict, List

import torch
from torch import Tensor


class ContextBank:
    def __init__(self, window: int = 64):
        if window < 1:
            raise ValueError("window must be >= 1")
        self.window = int(window)
        self._videos: Dict[str, List[Tensor]] = {}

    def __len__(self) -> int:
        return sum(len(v) for v in self._videos.values())

    def num_clips(self, video: str) -> int:
        return len(self._videos.get(video, ()))

    def append(self, video: str, entries: Tensor) -> None:
        """entries (n, tokens, d): the bank entries of n consecutive clips of `video`, in temporal order."""
        if entries.dim() != 3:
            raise ValueError("entries must be (n, tokens, d)")
        self._videos.setdefault(video, []).extend(e for e in entries)

    def span(self, video: str, index: int):
        """[lo, hi): the clips of the window centred on clip `index` -- window // 2 before it, the rest from it on, shifted to stay
        inside the video (so every clip of a long enough video sees exactly `window` entries)."""
        n = self.num_clips(video)
        if not 0 <= index < n:
            raise IndexError(f"clip {index} of '{video}' is not in the bank ({n} clips)")
        lo = max(0, index - self.window // 2)
        hi = min(n, lo + self.window)
        lo = max(0, hi - self.window)
        return lo, hi

    def window_for(self, video: str, index: int) -> Tensor:
        """(1, clips * tokens, d): the `lfb_features` / `bank` argument for clip `index` of `video`."""
        lo, hi = self.span(video, index)
        return torch.cat(self._videos[video][lo:hi], dim=0).unsqueeze(0).contiguous()

    def windows_for(self, video: str, indices) -> Tensor:
        """(B, clips * tokens, d) for a batch of clips of one video; all windows have the same length by construction."""
        return torch.cat([self.window_for(video, i) for i in indices], dim=0).contiguous()
