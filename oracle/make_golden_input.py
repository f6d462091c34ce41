"""Generate tests/golden/input_u8.npz from the UNMODIFIED reference input transforms (build container only).

TEST INFRASTRUCTURE.  Imports /root/reference/datasets/video_transforms.py as-is and runs the normalisation tail of the
reference's data pipeline (datasets/ava_frame.py:158-162: ``Compose([ToTensor(), Normalize(mean, std)])``, then :71-72
``torch.stack(imgs).permute(1, 0, 2, 3)``) on seeded synthetic PIL frames.  Stored: the uint8 frames and the fp32 clips, for the
reference's ImageNet constants and for a second, arbitrary mean / std; plus one 16x16 frame per channel holding all 256 byte
values (the complete value table of the transform).

    python oracle/make_golden_input.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("TUBER_REFERENCE", "/root/reference")
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]          # datasets/ava_frame.py:161
MEAN2, STD2 = [0.45, 0.5, 0.375], [0.225, 0.3, 0.25]


def reference_clip(T, frames_thwc: np.ndarray, mean, std) -> np.ndarray:
    tr = T.Compose([T.ToTensor(), T.Normalize(mean, std)])
    imgs, _ = tr([Image.fromarray(f) for f in frames_thwc], None)
    return torch.stack(imgs, dim=0).permute(1, 0, 2, 3).contiguous().numpy()      # (3,T,H,W)


def main():
    sys.path.insert(0, REF)
    import datasets.video_transforms as T                         # the reference's own transforms
    rng = np.random.default_rng(7)
    out = {}
    # clips: (2,4,12,20) -> T*H*W % 4 == 0 (vector path of the kernel); (1,3,7,9) -> 189 pixels (scalar path)
    for name, shape in (("a", (2, 4, 12, 20)), ("b", (1, 3, 7, 9))):
        fr = rng.integers(0, 256, size=shape + (3,), dtype=np.uint8)
        out[f"frames_{name}"] = fr
        out[f"clips_{name}"] = np.stack([reference_clip(T, f, MEAN, STD) for f in fr])
        out[f"clips2_{name}"] = np.stack([reference_clip(T, f, MEAN2, STD2) for f in fr])
    ramp = np.arange(256, dtype=np.uint8).reshape(1, 16, 16)
    table = np.stack([ramp[0]] * 3, axis=-1)[None]                # (1,16,16,3): every byte value in every channel
    out["frames_table"] = table[None]
    out["clips_table"] = reference_clip(T, table, MEAN, STD)[None]
    out["clips2_table"] = reference_clip(T, table, MEAN2, STD2)[None]
    out["mean"], out["std"], out["mean2"], out["std2"] = (np.asarray(v, dtype=np.float64) for v in (MEAN, STD, MEAN2, STD2))
    path = os.path.join(ROOT, "tests", "golden", "input_u8.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
