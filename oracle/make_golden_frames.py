"""Generate tests/golden/frames.npz with Pillow itself (build container; Pillow is what the reference's loader calls).

TEST INFRASTRUCTURE.  Seeded synthetic images are encoded as baseline JPEGs by Pillow at several sizes, qualities, chroma
subsamplings (4:4:4 / 4:2:2 / 4:2:0) and restart intervals; stored per case: the JPEG bytes, the pixels `Image.open` decodes them
to, and the pixels `Image.resize((w, h))` -- the reference's call, datasets/ava_frame.py:148-149, default filter -- makes of those.

    python oracle/make_golden_frames.py
"""
from __future__ import annotations

import io
import os

import numpy as np
import PIL
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# name: (height, width, quality, subsampling (0: 4:4:4, 1: 4:2:2, 2: 4:2:0), restart interval in MCU rows or 0, resize to (h, w))
CASES = {
    "s420_q75": (48, 64, 75, 2, 0, (32, 43)),
    "s420_odd": (37, 53, 90, 2, 0, (64, 91)),
    "s420_rst": (72, 88, 60, 2, 1, (36, 44)),
    "s422_q85": (40, 50, 85, 1, 0, (25, 31)),
    "s444_q95": (33, 47, 95, 0, 0, (33, 20)),
    "s420_tiny": (9, 5, 80, 2, 0, (16, 16)),
    "s420_ava": (120, 160, 75, 2, 0, (96, 128)),
    "s420_up": (60, 80, 50, 2, 2, (128, 171)),
}


def synth(h: int, w: int, seed: int) -> np.ndarray:
    """smooth gradients + blobs + hard edges + noise: exercises DC prediction, long zero runs, saturation and the clamps"""
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w, 3))
    for c in range(3):
        img[..., c] = 128 + 90 * np.sin(xx / (5.0 + 3 * c) + g.uniform(0, 6)) * np.cos(yy / (7.0 - c) + g.uniform(0, 6))
    for _ in range(6):
        cy, cx, r = g.uniform(0, h), g.uniform(0, w), g.uniform(2, max(3, min(h, w) / 3))
        m = (yy - cy) ** 2 + (xx - cx) ** 2 < r * r
        img[m] = g.uniform(0, 255, size=3)
    img[:, w // 2:w // 2 + 2] = 255
    img[h // 3:h // 3 + 1] = 0
    img += g.normal(0, 12, size=img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def main() -> None:
    out = {"pillow_version": np.array(PIL.__version__)}
    for i, (name, (h, w, q, ss, rst, (rh, rw))) in enumerate(CASES.items()):
        src = Image.fromarray(synth(h, w, 100 + i))
        buf = io.BytesIO()
        kw = {"restart_marker_rows": rst} if rst else {}
        src.save(buf, format="JPEG", quality=q, subsampling=ss, optimize=(i % 2 == 1), **kw)
        data = buf.getvalue()
        im = Image.open(io.BytesIO(data))
        im.load()
        assert im.mode == "RGB"
        out[name + "/jpeg"] = np.frombuffer(data, dtype=np.uint8)
        out[name + "/pixels"] = np.asarray(im)
        out[name + "/resized"] = np.asarray(im.resize((rw, rh)))            # the reference's call: (width, height), default filter
    # resize-only cases on raw pixels (both directions, one axis unchanged, identity)
    for j, (h, w, rh, rw) in enumerate([(64, 64, 24, 100), (50, 70, 50, 35), (31, 31, 31, 31), (200, 150, 256, 192), (360, 480, 256, 341)]):
        px = synth(h, w, 200 + j)
        out[f"resize{j}/src"] = px
        out[f"resize{j}/dst"] = np.asarray(Image.fromarray(px).resize((rw, rh)))
    path = os.path.join(ROOT, "tests", "golden", "frames.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
