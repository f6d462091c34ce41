"""Frame loading on the GPU: the reference loader's per-frame ``Image.open(path)`` + ``.resize((w, h))``
(datasets/ava_frame.py:146-150) behind the C-ABI (``tuber_frames_*``, csrc/frames.cu) -- bit-identical to Pillow.

    dec = FrameDecoder()
    frames = dec.decode([open(p, "rb").read() for p in paths], out_h, out_w)     # uint8 (n, out_h, out_w, 3) on the GPU
    out = model.forward_raw_u8(frames.view(B, T, out_h, out_w, 3))               # ToTensor + Normalize run on the device

Entropy (Huffman) decoding runs on host threads, everything after it -- dequantisation, inverse DCT, chroma upsampling,
colour conversion, Pillow's bicubic resample -- in CUDA kernels.  There is no CPU fallback: without the library or a GPU it raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib


class FrameDecodeError(RuntimeError):
    pass


def clip_size(orig_h: int, orig_w: int, resize_size: int) -> tuple:
    """(h, w) the reference resizes a frame of a video to: short side = resize_size, the other side scaled and truncated
    (datasets/ava_frame.py:86-91,127: ``int(nh), int(nw)``)."""
    if orig_w >= orig_h:
        nh, nw = resize_size, resize_size * (orig_w / orig_h)
    else:
        nw, nh = resize_size, resize_size * (orig_h / orig_w)
    return int(nh), int(nw)


class FrameDecoder:
    def __init__(self, device: Optional[torch.device] = None, host_threads: int = 0):
        if not torch.cuda.is_available():
            raise FrameDecodeError("tuber_b200.FrameDecoder needs a GPU: the frame path has no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._h = C.c_void_p()
        lib = _lib.load()
        with torch.cuda.device(self.device):
            st = lib.tuber_frames_create(C.byref(self._h), int(host_threads))
        if st != 0:
            raise FrameDecodeError(lib.tuber_frames_last_error().decode("utf-8", "replace"))

    def decode(self, jpegs: Sequence[bytes], out_h: int, out_w: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """JPEG byte strings -> RGB uint8 (n, out_h, out_w, 3) on the decoder's device, each frame decoded and resized as
        Pillow would (``Image.open(...).resize((out_w, out_h))``)."""
        n = len(jpegs)
        if n == 0:
            raise ValueError("no frames")
        if out is None:
            out = torch.empty((n, out_h, out_w, 3), dtype=torch.uint8, device=self.device)
        elif out.dtype != torch.uint8 or tuple(out.shape) != (n, out_h, out_w, 3) or not out.is_contiguous() or out.device != self.device:
            raise ValueError("out must be a contiguous uint8 (n, out_h, out_w, 3) tensor on the decoder's device")
        bufs = [bytes(j) for j in jpegs]                       # keep the byte strings alive over the call
        ptrs = (C.c_void_p * n)(*[C.cast(C.c_char_p(b), C.c_void_p) for b in bufs])
        sizes = (C.c_int64 * n)(*[len(b) for b in bufs])
        lib = _lib.load()
        with torch.cuda.device(self.device):
            st = lib.tuber_frames_decode(self._h, ptrs, sizes, n, int(out_h), int(out_w), C.c_void_p(out.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if st != 0:
            raise FrameDecodeError(lib.tuber_frames_last_error().decode("utf-8", "replace"))
        return out

    def load_clip(self, paths: Sequence[str], out_h: int, out_w: int) -> torch.Tensor:
        """The frames of one clip from JPEG files -> uint8 (T, out_h, out_w, 3) (ava_frame.py:146-150 for every sampled frame)."""
        data = []
        for p in paths:
            with open(p, "rb") as f:
                data.append(f.read())
        return self.decode(data, out_h, out_w)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            _lib.load().tuber_frames_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
