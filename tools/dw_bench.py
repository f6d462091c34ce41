"""Times the depthwise 3x3x3 operator (tuber_op_dwconv, C-ABI) alone at the shapes of the four stages of TubeR_CSN152_AVA21 at 8 clips
32x256x256, CUDA events over back-to-back launches on rotating buffers (so that layer1 / layer2 inputs come from HBM as in the forward),
and checks the result of every shape against torch's conv3d in fp64.  python tools/dw_bench.py [reps]"""
import ctypes as C
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import tuber_b200  # noqa: F401,E402
from tuber_b200 import _lib  # noqa: E402

SHAPES = [("layer1", 8, 32, 64, 64, 64, 1, 1), ("layer2", 8, 16, 32, 32, 128, 1, 1), ("layer3", 8, 8, 16, 16, 256, 1, 1),
          ("layer4", 8, 4, 16, 16, 512, 1, 1), ("layer2.0", 8, 32, 64, 64, 128, 2, 2), ("layer3.0", 8, 16, 32, 32, 256, 2, 2)]


def P(t):
    return C.c_void_p(t.data_ptr())


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    lib = _lib.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    torch.manual_seed(0)
    for name, b, t, h, w, c, stt, sts in SHAPES:
        nbuf = max(2, min(6, int(600e6 // (b * t * h * w * c * 4)) + 1))
        xs = [torch.randn(b, t, h, w, c, device="cuda") for _ in range(nbuf)]
        wt = torch.randn(c, 1, 3, 3, 3, device="cuda") * 0.3
        scale, shift = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
        wpk = wt.reshape(c, 27).t().contiguous()
        to, ho, wo = (t - 1) // stt + 1, (h - 1) // sts + 1, (w - 1) // sts + 1
        outs = [torch.empty((b * to * ho * wo, c), device="cuda") for _ in range(nbuf)]

        def run(i):
            _lib.check(lib.tuber_op_dwconv(P(xs[i % nbuf]), P(wpk), P(scale), P(shift), P(outs[i % nbuf]), b, t, h, w, c, stt, sts, st))
        for i in range(5):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000.0 / reps
        # check one clip of buffer 0 against torch (fp64)
        run(0)
        got = torch.empty((b * to * ho * wo, c), device="cuda")
        _lib.check(lib.tuber_op_from_split(P(outs[0]), P(got), b * to * ho * wo, c, st))
        x0 = xs[0][:1].permute(0, 4, 1, 2, 3).double()
        ref = F.conv3d(x0, wt.double(), stride=(stt, sts, sts), padding=1, groups=c)
        ref = torch.relu(ref * scale.double()[None, :, None, None, None] + shift.double()[None, :, None, None, None])
        g0 = got.view(b, to, ho, wo, c)[:1].permute(0, 4, 1, 2, 3).double()
        err = ((g0 - ref).abs().max() / ref.abs().max()).item()
        nbytes = (b * t * h * w + b * to * ho * wo) * c * 4
        print(f"{name:9s} B={b} T={t} {h}x{w} C={c} stride {stt},{sts}: {us:7.1f} us  {nbytes / us / 1e3:7.0f} GB/s  rel err {err:.1e}", flush=True)


if __name__ == "__main__":
    main()
