"""Long-term context path at BASELINE.json configs[3] scale on cuda:0: B clips 32x256x256, a 64-clip window (16 384 bank tokens).
Prints clips/s with and without the window (CUDA-graph replay, CUDA events) and, with TUBER_KPROF_DUMP=1, the per-launch profile."""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tuber_b200  # noqa: E402
from oracle import tuber_oracle as O  # noqa: E402  (weight / clip generator only)
from tuber_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="TubeR_CSN152_AVA22.yaml")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--window", type=int, default=64)
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()
cfg = tuber_b200.load_cfg(a.config, ["CONFIG.USE_LFB", True])
sd = O.make_state_dict(cfg, 0, "random")
sd.update(O.make_ltc_state_dict(cfg))
model, _, _ = tuber_b200.build_model(cfg)
model.load_state_dict(sd)
model = model.cuda().eval()
B = a.batch
clips = O.make_clips(B, 32, 256, 256, seed=2).cuda()
entries = torch.empty(model.bank_entry_shape(B, 32, 256, 256), device="cuda")
model.forward_raw(clips, bank_out=entries)
tokens = entries.shape[1]
bank = (torch.randn(1, a.window * tokens, 256, device="cuda") * entries.std() + entries.mean()).contiguous()
bank[0, :tokens] = entries[0]
model.use_cuda_graph(True)
res = {"config": a.config, "batch": B, "bank_tokens": int(bank.shape[1])}
out = None
for name, kw in (("plain", {}), ("with_window", {"bank": bank}), ("with_window_and_entries", {"bank": bank, "bank_out": entries})):
    for _ in range(3):
        out = model.forward_raw(clips, None, out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        model.forward_raw(clips, None, out, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    res[name] = {"ms_per_step": ms, "clips_per_s": B * 1e3 / ms}
lib = _lib.load()
_lib.check(lib.tuber_set_kernel_profiling(model.plan(), 1))
model.forward_raw(clips, None, out, bank=bank)
torch.cuda.synchronize()
n = C.c_int32()
_lib.check(lib.tuber_get_kernel_profile(model.plan(), None, 0, C.byref(n)))
stats = (_lib.TuberKernelStat * n.value)()
_lib.check(lib.tuber_get_kernel_profile(model.plan(), stats, n.value, C.byref(n)))
res["kernels_with_window"] = {s.name.decode(): {"launches": s.launches, "ms": round(s.ms, 4)} for s in stats}
print(json.dumps(res))
