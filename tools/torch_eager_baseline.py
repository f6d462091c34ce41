"""The "same box" bar SURVEY.md section 8d asks for next to the CPU baseline: the reference algorithm as stock PyTorch eager ops
(conv3d / batch_norm / linear / softmax / layer_norm through cuDNN + cuBLAS) ON THE B200, fp32 with TF32 off and on.  It runs the
oracle's functional restatement on CUDA tensors -- a measurement tool, not part of the product path or of bench.py.

    python tools/torch_eager_baseline.py [--config TubeR_CSN50_AVA21.yaml] [--batch 8] [--steps 10]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tuber_b200  # noqa: E402  (config loader only)
from oracle import tuber_oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="TubeR_CSN50_AVA21.yaml")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
cfg = tuber_b200.load_cfg(a.config)
sd = {k: v.cuda() for k, v in O.make_state_dict(cfg, 0, "random").items()}
clips = O.make_clips(a.batch, 32, 256, 256, seed=2).cuda()
torch.set_default_device("cuda")                      # the oracle's own torch.zeros / arange land on the GPU
res = {"config": a.config, "batch": a.batch, "steps": a.steps, "torch": torch.__version__, "device": torch.cuda.get_device_name(0)}
for name, tf32 in (("fp32", False), ("tf32", True)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    try:
        with torch.no_grad():
            for _ in range(3):
                O.forward(cfg, sd, clips, None)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                O.forward(cfg, sd, clips, None)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        res[name] = {"ms_per_step": ms, "clips_per_s": a.batch * 1e3 / ms}
    except Exception as exc:  # noqa: BLE001
        res[name] = {"error": repr(exc)[:300]}
print(json.dumps(res))
