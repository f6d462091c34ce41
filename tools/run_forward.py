"""Run N eager forwards of a config on cuda:0 (the target of the ncu command lines in profiles/README.md)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tuber_b200  # noqa: E402
from oracle import tuber_oracle as O  # noqa: E402  (weight / clip generator only)

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="TubeR_CSN50_AVA21.yaml")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--clip", type=int, nargs=3, default=[32, 256, 256])
ap.add_argument("--n", type=int, default=1)
a = ap.parse_args()
cfg = tuber_b200.load_cfg(a.config)
model, _, _ = tuber_b200.build_model(cfg)
model.load_state_dict(O.make_state_dict(cfg, 0, "random"))
model = model.cuda().eval()
clips = O.make_clips(a.batch, *a.clip, seed=2).cuda()
for _ in range(a.n):
    model.forward_raw(clips)
torch.cuda.synchronize()
print("ok")
