#!/usr/bin/env python
"""bench.py -- clips/s of the TubeR forward hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                         # the reference algorithm on the host CPU
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

A step = one forward over one batch of synthetic clips.  Workload at every N: the configuration BASELINE.json's metric is quoted
on -- TubeR_CSN152_AVA21.yaml shapes (CSN-152 backbone, 6+6 layers, 15 tubelet queries), random-init weights, 8 synthetic 32x256x256
clips per GPU per step (= BASELINE configs[2]'s 16-clip batch on two GPUs); clips shard over ranks with no data-path collective
inside the forward and one all-gather of packed detections per step ("weak" scaling: per-GPU batch fixed).
  value    whole-job clips/s, inputs resident in HBM, CUDA-graph replay of the launch sequence
  e2e      the same through the host entry points (tuber_forward_host_submit/_wait): pinned host clips -> H2D ->
           forward -> D2H of the detections, every step; two slots, so step i+1's copy overlaps step i's kernels
  roofline the kernel with the largest share of device time, from a per-launch CUDA-event profile of
           one extra forward (algorithmic bytes / flops per launch over the summed event time)
  cpu_baseline  the UNMODIFIED reference (baseline/_ref, placed by tools/install_reference.py: its own build_model(cfg) and
           model.eval()(NestedTensor) on the host cores, kind "reference"), else the CPU oracle port (kind "port"), on a bounded
           sample of the same workload
  also     the other BASELINE.json configs, same timing rules, fewer steps: configs[1] (CSN-50), configs[2] as STRONG scaling (16 clips
           in total: 16/N per GPU), configs[3] (long-term context bank, 64-clip window per clip), configs[4] (JHMDB, 16x256x256), and the
           headline model on the clip size the reference's evaluation transform really produces (32x256x341)
  same_box_baseline  the reference algorithm as stock PyTorch eager ops (cuDNN / cuBLAS) on the same B200, fp32 with TF32 off and on
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "clips/sec (32x256x256, CSN-152)"
UNIT = "clips/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "src": "fallback"}


def _ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu launch-list summary of the current round's build
    (profiles/r<round>_*_summary.json, newest round first, the round's `*final*` file before its earlier ones; written by tools/ncu_launches_summary.py from
    `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` over this very command)."""
    import glob
    import re

    def round_of(path):
        m = re.match(r"r(\d+)_", os.path.basename(path))
        return (int(m.group(1)) if m else 0, "final" in os.path.basename(path), os.path.basename(path))
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_summary.json")), key=round_of)
    for f in reversed(files):
        try:
            with open(f) as fh:
                k = json.load(fh)["kernels"].get(kernel)
            if k:
                return k["dram_bytes_per_launch"], os.path.relpath(f, ROOT)
        except (OSError, ValueError, KeyError):
            continue
    return None, None


def _bind_to_gpu_cpus(index: int):
    """Multi-GPU e2e is bound by host -> device copies (8 x 25 MB/clip streams): keep each rank's threads, and therefore its
    first-touched pinned buffers, on the CPUs NVML reports as local to its GPU.  Returns the CPU count or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def _cpu_forward(cfg, sd):
    """-> (fn(clips) running one CPU forward, kind).  kind "reference": the unmodified reference's own model (baseline/_ref);
    "port": the CPU oracle restatement (used when the reference copy did not travel)."""
    import contextlib
    import io
    from oracle import tuber_oracle as O
    if os.path.isdir(os.path.join(REF_DIR, "models")):
        try:
            if REF_DIR not in sys.path:
                sys.path.insert(0, REF_DIR)
            with contextlib.redirect_stdout(io.StringIO()):
                from models.tuber_ava import build_model as ref_build_model     # the reference's module, untouched
                from utils.misc import NestedTensor as RefNestedTensor
                model, _, _ = ref_build_model(cfg)
            model.load_state_dict(sd, strict=True)
            model.eval()

            def run(clips):
                mask = torch.zeros((clips.shape[0],) + tuple(clips.shape[3:]), dtype=torch.bool)
                with torch.no_grad():
                    return model(RefNestedTensor(clips, mask))
            return run, "reference"
        except Exception as exc:  # noqa: BLE001 -- fall back to the port, and say so on stderr
            print(f"bench.py: reference under baseline/_ref not usable ({exc!r}); timing the oracle port", file=sys.stderr)
    return (lambda clips: O.forward(cfg, sd, clips, None)), "port"


def _cpu_reference(cfg, sd, T, H, W, budget_s: float, max_clips: int):
    """Time the reference on the host CPU on a bounded sample: one clip per forward, all host threads."""
    from oracle import tuber_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run, kind = _cpu_forward(cfg, sd)
    clip = O.make_clips(1, T, H, W, seed=2)
    run(clip)                    # warm-up
    n, t0 = 0, time.perf_counter()
    while n < max_clips and (n == 0 or time.perf_counter() - t0 < budget_s):
        run(clip)
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, cores, kind, f"{n} forwards of 1 clip {T}x{H}x{W} after 1 warm-up, {dt:.1f} s"


def _same_box_baseline(cfg, sd, clips, steps=5):
    """SURVEY 8d's "same box" bar: the reference algorithm as stock PyTorch eager ops (conv3d / batch_norm / linear / softmax /
    layer_norm through cuDNN + cuBLAS) on this B200, fp32 with TF32 off and on, same clips.  The oracle's functional restatement on
    CUDA tensors -- a measurement, not a code path of the product."""
    from oracle import tuber_oracle as O
    res = {"what": "oracle restatement as PyTorch eager CUDA ops (cuDNN / cuBLAS), same weights and clips", "torch": torch.__version__,
           "batch": int(clips.shape[0])}
    sd_d = {k: v.cuda() for k, v in sd.items()}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    torch.set_default_device("cuda")
    try:
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cudnn.benchmark = True
            try:
                with torch.no_grad():
                    for _ in range(2):
                        O.forward(cfg, sd_d, clips, None)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(steps):
                        O.forward(cfg, sd_d, clips, None)
                    e1.record()
                    torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                res[name] = {"value": clips.shape[0] * 1e3 / ms, "unit": UNIT, "ms_per_step": ms, "steps": steps}
            except Exception as exc:  # noqa: BLE001
                res[name] = {"error": repr(exc)[:200]}
    finally:
        torch.set_default_device("cpu")
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = old
        del sd_d
        torch.cuda.empty_cache()
    return res


def _frame_loading_leg(B, T, H, W, model, seconds=2.0):
    """the reference loader's per-frame Image.open + resize (datasets/ava_frame.py:146-150) through
    tuber_b200.FrameDecoder (host Huffman decoding + CUDA kernels, bit-identical to Pillow) against Pillow itself on the host
    cores, on synthetic 360x480 JPEG frames resized to the clip size; then the whole chain JPEG bytes -> detections."""
    import io
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np
    import tuber_b200
    try:
        from PIL import Image
    except ImportError:
        return None
    from oracle.make_golden_frames import synth
    jpegs = []
    for i in range(T):
        buf = io.BytesIO()
        Image.fromarray(synth(360, 480, 500 + i)).save(buf, format="JPEG", quality=75)
        jpegs.append(buf.getvalue())
    jpegs = jpegs * B                                                        # one 8-clip step: B * T frames
    n = len(jpegs)
    dec = tuber_b200.FrameDecoder()
    frames = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        dec.decode(jpegs, H, W, out=frames)
    torch.cuda.synchronize()
    reps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        dec.decode(jpegs, H, W, out=frames)
        reps += 1
    torch.cuda.synchronize()
    gpu_fps = n * reps / (time.perf_counter() - t0)
    # the chain: JPEG bytes -> frames -> detections (decode of step i+1 on the host overlaps the forward of step i on the device)
    out = None
    for _ in range(2):
        out = model.forward_raw_u8(dec.decode(jpegs, H, W, out=frames).view(B, T, H, W, 3), None, out)
    torch.cuda.synchronize()
    reps2, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        out = model.forward_raw_u8(dec.decode(jpegs, H, W, out=frames).view(B, T, H, W, 3), None, out)
        reps2 += 1
    torch.cuda.synchronize()
    chain = B * reps2 / (time.perf_counter() - t0)
    dec.close()
    # Pillow, all host cores (decode and resize release the GIL)
    cores = len(os.sched_getaffinity(0))

    def one(b):
        return np.asarray(Image.open(io.BytesIO(b)).resize((W, H)))
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(one, jpegs[:cores]))
        done, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            list(ex.map(one, jpegs))
            done += n
        pil_fps = done / (time.perf_counter() - t0)
    return {"workload": f"{n} baseline 4:2:0 JPEG frames 360x480 (quality 75, {sum(map(len, jpegs)) // n} bytes on average) -> RGB uint8 "
                        f"{H}x{W} as the reference loader makes them (Image.open + resize, bicubic), one {B}-clip step per call",
            "value": gpu_fps, "unit": "frames/s", "clips_per_s_equivalent": gpu_fps / T,
            "jpeg_to_detections": {"value": chain, "unit": UNIT, "what": "FrameDecoder.decode + forward_raw_u8 per step, same stream"},
            "cpu_baseline": {"value": pil_fps, "unit": "frames/s", "cores": cores, "kind": "reference",
                             "sample": f"Pillow {__import__('PIL').__version__} Image.open + resize on {cores} threads, {done} frames"}}


def main():
    # stdout carries exactly one line, the JSON result: libraries that print to fd 1 (NCCL's version banner, ...) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="TubeR_CSN152_AVA21.yaml")
    ap.add_argument("--batch", type=int, default=8, help="clips per GPU per step")
    ap.add_argument("--clip", type=int, nargs=3, default=[32, 256, 256], metavar=("T", "H", "W"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads (other BASELINE configs, same-box baseline)")
    args = ap.parse_args()
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    T, H, W = args.clip

    import tuber_b200
    from oracle import tuber_oracle as O          # weights / clip generator + the CPU baseline (checker only)
    cfg = tuber_b200.load_cfg(args.config)
    sd = O.make_state_dict(cfg, seed=0, bn="random")
    workload = {"workload": f"{args.config} shapes (the configuration BASELINE.json's metric is quoted on; configs[2]'s backbone and heads), "
                            f"random-init weights, {args.batch} synthetic {T}x{H}x{W} clips per GPU per step",
                "global_batch": args.batch * world, "per_gpu_batch": args.batch,
                "parallelism": f"clip-sharded x{world}, one all-gather of packed detections per step",
                "l2": "inputs (201 MB of clips per step) and activations (GBs) exceed the 126 MB L2; no flush needed"}

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        per_step = max(1, min(args.batch, 2))
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        run, kind = _cpu_forward(cfg, sd)
        clips = O.make_clips(per_step, T, H, W, seed=2)
        for _ in range(max(1, min(args.warmup, 1))):
            run(clips)
        steps = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(steps):
            run(clips)
        dt = time.perf_counter() - t0
        v = per_step * steps / dt
        sample = (f"{steps} steps of {per_step} clips {T}x{H}x{W} (bounded sample of the {args.batch}-clip step), {dt:.1f} s; "
                  + ("the unmodified reference's build_model(cfg) / model.eval()(NestedTensor) from baseline/_ref" if kind == "reference"
                     else "oracle port (baseline/_ref absent); it evaluates the class branch once instead of DEC_LAYERS times"))
        emit({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                          "warmup": 1, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload,
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    # ---------------------------------------------------------------- this repo's arm (B200)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the tuber_b200 path has no CPU fallback")
    numa = _bind_to_gpu_cpus(local) if world > 1 else None   # pinned staging buffers land on the GPU's own NUMA node
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tuber_b200 import _lib
    lib = _lib.load()
    use_graph = not args.no_graph
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def build(config_name, overrides=(), extra_sd=None, state=None):
        c = tuber_b200.load_cfg(config_name, list(overrides))
        w = state if state is not None else O.make_state_dict(c, seed=0, bn="random")
        if extra_sd is not None:
            w = dict(w, **extra_sd(c))
        m, _, _ = tuber_b200.build_model(c)
        m.load_state_dict(w, strict=True)
        m = m.cuda().eval()
        m.use_cuda_graph(use_graph)
        return c, w, m

    def out_buffers(m, b):
        L, Q, NC = m.dec_layers, m.num_queries, m.num_class_out
        return {"pred_logits": torch.empty((b, L, Q, NC), device="cuda"), "pred_boxes": torch.empty((b, L, Q, 4), device="cuda"),
                "pred_logits_b": torch.empty((b, L, Q, 3) if m.dataset_mode == "ava" else (b, 2), device="cuda")}

    def timed(m, x, out, steps, n_global, **kw):
        """W warm-up steps, then `steps` steps between barrier + synchronize, CUDA events, max over ranks -> ms in total.
        Every step ends with the all-gather of the packed last-layer detections when there is more than one rank."""
        def step():
            m.forward_raw(x, None, out, **kw)
            if world > 1:
                last = {k: (v[:, -1] if v.dim() == 4 else v) for k, v in out.items()}
                tuber_b200.gather_detections(tuber_b200.pack_detections(last), n_global)
        for _ in range(warmup):
            step()
        sync_all()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    model = build(args.config, state=sd)[2]
    B = args.batch
    clips = O.make_clips(B, T, H, W, seed=2 + rank).cuda()
    out = out_buffers(model, B)
    for _ in range(warmup):
        model.forward_raw(clips, None, out)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(model, clips, out, args.steps, B * world)
    clocks = sampler.stop() if rank == 0 else None
    launches_per_step = lib.tuber_last_launches(model.plan())

    # ---- e2e: host buffers through the public host entry points (tuber_forward_host_submit / _wait): every step
    # copies that step's clips from pinned host memory (two alternating buffers) and reads its detections back;
    # the copy of step i+1 is in flight while step i computes
    h_clips = [torch.empty((B, 3, T, H, W), dtype=torch.float32).pin_memory() for _ in range(2)]
    for hc in h_clips:
        hc.copy_(clips)
    h_out = [model._host_out(B) for _ in range(2)]

    def pipelined(submit, bufs, n):
        submit(0, bufs[0], None, h_out[0])
        for i in range(1, n):
            submit(i & 1, bufs[i & 1], None, h_out[i & 1])
            model.forward_host_wait((i - 1) & 1)
        model.forward_host_wait((n - 1) & 1)

    def e2e_rate(submit, bufs, n):
        pipelined(submit, bufs, 3)
        sync_all()
        t0 = time.perf_counter()
        pipelined(submit, bufs, n)                  # returns after the last step's D2H copies have completed
        s = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(s, op=dist.ReduceOp.MAX)
        return B * world * n / float(s.item())

    e2e_steps = max(4, min(args.steps, 60))
    e2e_val = e2e_rate(model.forward_host_submit, h_clips, e2e_steps)
    h2d = h_clips[0].numel() * 4
    d2h = sum(v.numel() for v in h_out[0].values()) * 4
    # the same, one synchronous call per step (no overlap), for reference
    for _ in range(2):
        model.forward_host(h_clips[0], None, h_out[0])
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        model.forward_host(h_clips[0], None, h_out[0])
    e2e_sync_val = B * e2e_steps / (time.perf_counter() - t0)
    # host -> device ceiling of this box with every rank copying at once: the pinned clip buffer alone, no kernels
    sync_all()
    dst = torch.empty_like(clips)
    t0 = time.perf_counter()
    for i in range(8):
        dst.copy_(h_clips[i & 1], non_blocking=True)
    torch.cuda.synchronize()
    probe_s = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(probe_s, op=dist.ReduceOp.MAX)
    h2d_probe_gbs = 8 * h2d * world / float(probe_s.item()) / 1e9
    del dst
    # the same pipeline fed with decoded uint8 frames (B,T,H,W,3): 3 bytes per pixel cross the bus and the reference's ToTensor +
    # Normalize run on the device (tuber_forward_host_u8_submit)
    h_frames = [O.make_frames_u8(B, T, H, W, seed=20 + rank + 100 * i).pin_memory() for i in range(2)]
    e2e_u8_val = e2e_rate(model.forward_host_u8_submit, h_frames, e2e_steps)
    e2e_u8 = {"value": e2e_u8_val, "unit": UNIT, "h2d_bytes_per_step": h_frames[0].numel(), "d2h_bytes_per_step": d2h, "steps": e2e_steps,
              "api": "forward_host_u8_submit/_wait: uint8 (B,T,H,W,3) frames in pinned host memory, ToTensor + Normalize on the device "
                     "(+1 launch per step)"}
    del h_frames

    # ---- per-kernel profile of one forward (CUDA events around every launch) + stage timers, headline model
    kernels, roof, stage_ms, stages = [], None, {}, []
    if rank == 0:
        _lib.check(lib.tuber_set_kernel_profiling(model.plan(), 1))
        model.forward_raw(clips, None, out)
        torch.cuda.synchronize()
        n = C.c_int32()
        _lib.check(lib.tuber_get_kernel_profile(model.plan(), None, 0, C.byref(n)))
        stats = (_lib.TuberKernelStat * n.value)()
        _lib.check(lib.tuber_get_kernel_profile(model.plan(), stats, n.value, C.byref(n)))
        _lib.check(lib.tuber_set_kernel_profiling(model.plan(), 0))
        peaks = _peaks()
        tot_ms = sum(s.ms for s in stats) or 1.0
        for s in sorted(stats, key=lambda s: -s.ms):
            gbs = s.bytes / (s.ms * 1e-3) / 1e9 if s.ms > 0 else 0.0
            tfs = s.flops / (s.ms * 1e-3) / 1e12 if s.ms > 0 else 0.0
            kernels.append({"kernel": s.name.decode(), "launches": s.launches, "ms": round(s.ms, 4), "share": round(s.ms / tot_ms, 4),
                            "GB/s": round(gbs, 1), "TFLOP/s": round(tfs, 2), "hbm_frac": round(gbs / peaks["hbm_gbs"], 4),
                            "tensor_frac": round(tfs / peaks["bf16_tflops"], 4)})
        top = max(stats, key=lambda s: s.ms)
        t_hbm = top.bytes / (peaks["hbm_gbs"] * 1e9)
        # tensor-core work a kernel ISSUES per algorithmic FLOP, in bf16-rate units: the split kernels run three bf16 passes
        # (hi*hi + hi*mid + mid*hi); a kind::tf32 kernel runs one pass at half the bf16 rate
        passes = 2.0 if b"tf32" in top.name else 3.0
        on_tc = any(t in top.name for t in (b"tcgen05", b"gemm_bf16x3", b"gemm2_bf16x3", b"gemm_tf32", b"attention_tc", b"stem"))
        t_tc = passes * top.flops / (peaks["bf16_tflops"] * 1e12) if on_tc else 0.0
        if t_hbm >= t_tc:
            ach = top.bytes / (top.ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"]}
        else:
            ach = passes * top.flops / (top.ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
                    "note": f"issued bf16-equivalent FLOPs: {passes:g} tensor-core passes per contraction"}
        traffic, traffic_src = _ncu_traffic(top.name.decode())
        roof.update({"kernel": top.name.decode(), "launches_per_step": top.launches, "avg_launch_ms": top.ms / max(1, top.launches),
                     "algorithmic_bytes_per_launch": top.bytes / max(1, top.launches),
                     "algorithmic_flops_per_launch": top.flops / max(1, top.launches), "peak_source": peaks["src"], "traffic": traffic,
                     "traffic_source": traffic_src,
                     "how": "CUDA events around every launch of one extra forward (same stream, same buffers, after the timed region)"})
        stage_ms = model.stage_times_ms(clips)
        # each stage against the roofline that bounds it (BASELINE north_star): algorithmic bytes / flops of the stage's launches
        # (tuber_get_stage_work: operands read once, results written once, flops = 2 x MACs) over the stage's device time;
        # "tensor_frac_issued" counts the three bf16 passes the split GEMMs issue per contraction
        for name, w in model.stage_work().items():
            ms = stage_ms.get(name, 0.0)
            if ms <= 0:
                continue
            gbs, tfs = w["bytes"] / (ms * 1e-3) / 1e9, w["flops"] / (ms * 1e-3) / 1e12
            hf, tf = gbs / peaks["hbm_gbs"], tfs / peaks["bf16_tflops"]
            stages.append({"stage": name, "ms": round(ms, 4), "GB/s": round(gbs, 1), "TFLOP/s": round(tfs, 2), "hbm_frac": round(hf, 4),
                           "tensor_frac": round(tf, 4), "tensor_frac_issued": round(3 * tf, 4), "bound": "hbm" if hf >= 3 * tf else "tensor"})

    # ---- the other BASELINE.json configs, same timing rules (every rank takes part: the all-gather is inside the step)
    also = []
    steps2 = max(5, min(args.steps, 40))

    def also_entry(name, m, x, n_local, **kw):
        o = out_buffers(m, n_local)
        ms = timed(m, x, o, steps2, n_local * world, **kw)
        e = {"workload": name, "value": n_local * world * steps2 / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps2, "steps": steps2,
             "per_gpu_batch": n_local, "global_batch": n_local * world, "launches_per_step": lib.tuber_last_launches(m.plan())}
        also.append(e)
        return e

    if not args.no_also:
        # (1) the headline model on a clip of the size the reference's evaluation transform produces for 4:3 video (256 x 341)
        xw = O.make_clips(B, T, H, 341, seed=50 + rank).cuda()
        also_entry(f"{args.config} shapes, {B} synthetic {T}x{H}x341 clips per GPU per step (Resize_Custom output for 4:3 video, "
                   "datasets/video_transforms.py:213-228)", model, xw, B)
        del xw
        # (2) BASELINE configs[2] as written: 16 clips in total, sharded 16/N per GPU (strong scaling)
        if 16 % world == 0:
            nloc = 16 // world
            xs = clips[:nloc].contiguous() if nloc <= B else torch.cat((clips, O.make_clips(nloc - B, T, H, W, seed=70 + rank).cuda()))
            e = also_entry(f"BASELINE.json configs[2]: {args.config} shapes, 16 synthetic {T}x{H}x{W} clips in total sharded over {world} GPU(s) "
                           f"({nloc} per GPU)", model, xs, nloc)
            e["scaling"] = "strong"
            del xs
    frame_leg = None
    if world == 1 and not args.no_also:                                      # rank 0 at N = 1 only: host threads are not shared with other ranks
        frame_leg = _frame_loading_leg(B, T, H, W, model)
    del model, out, h_clips, h_out
    torch.cuda.empty_cache()
    if not args.no_also:
        # (3) BASELINE configs[1]: CSN-50 backbone, decode pool (round 1's headline workload)
        m2 = build("TubeR_CSN50_AVA21.yaml")[2]
        also_entry(f"BASELINE.json configs[1]: TubeR_CSN50_AVA21.yaml shapes, {B} synthetic {T}x{H}x{W} clips per GPU per step", m2, clips, B)
        del m2
        torch.cuda.empty_cache()
        # (4) BASELINE configs[4]: JHMDB head (320 queries, 22-way softmax, centre-slice pool), 16-frame clips
        m4 = build("Tuber_CSN152_JHMDB.yaml")[2]
        xj = O.make_clips(B, 16, H, W, seed=60 + rank).cuda()
        also_entry(f"BASELINE.json configs[4]: Tuber_CSN152_JHMDB.yaml shapes (320 queries), {B} synthetic 16x{H}x{W} clips per GPU per step", m4, xj, B)
        del m4, xj
        torch.cuda.empty_cache()
        # (5) BASELINE configs[3]: CSN-152 / decode pool + long-term context bank, one 64-clip window (16 384 tokens) PER CLIP
        m3 = build("TubeR_CSN152_AVA22.yaml", ["CONFIG.USE_LFB", True], extra_sd=O.make_ltc_state_dict)[2]
        entries = torch.empty(m3.bank_entry_shape(B, T, H, W), device="cuda")
        m3.forward_raw(clips, bank_out=entries)
        tokens = entries.shape[1]
        bank = (torch.randn(B, 64 * tokens, 256, device="cuda") * entries.std() + entries.mean()).contiguous()
        bank[:, :tokens] = entries
        e = also_entry(f"BASELINE.json configs[3]: TubeR_CSN152_AVA22.yaml shapes + long-term context bank (in-repo definition, parity unpinned: "
                       f"the reference never released the layer), 64-clip window = {64 * tokens} bank tokens per clip, {B} synthetic {T}x{H}x{W} "
                       "clips per GPU per step", m3, clips, B, bank=bank)
        e["bank_tokens_per_clip"] = int(64 * tokens)
        del m3, bank, entries
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = same_box = None
    if world == 1:                                                            # rank 0 at N = 1 only
        if not args.no_also:
            same_box = _same_box_baseline(cfg, sd, clips)
        if not args.no_cpu_baseline:
            v, cores, kind, sample = _cpu_reference(cfg, sd, T, H, W, budget_s=15.0, max_clips=32)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}

    value = B * world * args.steps / (ms_total * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload, "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "forward_host_submit/_wait (double-buffered: H2D of step i+1 overlaps step i)",
                    "synchronous_per_call_rank0": e2e_sync_val, "rank0_cpus_bound": numa,
                    "h2d_needed_gbs": e2e_val * h2d / B / 1e9, "h2d_probe_gbs": h2d_probe_gbs,
                    "h2d_probe": "all ranks copying their pinned fp32 clip buffer at once, no kernels: the host -> device ceiling of this box",
                    "u8": e2e_u8},
            "e2e_u8": e2e_u8,
            "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
            "cuda_graph": use_graph, "roofline": roof, "cpu_baseline": cpu, "same_box_baseline": same_box, "frame_loading": frame_leg, "also": also, "kernels": kernels,
            "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()}, "stages": stages,
            "arithmetic": "fp32 storage; GEMMs = 3-pass bf16 split (hi*hi+hi*lo+lo*hi) on tcgen05 with fp32 TMEM accumulation"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
