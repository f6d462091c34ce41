#!/usr/bin/env python
"""bench.py -- clips/s of the TubeR forward hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                         # the reference algorithm on the host CPU
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

A step = one forward over one batch of synthetic clips.  Workload at every N: BASELINE.json configs[1]
(TubeR_CSN50_AVA21.yaml shapes, random-init weights, 8 synthetic 32x256x256 clips per GPU); clips shard
over ranks with no data-path collective inside the forward and one all-gather of packed detections per
step ("weak" scaling: per-GPU batch fixed).
  value    whole-job clips/s, inputs resident in HBM, CUDA-graph replay of the launch sequence
  e2e      the same through the host entry points (tuber_forward_host_submit/_wait): pinned host clips -> H2D ->
           forward -> D2H of the detections, every step; two slots, so step i+1's copy overlaps step i's kernels
  roofline the kernel with the largest share of device time, from a per-launch CUDA-event profile of
           one extra forward (algorithmic bytes / flops per launch over the summed event time)
  cpu_baseline  the CPU oracle (a restatement of the reference on the same torch CPU ops) on a bounded
           sample of the same workload, on this box's host cores
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

METRIC = "clips/sec (32x256x256 synthetic clips, TubeR forward)"
UNIT = "clips/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "src": "fallback"}


def _ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the newest committed ncu launch-list summary (profiles/*_summary.json,
    written by tools/ncu_launches_summary.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_summary.json")), key=os.path.getmtime)
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


def _cpu_reference(cfg, sd, T, H, W, budget_s: float, max_clips: int):
    """Time the CPU oracle on a bounded sample: one clip per forward, all host threads."""
    from oracle import tuber_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    clip = O.make_clips(1, T, H, W, seed=2)
    O.forward(cfg, sd, clip, None)                    # warm-up
    n, t0 = 0, time.perf_counter()
    while n < max_clips and (n == 0 or time.perf_counter() - t0 < budget_s):
        O.forward(cfg, sd, clip, None)
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, cores, f"{n} forwards of 1 clip {T}x{H}x{W} after 1 warm-up, {dt:.1f} s"


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
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="TubeR_CSN50_AVA21.yaml")
    ap.add_argument("--batch", type=int, default=8, help="clips per GPU per step")
    ap.add_argument("--clip", type=int, nargs=3, default=[32, 256, 256], metavar=("T", "H", "W"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary CSN-152 measurement")
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
    workload = {"workload": f"{args.config} shapes, random-init weights, {args.batch} synthetic {T}x{H}x{W} clips per GPU per step "
                            f"(BASELINE.json configs[1])", "global_batch": args.batch * world, "per_gpu_batch": args.batch,
                "parallelism": f"clip-sharded x{world}, one all-gather of packed detections per step",
                "l2": "inputs (201 MB of clips per step) and activations (GBs) exceed the 126 MB L2; no flush needed"}

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        per_step = max(1, min(args.batch, 2))
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        clips = O.make_clips(per_step, T, H, W, seed=2)
        for _ in range(max(1, min(args.warmup, 1))):
            O.forward(cfg, sd, clips, None)
        steps = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(steps):
            O.forward(cfg, sd, clips, None)
        dt = time.perf_counter() - t0
        v = per_step * steps / dt
        sample = f"{steps} steps of {per_step} clips {T}x{H}x{W} (bounded sample of the {args.batch}-clip step), {dt:.1f} s"
        emit({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                          "warmup": 1, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload,
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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
    model, _, _ = tuber_b200.build_model(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    B = args.batch
    lo, _ = tuber_b200.shard_range(B * world, rank, world)
    clips = O.make_clips(B, T, H, W, seed=2 + rank).cuda()
    L, Q, NC = model.dec_layers, model.num_queries, model.num_class_out
    out = {"pred_logits": torch.empty((B, L, Q, NC), device="cuda"), "pred_boxes": torch.empty((B, L, Q, 4), device="cuda"),
           "pred_logits_b": torch.empty((B, L, Q, 3) if model.dataset_mode == "ava" else (B, 2), device="cuda")}
    model.use_cuda_graph(not args.no_graph)

    def step():
        model.forward_raw(clips, None, out)
        if world > 1:
            last = {k: (v[:, -1] if v.dim() == 4 else v) for k, v in out.items()}
            tuber_b200.gather_detections(tuber_b200.pack_detections(last), B * world)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    launches_per_step = lib.tuber_last_launches(model.plan())

    # ---- e2e: host buffers through the public host entry points (tuber_forward_host_submit / _wait): every step
    # copies that step's clips from pinned host memory (two alternating buffers) and reads its detections back;
    # the copy of step i+1 is in flight while step i computes
    h_clips = [torch.empty((B, 3, T, H, W), dtype=torch.float32).pin_memory() for _ in range(2)]
    for hc in h_clips:
        hc.copy_(clips)
    h_out = [model._host_out(B) for _ in range(2)]

    def e2e_run(n):
        model.forward_host_submit(0, h_clips[0], None, h_out[0])
        for i in range(1, n):
            model.forward_host_submit(i & 1, h_clips[i & 1], None, h_out[i & 1])
            model.forward_host_wait((i - 1) & 1)
        model.forward_host_wait((n - 1) & 1)

    e2e_run(3)
    sync_all()
    e2e_steps = max(4, min(args.steps, 50))
    t0 = time.perf_counter()
    e2e_run(e2e_steps)                              # returns after the last step's D2H copies have completed
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = B * world * e2e_steps / float(e2e_s.item())
    h2d = h_clips[0].numel() * 4
    d2h = sum(v.numel() for v in h_out[0].values()) * 4
    # the same, one synchronous call per step (no overlap), for reference
    for _ in range(2):
        model.forward_host(h_clips[0], None, h_out[0])
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        model.forward_host(h_clips[0], None, h_out[0])
    e2e_sync_val = B * e2e_steps / (time.perf_counter() - t0)
    # the same pipeline fed with decoded uint8 frames (B,T,H,W,3): 3 bytes per pixel cross the bus and the reference's ToTensor +
    # Normalize run on the device (tuber_forward_host_u8_submit) -- reported beside e2e, not instead of it
    h_frames = [O.make_frames_u8(B, T, H, W, seed=20 + rank + 100 * i).pin_memory() for i in range(2)]

    def e2e_u8_run(n):
        model.forward_host_u8_submit(0, h_frames[0], None, h_out[0])
        for i in range(1, n):
            model.forward_host_u8_submit(i & 1, h_frames[i & 1], None, h_out[i & 1])
            model.forward_host_wait((i - 1) & 1)
        model.forward_host_wait((n - 1) & 1)

    e2e_u8_run(3)
    sync_all()
    t0 = time.perf_counter()
    e2e_u8_run(e2e_steps)
    e2e_u8_s = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_u8_s, op=dist.ReduceOp.MAX)
    e2e_u8_val = B * world * e2e_steps / float(e2e_u8_s.item())

    # ---- the metric's other backbone: BASELINE.json quotes "clips/sec (32x256x256, CSN-152)", its configs[2] is
    # TubeR_CSN152_AVA21.yaml sharded over the GPUs -- same batch per GPU, same timing rules, reported beside the headline workload
    also = []
    if not args.no_also and args.config != "TubeR_CSN152_AVA21.yaml":
        cfg2 = tuber_b200.load_cfg("TubeR_CSN152_AVA21.yaml")
        model2, _, _ = tuber_b200.build_model(cfg2)
        model2.load_state_dict(O.make_state_dict(cfg2, seed=0, bn="random"), strict=True)
        model2 = model2.cuda().eval()
        model2.use_cuda_graph(not args.no_graph)
        out2 = {"pred_logits": torch.empty((B, model2.dec_layers, model2.num_queries, model2.num_class_out), device="cuda"),
                "pred_boxes": torch.empty((B, model2.dec_layers, model2.num_queries, 4), device="cuda"),
                "pred_logits_b": torch.empty((B, model2.dec_layers, model2.num_queries, 3), device="cuda")}
        for _ in range(warmup):
            model2.forward_raw(clips, None, out2)
        steps2 = max(5, min(args.steps, 30))
        sync_all()
        e0.record()
        for _ in range(steps2):
            model2.forward_raw(clips, None, out2)
        e1.record()
        sync_all()
        ms2 = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        also.append({"workload": f"TubeR_CSN152_AVA21.yaml shapes (BASELINE.json configs[2] backbone), {B} synthetic {T}x{H}x{W} clips per GPU per step",
                     "value": B * world * steps2 / (float(ms2.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(ms2.item()) / steps2,
                     "steps": steps2, "launches_per_step": lib.tuber_last_launches(model2.plan())})
        del model2, out2
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel profile of one forward (CUDA events around every launch)
    _lib.check(lib.tuber_set_kernel_profiling(model.plan(), 1))
    model.forward_raw(clips, None, out)
    torch.cuda.synchronize()
    n = C.c_int32()
    _lib.check(lib.tuber_get_kernel_profile(model.plan(), None, 0, C.byref(n)))
    stats = (_lib.TuberKernelStat * n.value)()
    _lib.check(lib.tuber_get_kernel_profile(model.plan(), stats, n.value, C.byref(n)))
    _lib.check(lib.tuber_set_kernel_profiling(model.plan(), 0))
    peaks = _peaks()
    kernels = []
    tot_ms = sum(s.ms for s in stats) or 1.0
    for s in sorted(stats, key=lambda s: -s.ms):
        gbs = s.bytes / (s.ms * 1e-3) / 1e9 if s.ms > 0 else 0.0
        tfs = s.flops / (s.ms * 1e-3) / 1e12 if s.ms > 0 else 0.0
        kernels.append({"kernel": s.name.decode(), "launches": s.launches, "ms": round(s.ms, 4), "share": round(s.ms / tot_ms, 4),
                        "GB/s": round(gbs, 1), "TFLOP/s": round(tfs, 2), "hbm_frac": round(gbs / peaks["hbm_gbs"], 4),
                        "tensor_frac": round(tfs / peaks["bf16_tflops"], 4)})
    top = max(stats, key=lambda s: s.ms)
    t_hbm = top.bytes / (peaks["hbm_gbs"] * 1e9)
    t_tc = 3.0 * top.flops / (peaks["bf16_tflops"] * 1e12) if (b"tcgen05" in top.name or b"gemm_bf16x3" in top.name or b"gemm2_bf16x3" in top.name) else 0.0   # bf16x3: 3 MMA passes
    if t_hbm >= t_tc:
        ach = top.bytes / (top.ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"]}
    else:
        ach = 3.0 * top.flops / (top.ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"]}
    traffic, traffic_src = _ncu_traffic(top.name.decode())
    roof.update({"kernel": top.name.decode(), "launches_per_step": top.launches, "avg_launch_ms": top.ms / max(1, top.launches),
                 "algorithmic_bytes_per_launch": top.bytes / max(1, top.launches), "peak_source": peaks["src"], "traffic": traffic,
                 "traffic_source": traffic_src,
                 "how": "CUDA events around every launch of one extra forward (same stream, same buffers, after the timed region)"})
    stage_ms = model.stage_times_ms(clips)

    cpu = None
    if not args.no_cpu_baseline and world == 1:                 # rank 0 at N = 1 only
        v, cores, sample = _cpu_reference(cfg, sd, T, H, W, budget_s=15.0, max_clips=32)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    value = B * world * args.steps / (ms_total * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload, "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "forward_host_submit/_wait (double-buffered: H2D of step i+1 overlaps step i)",
                    "synchronous_per_call_rank0": e2e_sync_val, "rank0_cpus_bound": numa},
            "e2e_u8": {"value": e2e_u8_val, "unit": UNIT, "h2d_bytes_per_step": h_frames[0].numel(), "d2h_bytes_per_step": d2h,
                       "steps": e2e_steps, "api": "forward_host_u8_submit/_wait: uint8 (B,T,H,W,3) frames in pinned host memory, "
                                                  "ToTensor + Normalize on the device (+1 launch per step)"},
            "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
            "cuda_graph": not args.no_graph, "roofline": roof, "cpu_baseline": cpu, "also": also, "kernels": kernels,
            "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
            "arithmetic": "fp32 storage; GEMMs = 3-pass bf16 split (hi*hi+hi*lo+lo*hi) on tcgen05 with fp32 TMEM accumulation"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
