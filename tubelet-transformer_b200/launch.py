"""One process per GPU: the launcher of the evaluation scripts (reference pipelines/launch.py:20-50).

``spawn_workers(main, cfg)`` keeps the reference's call form -- ``eval_tuber_ava.py`` ends with ``spawn_workers(main_worker, cfg)``
(:58) -- and its ``DDP_CONFIG`` surface: DISTRIBUTED, WORLD_SIZE (nodes), WORLD_RANK (this node), GPU_WORLD_SIZE / GPU_WORLD_RANK
(filled in here), GPU (this process's device), DIST_URL, DIST_BACKEND, AUTO_RANK_MATCH + WOLRD_URLS (sic).  Differences, all on the
host side of the path:

* every worker is bound to the CPUs NVML reports as local to its GPU *before* it allocates anything, so the pinned staging buffers of
  the host entry points (tuber_forward_host*) land on the GPU's own NUMA node (the 8-GPU host path is bound by host -> device copies);
* the rank of this node is matched against the machine's own addresses without opening a socket to a public resolver
  (the reference connects to 8.8.8.8, launch.py:8-17) -- the GPU boxes have no route out;
* ``nprocs`` / ``device_count`` can be given explicitly (CPU tests run two gloo workers on a box without GPUs).
"""
from __future__ import annotations

import os
import socket
from typing import Callable, Iterable, Optional

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def local_addresses() -> set:
    """Addresses of this machine that a peer could have listed in WOLRD_URLS."""
    out = {"127.0.0.1", "localhost"}
    try:
        host = socket.gethostname()
        out.add(host)
        out.update(info[4][0] for info in socket.getaddrinfo(host, None))
    except OSError:
        pass
    return out


def match_node_rank(urls: Iterable[str]) -> int:
    """Index of this machine in the node list, -1 when absent (launch.py:8-17)."""
    mine = local_addresses()
    for i, url in enumerate(urls):
        if url in mine:
            return i
    return -1


def bind_to_gpu_cpus(index: int) -> Optional[int]:
    """Restrict this process to the CPUs local to GPU `index`; returns how many, or None when NVML cannot say."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = ((os.cpu_count() or 1) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1} & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:  # noqa: BLE001 -- no NVML, no GPU, or a container that forbids it: run unbound
        pass
    return None


def main_worker(gpu: int, ngpus_per_node: int, main: Callable, cfg) -> None:
    """One worker (launch.py:37-50): fill in the per-process DDP_CONFIG fields, join the process group, run ``main(cfg)``."""
    ddp = cfg.DDP_CONFIG
    ddp.GPU = gpu
    if torch.cuda.is_available():
        bind_to_gpu_cpus(gpu)
        torch.cuda.set_device(gpu)
        torch.backends.cudnn.benchmark = True                      # launch.py:39 (only the reference-side modules use cuDNN)
    if ddp.DISTRIBUTED:
        ddp.GPU_WORLD_RANK = ddp.WORLD_RANK * ngpus_per_node + gpu
        kw = {}
        if ddp.DIST_BACKEND == "nccl" and torch.cuda.is_available():
            kw["device_id"] = torch.device("cuda", gpu)
        dist.init_process_group(backend=ddp.DIST_BACKEND, init_method=ddp.DIST_URL, world_size=ddp.GPU_WORLD_SIZE,
                                rank=ddp.GPU_WORLD_RANK, **kw)
    try:
        main(cfg)
    finally:
        if ddp.DISTRIBUTED and dist.is_initialized():
            dist.destroy_process_group()


def spawn_workers(main: Callable, cfg, nprocs: Optional[int] = None) -> None:
    """``spawn_workers(main_worker, cfg)`` of the reference's entry scripts (launch.py:20-34)."""
    ddp = cfg.DDP_CONFIG
    if ddp.AUTO_RANK_MATCH:
        urls = list(ddp.WOLRD_URLS)
        if not urls or urls[0] not in ddp.DIST_URL or len(urls) != ddp.WORLD_SIZE:
            raise ValueError("DDP_CONFIG: WOLRD_URLS must list WORLD_SIZE nodes and start with the host of DIST_URL")
        ddp.WORLD_RANK = match_node_rank(urls)
        if ddp.WORLD_RANK < 0:
            raise RuntimeError(f"this machine ({sorted(local_addresses())}) is not in DDP_CONFIG.WOLRD_URLS {urls}")
    ngpus = nprocs if nprocs is not None else torch.cuda.device_count()
    if ddp.DISTRIBUTED:
        if ngpus < 1:
            raise RuntimeError("spawn_workers: no CUDA device (pass nprocs= for CPU-side tests)")
        ddp.GPU_WORLD_SIZE = ngpus * ddp.WORLD_SIZE
        mp.spawn(main_worker, nprocs=ngpus, args=(ngpus, main, cfg))
    else:
        main_worker(ddp.GPU, ngpus, main, cfg)
