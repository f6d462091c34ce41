"""CPU: the launcher (tuber_b200/launch.py; reference pipelines/launch.py:20-50) with two gloo workers."""
import os
import socket

import pytest
import torch
import torch.distributed as dist

import tuber_b200
from tuber_b200 import launch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _main(cfg):
    # what an entry script's main_worker sees: its rank fields are filled in and the process group is up
    ddp = cfg.DDP_CONFIG
    assert dist.is_initialized() and dist.get_world_size() == ddp.GPU_WORLD_SIZE == 2
    assert dist.get_rank() == ddp.GPU_WORLD_RANK == ddp.GPU
    lo, hi = tuber_b200.shard_range(5, ddp.GPU_WORLD_RANK, ddp.GPU_WORLD_SIZE)
    rows = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 3)
    got = tuber_b200.gather_detections(rows, 5)
    assert torch.equal(got[:, 0], torch.arange(5, dtype=torch.float32))
    with open(os.path.join(cfg.CONFIG.LOG.BASE_PATH, f"{ddp.GPU_WORLD_RANK}.txt"), "w") as f:
        f.write("ok")


def test_spawn_workers_two_gloo_ranks(tmp_path):
    cfg = tuber_b200.load_cfg("TubeR_CSN50_AVA21.yaml", ["DDP_CONFIG.DIST_BACKEND", "gloo", "DDP_CONFIG.DIST_URL", f"tcp://127.0.0.1:{_free_port()}",
                                                        "CONFIG.LOG.BASE_PATH", str(tmp_path)])
    launch.spawn_workers(_main, cfg, nprocs=2)
    assert sorted(os.listdir(tmp_path)) == ["0.txt", "1.txt"]
    assert cfg.DDP_CONFIG.WORLD_RANK == 0 and cfg.DDP_CONFIG.GPU_WORLD_SIZE == 2


def test_node_rank_matching_and_errors():
    assert launch.match_node_rank(["10.255.255.1", "127.0.0.1"]) == 1
    assert launch.match_node_rank(["10.255.255.1"]) == -1
    cfg = tuber_b200.load_cfg("TubeR_CSN50_AVA21.yaml", ["DDP_CONFIG.WOLRD_URLS", ["10.255.255.1"], "DDP_CONFIG.DIST_URL", "tcp://10.255.255.1:1"])
    with pytest.raises(RuntimeError):
        launch.spawn_workers(_main, cfg, nprocs=2)
    cfg = tuber_b200.load_cfg("TubeR_CSN50_AVA21.yaml", ["DDP_CONFIG.WORLD_SIZE", 2])
    with pytest.raises(ValueError):
        launch.spawn_workers(_main, cfg, nprocs=2)
