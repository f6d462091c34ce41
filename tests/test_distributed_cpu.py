"""CPU, world_size 2 over gloo: the clip sharding + single all-gather of packed detections."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import tuber_b200


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        full = {"pred_logits": torch.randn(n_clips, 15, 80, generator=g), "pred_boxes": torch.rand(n_clips, 15, 4, generator=g),
                "pred_logits_b": torch.randn(n_clips, 15, 3, generator=g)}
        lo, hi = tuber_b200.shard_range(n_clips, rank, world)
        local = {k: v[lo:hi] for k, v in full.items()}          # what this rank's forward would produce
        gathered = tuber_b200.gather_detections(tuber_b200.pack_detections(local), n_clips)
        back = tuber_b200.unpack_detections(gathered, 15, 80, ava=True)
        ok = all(torch.equal(back[k], full[k]) for k in full)
        q.put((rank, ok, tuple(gathered.shape)))
    finally:
        dist.destroy_process_group()


def _run(n_clips):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ok, shape in res:
        assert ok, rank
        assert shape == (n_clips, 15 * 87)


def test_gather_even_shards():
    _run(8)


def test_gather_uneven_shards():
    _run(5)


def test_gather_single_process_is_identity():
    x = torch.randn(3, 10)
    assert tuber_b200.gather_detections(x, 3) is x
