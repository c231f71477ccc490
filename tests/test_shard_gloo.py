"""N>1 host logic on CPU: frame sharding + the output gather with the gloo
backend, world_size 2 and 3 (the GPU box runs the same code over NCCL)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rawspeed_b200 import shard


def test_partition_covers_every_frame_once():
    for n in (0, 1, 7, 256):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                f = shard.frames_of_rank(n, r, world)
                assert all(shard.owner_of_frame(i, world) == r for i in f)
                assert len(f) <= shard.max_frames_per_rank(n, world)
                seen += f
            assert sorted(seen) == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nframes, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard.frames_of_rank(nframes, rank, world)
        # "decode": frame i is a 4x6 uint16 image filled with a function of i
        local = torch.stack([torch.full((4, 6), 1000 + i, dtype=torch.int16) for i in mine]) \
            if mine else torch.zeros((0, 4, 6), dtype=torch.int16)
        full = shard.gather_frames(local, nframes, dist)
        assert full.shape == (nframes, 4, 6)
        for i in range(nframes):
            assert int(full[i, 0, 0]) == 1000 + i and bool((full[i] == 1000 + i).all())
        # the copy-free form bench.py times: preallocated [world, per, ...] result, the
        # collective's own layout (frame r + k*world at [r, k])
        per = shard.max_frames_per_rank(nframes, world)
        buf = torch.full((world, per, 4, 6), -1, dtype=torch.int16)
        got = shard.gather_frames(local, nframes, dist, out=buf, reorder=False)
        assert got.data_ptr() == buf.data_ptr()
        for i in range(nframes):
            assert bool((got[i % world, i // world] == 1000 + i).all())
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nframes", [(2, 7), (2, 8), (3, 5)])
def test_gather_over_gloo(world, nframes):
    ctx = mp.get_context("spawn")
    ok = ctx.Array("i", [0] * world)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nframes, ok)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs)
    assert list(ok) == [1] * world
