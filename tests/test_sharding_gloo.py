"""world_size-2 gloo coverage of the N>1 host path: partition + the single padded gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from arp_b200.sharding import gather_rows, partition_episodes


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lens = np.random.default_rng(0).integers(3, 40, size=23)
        off = np.concatenate([[0], np.cumsum(lens)])
        parts = partition_episodes(off, world)
        a, b = parts[rank]
        lo, hi = int(off[a]), int(off[b])
        # stand-in for this rank's labels: value = global row index, so order errors are visible
        local = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1).repeat(1, 2, 4)
        rows = [int(off[y] - off[x]) for x, y in parts]
        full = gather_rows(local, rows, dst=0)
        if rank == 0:
            ok = full.shape == (int(off[-1]), 2, 4) and torch.equal(full[:, 0, 0], torch.arange(int(off[-1]), dtype=torch.float32))
            q.put(bool(ok))
        else:
            assert full is None
        # empty shard on one rank
        rows2 = [5, 0] if world == 2 else [5] + [0] * (world - 1)
        loc2 = torch.ones(rows2[rank], 3)
        full2 = gather_rows(loc2, rows2, dst=0)
        if rank == 0:
            q.put(bool(full2.shape == (5, 3)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_gather_rows_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True and q.get(timeout=5) is True
