"""world_size-2/3 gloo coverage of the N>1 host path: partition + the single padded gather, and the whole
label_reward() orchestration (shards, slabs, gather, rank-0 write) against a single-process run."""
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


# ------------------------------------------------------------------------------------------------
# the whole N>1 host path of label_reward(): partition -> per-rank slabs -> one gather -> rank-0 write.
# The GPU labeler is replaced by a deterministic per-frame stand-in (test-only; the product has no CPU path).
# ------------------------------------------------------------------------------------------------
class _StubLabeler:
    goal = False

    class _Eng:
        device = torch.device("cpu")

    def __init__(self, *a, **k):
        self.engine = self._Eng()

    def label_slab(self, ob, ep_offsets, num_frames):
        ob = np.asarray(ob)
        last = ob[:, -1] if ob.ndim == 5 else ob
        r = last.reshape(len(last), -1).astype(np.float64).mean(axis=1).astype(np.float32)   # depends on the frame only
        g = np.zeros_like(r)
        for lo, hi in zip(ep_offsets[:-1], ep_offsets[1:]):
            acc = np.float32(0)
            for t in range(hi - 1, lo - 1, -1):                 # label_reward.py:247-254 order
                acc = np.float32(r[t] + acc)
                g[t] = acc
        idx = np.arange(len(r))
        start = np.repeat(ep_offsets[:-1], np.diff(ep_offsets))
        win = np.maximum(start[:, None], idx[:, None] - (num_frames - 1 - np.arange(num_frames))[None, :])
        return r, g, r[win], g[win]

    def close(self):
        pass


def _make_store(path, seed=3):
    from arp_b200.store import NpyStore
    from arp_b200.synth import make_dataset, write_dataset
    data = make_dataset(n_episodes=11, len_lo=2, len_hi=17, size=8, num_frames=4, seed=seed, tail_rows=3)
    s = NpyStore(path, "w")
    write_dataset(s, data)
    s.close()


def _label(path, distributed):
    import arp_b200.label_reward as lr
    real = lr.RewardLabeler
    lr.RewardLabeler = _StubLabeler
    try:
        lr.label_reward("coinrun", "hard", 500, 0, "the goal is to collect the coin.", ".", data_path=str(path),
                        model_type="clip", env_type="none", slab_frames=16, distributed=distributed)
    finally:
        lr.RewardLabeler = real          # later tests in this process get the real labeler back


def _worker_label(rank, world, port, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _label(path, True)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_label_reward_host_path_gloo_equals_single_process(tmp_path, world):
    from arp_b200.store import NpyStore
    single, multi = tmp_path / "single", tmp_path / "multi"
    _make_store(single)
    _make_store(multi)
    _label(single, False)
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker_label, args=(r, world, port, multi)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    a, b = NpyStore(single, "r"), NpyStore(multi, "r")
    keys = [k for k in a.keys() if k.startswith("ob_")]
    assert keys == ["ob_clip_pos_rtg", "ob_clip_reward"] and sorted(k for k in b.keys() if k.startswith("ob_")) == keys
    for k in keys:
        x, y = np.array(a[k][:]), np.array(b[k][:])
        assert x.shape == y.shape and x.dtype == y.dtype == np.float32 and np.array_equal(x, y), k
    assert np.array(a["ob_clip_reward"][:]).shape[0] == int(np.nonzero(np.array(a["done"][:, -1]))[0][-1]) + 1   # tail rows unlabeled
    a.close()
    b.close()
