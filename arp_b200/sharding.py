"""Episode sharding across GPUs. The reference labeler is single-process
(label_reward.py:265); episodes are independent, so ranks take contiguous episode ranges balanced by
frame count and the only communication is one gather of the per-frame reward / return-to-go arrays.
Contiguous ranges keep output order = file order for the append-by-episode writer (:276-287).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def partition_episodes(ep_offsets: np.ndarray, world_size: int) -> list[tuple[int, int]]:
    """Split episodes [0, n) into `world_size` contiguous ranges with ~equal frame counts.

    Cut points are the episode boundaries closest to k*T/world (prefix-sum search); ranges may be
    empty when there are fewer episodes than ranks."""
    off = np.asarray(ep_offsets, dtype=np.int64)
    n = len(off) - 1
    total = int(off[-1] - off[0])
    cuts = [0]
    for k in range(1, world_size):
        target = off[0] + total * k / world_size
        j = int(np.searchsorted(off, target, side="left"))
        if j > 0 and j <= n and abs(off[j - 1] - target) <= abs(off[min(j, n)] - target):
            j -= 1
        cuts.append(min(max(j, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world_size)]


def gather_rows(local: torch.Tensor, rows_per_rank: list[int], dst: int = 0, group=None):
    """Variable-length gather along dim 0 to rank `dst` through one padded all_gather.

    `local` is [n_r, ...] on this rank's device (NCCL) or CPU (gloo). Returns the concatenation in rank
    order on `dst`, None elsewhere. <= 8 B/frame: latency-bound, so a single collective is the right size."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert len(rows_per_rank) == world and local.shape[0] == rows_per_rank[rank]
    pad = max(max(rows_per_rank), 1)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty((world * pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    if rank != dst:
        return None
    parts = [out[r * pad: r * pad + rows_per_rank[r]] for r in range(world)]
    return torch.cat(parts, dim=0)
