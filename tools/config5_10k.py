"""BASELINE.json configs[4]: episode-sharded labeling of 10k synthetic episodes (~1.43 M frames of 64x64) on N GPUs with one
NCCL gather of the reward / return-to-go arrays. Frames are generated on each rank's device from a generator keyed by the
episode index (so any sharding sees the same frames). Rank 0 checks the gathered arrays:
  * episode boundaries / scan order / window stack: the C oracle's scan over OUR gathered rewards must equal OUR gathered
    return-to-go bit for bit, for all 10k episodes;
  * sharding invariance: a sample of episodes re-labeled on rank 0 alone gives the same rewards bit for bit.

  torchrun --nproc-per-node N tools/config5_10k.py        (writes gpurun_out/config5_n<N>.json on rank 0)
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import os as _os
_os.environ.setdefault("ARP_ALLOW_STANDIN_TOKENIZER", "1")   # random-init weights: the deterministic stand-in token ids
from arp_b200 import capi  # noqa: E402
from arp_b200.sharding import gather_rows, partition_episodes  # noqa: E402
from arp_b200.text_tower import clip_text_embedding  # noqa: E402
from arp_b200.tokenizer import tokenize  # noqa: E402
from arp_b200.weights import random_clip_state_dict  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N_EPS, SIZE, F = 10_000, 64, 8
lens = np.random.default_rng(5).integers(32, 256, size=N_EPS)
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
shards = partition_episodes(off, world)
e_lo, e_hi = shards[rank]
lo, hi = int(off[e_lo]), int(off[e_hi])
T = hi - lo


def episode_frames(e: int) -> torch.Tensor:
    g = torch.Generator(device=dev).manual_seed(1_000_003 * 7 + e)
    return torch.randint(0, 256, (int(lens[e]), 1, SIZE, SIZE, 3), dtype=torch.uint8, device=dev, generator=g)


ob = torch.empty(T, 1, SIZE, SIZE, 3, dtype=torch.uint8, device=dev)
for e in range(e_lo, e_hi):
    ob[int(off[e]) - lo:int(off[e + 1]) - lo] = episode_frames(e)
sd = random_clip_state_dict("ViT-B/16", seed=0, device="cpu")
eng = capi.Engine(device=local, patch=16, in_h=SIZE, in_w=SIZE, max_batch=512)
eng.load_state_dict(sd)
emb, scale = clip_text_embedding(sd, tokenize(["the goal is to collect the coin."], standin=True), dev)
eng.set_text(emb, scale)
off_local = torch.from_numpy(off[e_lo:e_hi + 1] - lo).to(dev)
rows = [int(off[b] - off[a]) for a, b in shards]


def step():
    r, g, rs, gs = eng.label(ob, off_local, F)
    both = torch.stack([r, g], dim=1)
    return gather_rows(both, rows, dst=0) if world > 1 else both


step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
full = step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
secs = torch.tensor([time.perf_counter() - t0], device=dev)
if world > 1:
    dist.all_reduce(secs, op=dist.ReduceOp.MAX)
if rank == 0:
    from oracle import cport
    full = full.cpu().numpy()
    r_all, g_all = np.ascontiguousarray(full[:, 0]), np.ascontiguousarray(full[:, 1])
    ok_scan = True
    for a, b in zip(off[:-1], off[1:]):
        ok_scan &= np.array_equal(cport.discount_cumsum(r_all[a:b]), g_all[a:b])
    # sharding invariance on episodes from every shard
    sample = sorted({s[0] for s in shards if s[1] > s[0]} | {s[1] - 1 for s in shards if s[1] > s[0]} | {N_EPS // 2})
    ok_inv = True
    for e in sample:
        fr = episode_frames(e)
        r1 = eng.compute_reward(fr).cpu().numpy()
        ok_inv &= np.array_equal(r1, r_all[off[e]:off[e + 1]])
    out = {"config": "BASELINE configs[4]: 10k episodes, 64x64, ViT-B/16 clip", "n_gpus": world, "episodes": N_EPS,
           "frames": int(off[-1]), "seconds": float(secs), "frames_per_s": float(off[-1] / float(secs)),
           "rows_per_rank": rows, "scan_bit_exact_all_episodes": bool(ok_scan),
           "rewards_identical_to_single_rank_on_sample": bool(ok_inv), "sample_episodes": [int(x) for x in sample]}
    print(json.dumps(out))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"config5_n{world}.json").write_text(json.dumps(out, indent=1))
eng.close()
if world > 1:
    dist.destroy_process_group()
