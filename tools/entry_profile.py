"""Dev tool: where does the drop-in's wall clock go? arp_label_host on pinned memory vs on a memory-mapped store through
the native stager (ARP_STAGER_THREADS), then label_reward() with a phase breakdown."""
import os
import shutil
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ.setdefault("ARP_ALLOW_STANDIN_TOKENIZER", "1")
from arp_b200 import capi  # noqa: E402
from arp_b200.label_reward import label_reward  # noqa: E402
from arp_b200.text_tower import clip_text_embedding  # noqa: E402
from arp_b200.tokenizer import tokenize  # noqa: E402
from arp_b200.weights import random_clip_state_dict  # noqa: E402

episodes = int(sys.argv[1]) if len(sys.argv) > 1 else 200
S, F = 2, 8
lens = np.random.default_rng(1).integers(32, 256, size=episodes)
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
T = int(off[-1])
base = "/dev/shm" if shutil.disk_usage("/dev/shm").free > T * S * 196608 * 1.2 else tempfile.gettempdir()
d = Path(tempfile.mkdtemp(prefix="arp_entry_", dir=base))
print(f"{T} rows, store under {d}", flush=True)
mm = np.lib.format.open_memmap(d / "ob.npy", mode="w+", dtype=np.uint8, shape=(T, S, 256, 256, 3))
rng = np.random.default_rng(0)
blk = rng.integers(0, 256, size=(512, S, 256, 256, 3), dtype=np.uint8)
for lo in range(0, T, 512):
    mm[lo:lo + 512] = blk[:min(512, T - lo)]
mm.flush()
del mm
done = np.zeros((T, F), np.float32)
done[off[1:] - 1, -1] = 1.0
np.save(d / "done.npy", done)

t0 = time.perf_counter()
sd = random_clip_state_dict("ViT-B/16", 0, "cpu")
eng = capi.Engine(device=0, patch=16, in_h=256, in_w=256, max_batch=1024)
t1 = time.perf_counter()
eng.load_state_dict(sd)
t2 = time.perf_counter()
emb, scale = clip_text_embedding(sd, tokenize(["the goal is to collect the coin."]), eng.device)
eng.set_text(emb, scale)
torch.cuda.synchronize()
t3 = time.perf_counter()
print(f"engine create {t1 - t0:.3f} s (incl. random weights)  load_state_dict {t2 - t1:.3f} s  text tower {t3 - t2:.3f} s", flush=True)

ob = np.load(d / "ob.npy", mmap_mode="r")
pinned = torch.from_numpy(np.ascontiguousarray(ob)).pin_memory()
for name, src in (("pinned", pinned), ("memmap", ob)):
    for threads in ((None,) if name == "pinned" else (2, 4, 8)):
        if threads:
            os.environ["ARP_STAGER_THREADS"] = str(threads)
        ob2 = np.load(d / "ob.npy", mmap_mode="r") if name == "memmap" else src      # fresh mapping: page tables cold
        for rep in range(2):
            t0 = time.perf_counter()
            eng.label_host(ob2, off, F)
            dt = time.perf_counter() - t0
            print(f"label_host {name:7s} threads={threads} rep={rep}: {dt:.3f} s  {T / dt:8.0f} frames/s", flush=True)
del pinned
eng.close()
os.environ["ARP_STAGER_THREADS"] = "4"
kw = dict(env_name="coinrun", distribution_mode="hard", num_levels=500, start_level=0, text="the goal is to collect the coin.",
          base_path=".", data_path=str(d), model_type="clip", clip_state_dict=sd, env_type="none", max_batch=1024, distributed=False)
for rep in range(2):
    t0 = time.perf_counter()
    label_reward(**kw)
    dt = time.perf_counter() - t0
    print(f"label_reward() rep={rep}: {dt:.3f} s  {T / dt:8.0f} frames/s", flush=True)
shutil.rmtree(d, ignore_errors=True)
