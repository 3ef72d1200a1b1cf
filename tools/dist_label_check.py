"""Dev/validation tool: run the drop-in label_reward() under torchrun (NCCL) and compare the sharded result with a
single-rank run of the same dataset. Usage: torchrun --nproc-per-node N tools/dist_label_check.py"""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import os as _os
_os.environ.setdefault("ARP_ALLOW_STANDIN_TOKENIZER", "1")   # random-init weights: the deterministic stand-in token ids
from arp_b200.label_reward import label_reward  # noqa: E402
from arp_b200.store import NpyStore  # noqa: E402
from arp_b200.synth import make_dataset, write_dataset  # noqa: E402
from arp_b200.weights import random_clip_state_dict  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sd = random_clip_state_dict("ViT-B/32", seed=0, device="cpu")
base = Path(tempfile.gettempdir()) / "arp_dist_check"
if rank == 0:
    import shutil
    shutil.rmtree(base, ignore_errors=True)
    for name in ("sharded", "single"):
        s = NpyStore(base / name, "w")
        write_dataset(s, make_dataset(n_episodes=37, len_lo=5, len_hi=60, size=64, num_frames=4, seed=11, tail_rows=7))
        s.close()
dist.barrier()
kw = dict(model_type="clip", clip_state_dict=sd, arch="ViT-B/32", max_batch=64, slab_frames=300, env_type="none")
label_reward("coinrun", "hard", 500, 0, "the goal is to collect the coin.", str(base), data_path=str(base / "sharded"), **kw)
if rank == 0:
    label_reward("coinrun", "hard", 500, 0, "the goal is to collect the coin.", str(base), data_path=str(base / "single"),
                 distributed=False, **kw)
    a, b = NpyStore(base / "sharded", "r"), NpyStore(base / "single", "r")
    ok = True
    for k in ("ob_clip_reward", "ob_clip_pos_rtg"):
        x, y = np.array(a[k][:]), np.array(b[k][:])
        same = x.shape == y.shape and np.array_equal(x, y)
        ok &= same
        print(f"{k}: shape {x.shape} identical to single-rank run: {same}")
    print("DIST_LABEL_OK" if ok else "DIST_LABEL_MISMATCH", f"world={world}")
dist.barrier()
dist.destroy_process_group()
