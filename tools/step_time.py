"""Dev tool: steady-state time of the labeling step for a list of max_batch values (resident frames).
  python tools/step_time.py 256,384,512,768 [frames]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from arp_b200 import capi  # noqa: E402
from arp_b200.weights import random_clip_state_dict  # noqa: E402

batches = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "512").split(",")]
T = int(sys.argv[2]) if len(sys.argv) > 2 else 6144
dev = torch.device("cuda", 0)
sd = random_clip_state_dict("ViT-B/16", seed=0, device="cpu")
ob = torch.randint(0, 256, (T, 1, 256, 256, 3), dtype=torch.uint8, device=dev)
off = torch.arange(0, T + 1, 64, dtype=torch.int64, device=dev)
text = torch.nn.functional.normalize(torch.randn(1, 512), dim=1)
for B in batches:
    e = capi.Engine(device=0, patch=16, in_h=256, in_w=256, max_batch=B)
    e.load_state_dict(sd)
    e.set_text(text, 14.2857)
    for _ in range(2):
        e.label(ob, off, 8)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(4):
        e.label(ob, off, 8)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 4
    e.profile_begin()
    e.label(ob, off, 8)
    prof = e.profile_end()
    print(f"max_batch={B}: {ms:.2f} ms / {T} frames = {T / ms * 1e3:.0f} frames/s   by class ms: "
          + ", ".join(f"{k} {v['total_ms']:.1f}" for k, v in prof.items() if v["launches"]), flush=True)
    e.close()
