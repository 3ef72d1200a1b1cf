"""Dev tool: a short, fixed run of the hot path for ncu (launch list / --set full captures).

  python tools/profile_step.py [--batch 256] [--chunks 2] [--px 256]
Runs `chunks` chunks of `batch` frames of ViT-B/16 `clip` labeling with random-init weights.
Numbers printed under a profiler are never benchmark values.
"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from arp_b200 import capi  # noqa: E402
from arp_b200.weights import random_clip_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--chunks", type=int, default=2)
ap.add_argument("--px", type=int, default=256)
ap.add_argument("--head", type=int, default=capi.HEAD_CLIP)
args = ap.parse_args()

dev = torch.device("cuda", 0)
e = capi.Engine(device=0, patch=16, in_h=args.px, in_w=args.px, max_batch=args.batch, head=args.head)
e.load_state_dict(random_clip_state_dict("ViT-B/16", seed=0, device=dev))
e.set_text(torch.nn.functional.normalize(torch.randn(1, e.feat_dim), dim=1), 14.2857)
T = args.batch * args.chunks
ob = torch.randint(0, 256, (T, 1, args.px, args.px, 3), dtype=torch.uint8, device=dev)
off = torch.arange(0, T + 1, 64, dtype=torch.int64)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
e.label(ob, off, 8)
t1.record()
torch.cuda.synchronize()
print(f"frames={T} ms={t0.elapsed_time(t1):.3f} launches={e.launch_count}")
