"""Dev tool: the LayerNorm-folded pipeline (ARP_LN_FOLD=1) against the standalone-LN pipeline (ARP_LN_FOLD=0)
and against the fp32 torch model (oracle shim) on the same frames, with NON-trivial LN gamma/beta and a
shifted residual stream, plus per-chunk timing of both. Run under gpurun.

  python tools/fold_check.py [--batch 512] [--chunks 4]
"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "shims"))
from arp_b200 import capi  # noqa: E402
from arp_b200.weights import random_clip_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--chunks", type=int, default=4)
ap.add_argument("--px", type=int, default=64)
args = ap.parse_args()

dev = torch.device("cuda", 0)
sd = random_clip_state_dict("ViT-B/16", seed=0, device="cpu")
g = torch.Generator().manual_seed(7)
for k in list(sd):
    if ".ln_" in k or "ln_pre" in k or "ln_post" in k:
        if k.endswith("weight"):
            sd[k] = sd[k] * (1.0 + 0.5 * torch.randn(sd[k].shape, generator=g))
        else:
            sd[k] = sd[k] + 0.3 * torch.randn(sd[k].shape, generator=g)
# a residual stream with a mean and a few large channels (what pretrained CLIP looks like)
sd["visual.class_embedding"] = sd["visual.class_embedding"] + 0.5
pe = sd["visual.positional_embedding"].clone()
pe[:, :4] += 8.0
pe += 0.2
sd["visual.positional_embedding"] = pe

text = torch.nn.functional.normalize(torch.randn(1, 512, generator=g), dim=1)
T = 48
ob = torch.randint(0, 256, (T, 1, args.px, args.px, 3), dtype=torch.uint8, generator=g)


def run(fold: int):
    os.environ["ARP_LN_FOLD"] = str(fold)
    e = capi.Engine(device=0, patch=16, in_h=args.px, in_w=args.px, max_batch=32)
    e.load_state_dict(sd)
    e.set_text(text, 14.2857)
    feat = e.encode_image(ob.to(dev)).cpu()
    e.close()
    return feat


f0, f1, f2 = run(0), run(1), run(2)

import clip as clip_shim  # noqa: E402
from oracle import port  # noqa: E402
model = clip_shim.build("ViT-B/16", seed=0)
model.load_state_dict(sd, strict=True)
model = model.float().eval().to(dev)
tf = port.transform_pil(False, args.px)
imgs = torch.stack([tf(ob[t, -1].numpy()) for t in range(T)]).to(dev)
with torch.no_grad():
    ref = model.encode_image(imgs).float().cpu()


def cosdist(a, b):
    a = torch.nn.functional.normalize(a.double(), dim=1)
    b = torch.nn.functional.normalize(b.double(), dim=1)
    return float((1 - (a * b).sum(1)).max())


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


print(f"feature relerr vs fp32 torch: no-fold {rel(f0, ref):.3e}  fold1 {rel(f1, ref):.3e}  fold2 {rel(f2, ref):.3e}")
print(f"max (1 - cos) vs fp32 torch:  no-fold {cosdist(f0, ref):.3e}  fold1 {cosdist(f1, ref):.3e}  fold2 {cosdist(f2, ref):.3e}")
ok = max(rel(f1, ref), rel(f2, ref)) < 2.0 * max(rel(f0, ref), 2e-3)
print("FOLD NUMERICS", "OK" if ok else "BAD")

# ---- timing: chunks of `batch` 256x256 frames, both pipelines ----
Tt = args.batch * args.chunks
obt = torch.randint(0, 256, (Tt, 1, 256, 256, 3), dtype=torch.uint8, device=dev)
off = torch.arange(0, Tt + 1, 64, dtype=torch.int64, device=dev)
sd0 = random_clip_state_dict("ViT-B/16", seed=0, device="cpu")
for fold in (0, 1, 2):
    os.environ["ARP_LN_FOLD"] = str(fold)
    e = capi.Engine(device=0, patch=16, in_h=256, in_w=256, max_batch=args.batch)
    e.load_state_dict(sd0)
    e.set_text(text, 14.2857)
    for _ in range(2):
        e.label(obt, off, 8)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(3):
        e.label(obt, off, 8)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 3
    e.profile_begin()
    e.label(obt, off, 8)
    prof = e.profile_end()
    print(f"fold={fold}: {ms:.2f} ms / {Tt} frames = {Tt / ms * 1e3:.0f} frames/s   by class ms: "
          + ", ".join(f"{k} {v['total_ms']:.1f}" for k, v in prof.items() if v["launches"]))
    e.close()
sys.exit(0 if ok else 1)
