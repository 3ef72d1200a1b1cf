"""Labeling rate of the BASELINE.json configurations that are parity cases, not bench lines (configs[2], [3]):
`clip_ft` / `clip_multiscale_ensemble` adapter heads and multi-instruction labeling, next to plain `clip`, on resident
frames through the same RewardLabeler the drop-in uses.

  python tools/config_rates.py [frames]        # writes gpurun_out/config_rates.json
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import os as _os
_os.environ.setdefault("ARP_ALLOW_STANDIN_TOKENIZER", "1")   # random-init weights: the deterministic stand-in token ids
from arp_b200.instructions import get_clip_instruct  # noqa: E402
from arp_b200.label_reward import RewardLabeler  # noqa: E402
from arp_b200.weights import random_adapter_state_dict, random_clip_state_dict  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 6144
dev = torch.device("cuda", 0)
clip_sd = random_clip_state_dict("ViT-B/16", 0, "cpu")
adapter_sd = random_adapter_state_dict("ViT-B/16", seed=1, device="cpu", clip_sd=clip_sd)
MAZE2 = ["The agent must navigate a maze to find the line.", "The agent navigate to the yellow object."]
MAZE3 = MAZE2 + ["navigate a maze to collect the yellow cheese."]
cases = [
    ("clip_256_1text (configs[1])", dict(model_type="clip", text=get_clip_instruct("coinrun")), 256),
    ("clip_64_3texts_first (configs[3])", dict(model_type="clip", text=MAZE3), 64),
    ("clip_64_3texts_mean", dict(model_type="clip", text=MAZE3, reduce="mean"), 64),
    ("clip_ft_64 (configs[2])", dict(model_type="clip_ft", text=get_clip_instruct("maze"), model_ckpt_dir=adapter_sd), 64),
    ("clip_ft_256_2texts", dict(model_type="clip_ft", text=MAZE2, model_ckpt_dir=adapter_sd), 256),
    ("clip_multiscale_ensemble_64 (configs[2])", dict(model_type="clip_multiscale_ensemble", text=get_clip_instruct("maze"),
                                                     model_ckpt_dir=adapter_sd), 64),
]
out = {}
for name, kw, size in cases:
    lab = RewardLabeler(frame_hw=(size, size), clip_state_dict=clip_sd, max_batch=1024, **kw)
    e = lab.engine
    ob = torch.randint(0, 256, (T, 1, size, size, 3), dtype=torch.uint8, device=dev)
    off = torch.arange(0, T + 1, 64, dtype=torch.int64, device=dev)
    for _ in range(2):
        e.label(ob, off, 8)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(3):
        e.label(ob, off, 8)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 3
    e.profile_begin()
    e.label(ob, off, 8)
    prof = e.profile_end()
    out[name] = {"frames": T, "frame_px": size, "ms": ms, "frames_per_s": T / ms * 1e3,
                 "by_class_ms": {k: round(v["total_ms"], 2) for k, v in prof.items() if v["launches"]}}
    print(name, json.dumps(out[name]), flush=True)
    lab.close()
    del ob
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "config_rates.json").write_text(json.dumps(out, indent=1))
