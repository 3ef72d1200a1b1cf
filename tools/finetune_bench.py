"""Measurement for the widened row SURVEY.md §8(f)4: one adapter fine-tuning step (forward + backward + AdamW) at the
reference's batch size (finetune.py:27: 64 samples -> 192 images through the frozen CLIP tower, 256 goal-conditioned),
split into the frozen tower (native library) and the autograd half (PyTorch), next to the same step with the frozen
tower evaluated by PyTorch eager fp16 on the same GPU (the oracle's restatement of openai/CLIP, i.e. what the
reference runs: clip.load keeps fp16 weights on CUDA).

  python tools/finetune_bench.py            # writes gpurun_out/finetune_bench.json
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "shims"))
from arp_b200.finetune import CLIPMultiscaleAdapter  # noqa: E402
from arp_b200.weights import random_clip_state_dict  # noqa: E402

dev = torch.device("cuda", 0)
B, SIZE = 64, 64
rng = np.random.default_rng(0)
clip_sd = random_clip_state_dict("ViT-B/16", 0, "cpu")
out = {}


def events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def batch_of(goal):
    b = {f"image{i}": {"ob": torch.from_numpy(rng.integers(0, 256, size=(B, SIZE, SIZE, 3), dtype=np.uint8)).to(dev)} for i in range(4)}
    tok = torch.zeros(B, 77, dtype=torch.long)
    tok[:, 0], tok[:, 1:9], tok[:, 9] = 49406, torch.randint(1000, 40000, (1, 8)), 49407
    b["instruct"] = tok.to(dev)
    b["r"] = torch.from_numpy(rng.integers(0, 2, size=B).astype(np.float32)).to(dev)
    b["action"] = torch.from_numpy(rng.integers(0, 15, size=B).astype(np.int64)).to(dev)
    return b


for goal in (False, True):
    m = CLIPMultiscaleAdapter(clip_state_dict=clip_sd, use_discrete_action=True, action_dim=15, use_vip_loss=True,
                              use_id_loss=True, goal_conditioned=goal, init="normal", device=dev, max_batch=256)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=0.001)
    batch = batch_of(goal)
    n_img = B * (4 if goal else 3)

    def step():
        loss = m(batch)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = events()
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / 10
    # the frozen tower alone (what the native library replaces)
    imgs = m.preprocess(torch.cat([batch[f"image{i}"]["ob"] for i in range(4 if goal else 3)]), train=True)
    m.clip_model.image_taps(imgs)
    e0, e1 = events()
    e0.record()
    for _ in range(10):
        m.clip_model.image_taps(imgs)
    e1.record()
    torch.cuda.synchronize()
    tower_ms = e0.elapsed_time(e1) / 10
    # the same tower, PyTorch eager fp16 with forward hooks (the reference's way)
    import clip as clip_shim
    ref = clip_shim.build("ViT-B/16", 0)
    ref.load_state_dict(clip_sd, strict=True)
    ref = ref.to(dev).eval()
    for mod in ref.modules():                # openai/CLIP convert_weights: fp16 conv / linear / attention / projections,
        if isinstance(mod, (torch.nn.Conv2d, torch.nn.Linear)):                      # LayerNorm parameters stay fp32
            mod.half()
        if isinstance(mod, torch.nn.MultiheadAttention):
            for a in ("in_proj_weight", "in_proj_bias"):
                getattr(mod, a).data = getattr(mod, a).data.half()
    for name in ("text_projection", "positional_embedding"):
        getattr(ref, name).data = getattr(ref, name).data.half()
    for name in ("proj", "class_embedding", "positional_embedding"):
        getattr(ref.visual, name).data = getattr(ref.visual, name).data.half()
    acts = {}
    for name, mod in ref.named_modules():
        if name.startswith("visual.transformer.resblocks.") and name.split(".")[-1].isdigit():
            mod.register_forward_hook(lambda mod_, i, o, name=name: acts.__setitem__(name, o.detach()))
    with torch.no_grad():
        roles = torch.split(imgs.half(), B)
        for _ in range(3):
            for x in roles:
                ref.encode_image(x)
        torch.cuda.synchronize()
        e0, e1 = events()
        e0.record()
        for _ in range(10):
            for x in roles:                      # one encode_image call per image role (:207-215)
                ref.encode_image(x)
        e1.record()
        torch.cuda.synchronize()
    eager_ms = e0.elapsed_time(e1) / 10
    key = "finetune_step_goal" if goal else "finetune_step_text"
    out[key] = {"batch": B, "images_per_step": n_img, "step_ms": step_ms, "frozen_tower_ms": tower_ms,
                "autograd_half_ms": step_ms - tower_ms, "frozen_tower_images_per_s": 1e3 * n_img / tower_ms,
                "torch_eager_fp16_tower_ms": eager_ms, "tower_speedup_vs_torch_eager_fp16": eager_ms / tower_ms,
                "step_ms_with_torch_eager_tower": step_ms - tower_ms + eager_ms}
    print(key, out[key], flush=True)
    del m, opt, ref
    torch.cuda.empty_cache()

(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "finetune_bench.json").write_text(json.dumps(out, indent=1))
