"""ORACLE / TEST INFRASTRUCTURE — not part of the shipped product.

Runs the REFERENCE's own, unmodified `arp_dt/label_reward.py` (and, for clip_ft,
`finetune_module/clip_multiscale_adapter.py`) from /root/reference on top of the shims in
oracle/shims (clip = restated openai/CLIP with seeded random weights; h5py = in-memory store) and the
stubs in oracle/stubs.py. Only possible where /root/reference is mounted (the build container);
its outputs are committed as golden fixtures by oracle/make_golden.py.
"""
from __future__ import annotations

import sys
import time
from functools import partial
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import stubs  # noqa: E402


def run_reference_label_reward(data: dict, *, arch: str = "ViT-B/16", seed: int = 0, model_type: str = "clip",
                               text="the goal is to collect the coin.", use_crop: bool = False,
                               inst_type: str = "none", model_ckpt_dir=None, image_key: str = "ob",
                               threads: int | None = None, store_path: str = "mem://oracle") -> dict:
    """data: dict from arp_b200.synth.make_dataset. Returns the datasets the reference wrote + timing."""
    lr = stubs.import_reference()
    import clip as clip_shim
    import h5py as h5_shim

    clip_shim.OVERRIDE_ARCH = arch      # the reference hard-codes "ViT-B/16" (label_reward.py:126)
    clip_shim.SEED = seed
    if threads:
        torch.set_num_threads(threads)
    if model_type.startswith("clip_") and model_type != "clip_goal_conditioned":
        # pinned torchvision 0.12 resizes tensors WITHOUT antialias; this container's 0.26 defaults to
        # antialias=True for tensors. Restore the pinned behaviour (SURVEY.md §8 a3').
        import finetune_module.clip_multiscale_adapter as cma
        from torchvision.transforms.functional import resize as tv_resize
        cma.resize = partial(tv_resize, antialias=False)
    h5_shim.reset()
    f = h5_shim.File(store_path, "w")
    for k in ("done", "reward", "act"):
        f.create_dataset(k, data=data[k])
    f.create_dataset(image_key, data=data["ob"])
    t0 = time.perf_counter()
    lr.label_reward(env_name="coinrun", distribution_mode="hard", num_levels=500, start_level=0, text=text,
                    base_path=".", data_path=store_path, image_keys=image_key, model_type=model_type,
                    model_ckpt_dir=model_ckpt_dir, use_crop=use_crop, inst_type=inst_type, env_type="none")
    secs = time.perf_counter() - t0
    g = h5_shim.File(store_path, "a")
    out = {k: np.array(g[k][:]) for k in g.keys() if k.startswith(f"{image_key}_")}
    out["_seconds"] = secs
    return out


def reference_clip_state_dict(arch: str, seed: int) -> dict:
    stubs.install()
    import clip as clip_shim
    return {k: v.clone() for k, v in clip_shim.build(arch, seed).state_dict().items()}


def reference_adapter(arch: str = "ViT-B/16", seed: int = 0, adapter_seed: int = 1, fast_init: bool = True):
    """Instantiate the reference's CLIPMultiscaleAdapter (random init). fast_init replaces the O(n^3)
    orthogonal init of the 13312x6656 matrices with a scaled gaussian (values are test inputs, not behaviour)."""
    stubs.import_reference()
    import clip as clip_shim
    clip_shim.OVERRIDE_ARCH, clip_shim.SEED = arch, seed
    import finetune_module.clip_multiscale_adapter as cma
    from torchvision.transforms.functional import resize as tv_resize
    cma.resize = partial(tv_resize, antialias=False)
    torch.manual_seed(adapter_seed)
    if fast_init:
        orig = torch.nn.init.orthogonal_
        torch.nn.init.orthogonal_ = lambda w, gain=1.0: torch.nn.init.normal_(w, std=gain / (w.shape[1] ** 0.5))
        try:
            model = cma.CLIPMultiscaleAdapter(device=torch.device("cpu"), use_discrete_action=True, action_dim=15)
        finally:
            torch.nn.init.orthogonal_ = orig
    else:
        model = cma.CLIPMultiscaleAdapter(device=torch.device("cpu"), use_discrete_action=True, action_dim=15)
    return model.eval()
