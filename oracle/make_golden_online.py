"""ORACLE / TEST INFRASTRUCTURE — not part of the shipped product.

Golden vectors for the ONLINE reward: executes the REFERENCE's unmodified
`arp_dt/envs/vl_reward.py` (get_torch_clip_reward, get_torch_clip_goal_conditioned_reward,
get_torch_clip_adapter_reward, get_torch_clip_adapter_goal_conditioned_reward) on seeded frames, on top of
the same shims / stubs as oracle/run_reference.py, and writes tests/golden/online/*.npz.
Build container only (needs /root/reference):   python -m oracle.make_golden_online
"""
from __future__ import annotations

import importlib
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from arp_b200.synth import structured_frames  # noqa: E402
from arp_b200.weights import random_adapter_state_dict  # noqa: E402
from oracle import stubs  # noqa: E402
from oracle.make_golden import digest, state_digest  # noqa: E402
from oracle.run_reference import reference_adapter, reference_clip_state_dict  # noqa: E402

OUT = ROOT / "tests" / "golden" / "online"
COIN = "the goal is to collect the coin."
MAZE = ["The agent must navigate a maze to find the line.", "The agent navigate to the yellow object."]

CASES = {
    # vl_type, arch, frame size, text (str -> row 0, list -> mean over texts: vl_reward.py:19-22), crop
    "o1_clip_b32_64": dict(vl_type="clip", arch="ViT-B/32", seed=0, size=64, text=COIN, use_crop=False, n=6, frames_seed=10),
    "o2_clip_b16_256_list_crop": dict(vl_type="clip", arch="ViT-B/16", seed=0, size=256, text=MAZE, use_crop=True, n=3, frames_seed=11),
    "o3_clip_goal_b32_64": dict(vl_type="clip_goal_conditioned", arch="ViT-B/32", seed=1, size=64, text=None, use_crop=False, n=5, frames_seed=12),
    "o4_clip_goal_b32_crop": dict(vl_type="clip_goal_conditioned", arch="ViT-B/32", seed=1, size=64, text=None, use_crop=True, n=4, frames_seed=13),
    "o5_clipft_b16_64_list": dict(vl_type="clip_ft", arch="ViT-B/16", seed=0, adapter_seed=1, size=64, text=MAZE, use_crop=False, n=3, frames_seed=14),
    "o6_clipft_goal_b16_64": dict(vl_type="clip_ft_goal_conditioned", arch="ViT-B/16", seed=0, adapter_seed=1, size=64, text=None, use_crop=False, n=3, frames_seed=15),
}


def make_case(name: str):
    spec = dict(CASES[name])
    stubs.import_reference()
    import clip as clip_shim
    vl = importlib.import_module("arp_dt.envs.vl_reward")          # the reference's own file, unmodified
    clip_shim.OVERRIDE_ARCH, clip_shim.SEED = spec["arch"], spec["seed"]
    rng = np.random.default_rng(spec["frames_seed"])
    frames = structured_frames(spec["n"] + 1, spec["size"], rng)   # last frame = goal image
    obs, goal = frames[:-1], frames[-1]
    model, preprocess = clip_shim.load("ViT-B/16", device="cpu")   # main_procgen.py:570 (the shim honours OVERRIDE_ARCH)
    if spec["vl_type"].startswith("clip_ft"):
        clip_sd = reference_clip_state_dict(spec["arch"], spec["seed"])
        sd = random_adapter_state_dict(spec["arch"], seed=spec["adapter_seed"], device="cpu", clip_sd=clip_sd)
        orig = torch.nn.init.orthogonal_
        torch.nn.init.orthogonal_ = lambda w, gain=1.0: w         # the checkpoint overwrites every adapter tensor
        try:
            model = reference_adapter(spec["arch"], spec["seed"], spec["adapter_seed"])
        finally:
            torch.nn.init.orthogonal_ = orig
        model.load_state_dict(sd, strict=False)                    # main_procgen.py:585-587
        model.eval()
        wdig = state_digest(sd)
    else:
        wdig = state_digest(reference_clip_state_dict(spec["arch"], spec["seed"]))
    cm = (model, preprocess)
    fn = {"clip": vl.get_torch_clip_reward, "clip_goal_conditioned": vl.get_torch_clip_goal_conditioned_reward,
          "clip_ft": vl.get_torch_clip_adapter_reward,
          "clip_ft_goal_conditioned": vl.get_torch_clip_adapter_goal_conditioned_reward}[spec["vl_type"]]
    second = goal if "goal" in spec["vl_type"] else spec["text"]
    out = [np.asarray(fn(cm, o, second, use_crop=spec["use_crop"])) for o in obs]
    rewards = np.stack(out)
    meta = dict(spec, name=name, weights_digest=wdig, frames_digest=digest([frames]), torch=torch.__version__,
                generator="oracle/make_golden_online.py -> /root/reference/arp_dt/envs/vl_reward.py (unmodified)")
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / f"{name}.npz", meta=json.dumps(meta), reward=rewards)
    print(f"{name}: reward {rewards.shape} {rewards.dtype} {rewards.reshape(-1)[:3]}")


if __name__ == "__main__":
    if not stubs.reference_available():
        sys.exit("needs /root/reference (build container only)")
    for n in (sys.argv[1:] or list(CASES)):
        make_case(n)
