"""Shared helpers for the parity tests: load golden fixtures and rebuild their seeded inputs/weights."""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT / "oracle" / "shims"))

# tolerances of the bf16 tensor-core pipeline against the fp32 reference, stated once:
#   BASELINE.json north_star: rewards within 1e-3 relative (bf16). With RANDOM-INIT weights the
#   image/text cosine is ~1e-2, so a relative bound on the reward is ill-conditioned;
#   the well-posed statement is on the cosine itself:  |Δcos| = |Δreward| / exp(logit_scale).
TOL_COS_ABS = 5e-4          # absolute error of the cosine similarity, bf16 pipeline
TOL_REWARD_REL_TO_MAX = 2e-2   # max|Δr| / max|r_ref| at random init (reported, loose)
TOL_REWARD_REL_CORRELATED = 1e-3  # north_star's 1e-3 relative, on the well-conditioned correlated-text case
TOL_REWARD_REL_OPERATING_POINT = 1e-3   # north_star's 1e-3, per frame, at cos 0.15-0.35 and scale 100 (default build)
LOGIT_SCALE_RANDOM_INIT = 1.0 / 0.07
# fp32 verification path (precision="fp32") against the fp32 reference — north_star's 1e-5:
TOL_F32_REL = 1e-5          # max|Δr| / max|r_ref| per golden, and per-frame relative on the correlated-text case
TOL_F32_COS_ABS = 1e-6      # |Δcos|


def golden_names():
    return sorted(p.stem for p in GOLDEN.glob("*.npz"))


def load_golden(name: str):
    z = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    arrays = {k: z[k] for k in z.files if k != "meta"}
    return meta, arrays


def digest(arrs) -> str:
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def state_digest(sd: dict) -> str:
    keys = sorted(sd)
    return digest([sd[k].detach().cpu().float().numpy() for k in keys[:: max(1, len(keys) // 24)]])


def rebuild_inputs(meta: dict):
    """(dataset dict, clip state_dict, adapter state_dict | None) exactly as oracle/make_golden.py built them."""
    import clip as clip_shim  # oracle shim
    from arp_b200.synth import make_dataset
    from arp_b200.weights import random_adapter_state_dict
    data = make_dataset(**meta["data"])
    assert digest([data["ob"][:, -1]]) == meta["frames_digest"], "synthetic frames drifted from the golden recipe"
    clip_sd = {k: v.clone() for k, v in clip_shim.build(meta["arch"], meta["seed"]).state_dict().items()}
    adapter_sd = None
    if meta["model_type"].startswith("clip_") and meta["model_type"] != "clip_goal_conditioned":
        adapter_sd = random_adapter_state_dict(meta["arch"], seed=meta["adapter_seed"], device="cpu", clip_sd=clip_sd)
        assert state_digest(adapter_sd) == meta["weights_digest"], "adapter weights drifted from the golden recipe"
    else:
        assert state_digest(clip_sd) == meta["weights_digest"], "CLIP shim weights drifted from the golden recipe"
    return data, clip_sd, adapter_sd
