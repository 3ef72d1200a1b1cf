"""Online (per environment step) reward — SURVEY.md §8(f)1, reference arp_dt/envs/vl_reward.py.

Fixtures under tests/golden/online/ were produced by the reference's UNMODIFIED vl_reward.py
(oracle/make_golden_online.py). CPU tests pin the oracle port against them; GPU tests run the product
(arp_b200.online -> arp_online_reward, one CUDA graph per call) against the same fixtures."""
import json

import numpy as np
import pytest
import torch

from _util import (GOLDEN, LOGIT_SCALE_RANDOM_INIT, TOL_COS_ABS, TOL_F32_COS_ABS, digest, state_digest)

ONLINE = GOLDEN / "online"
NAMES = sorted(p.stem for p in ONLINE.glob("*.npz"))


def load(name):
    z = np.load(ONLINE / f"{name}.npz", allow_pickle=False)
    return json.loads(str(z["meta"])), z["reward"]


def rebuild(meta):
    """(obs frames, goal frame, clip state_dict, adapter state_dict | None) as make_golden_online.py built them."""
    import clip as clip_shim  # oracle shim (sys.path set by _util)
    from arp_b200.synth import structured_frames
    from arp_b200.weights import random_adapter_state_dict
    frames = structured_frames(meta["n"] + 1, meta["size"], np.random.default_rng(meta["frames_seed"]))
    assert digest([frames]) == meta["frames_digest"]
    clip_sd = {k: v.clone() for k, v in clip_shim.build(meta["arch"], meta["seed"]).state_dict().items()}
    adapter_sd = None
    if meta["vl_type"].startswith("clip_ft"):
        adapter_sd = random_adapter_state_dict(meta["arch"], seed=meta["adapter_seed"], device="cpu", clip_sd=clip_sd)
        assert state_digest(adapter_sd) == meta["weights_digest"]
    else:
        assert state_digest(clip_sd) == meta["weights_digest"]
    return frames[:-1], frames[-1], clip_sd, adapter_sd


def test_online_fixtures_exist():
    assert len(NAMES) == 6


@pytest.mark.parametrize("name", NAMES)
def test_oracle_port_reproduces_reference_online_reward(name):
    from oracle import port
    meta, gold = load(name)
    obs, goal, clip_sd, adapter_sd = rebuild(meta)
    vt = meta["vl_type"]
    if adapter_sd is not None:
        ad = port.AdapterOracle(adapter_sd, meta["arch"])
        if vt == "clip_ft":
            got = np.stack([port.online_adapter_reward(ad, o, meta["text"], meta["use_crop"]) for o in obs])
        else:
            got = np.array([port.online_goal_reward(ad.encode_image, o, goal, meta["use_crop"]) for o in obs])
    else:
        model = port.clip_shim.build(meta["arch"], meta["seed"]).float().eval()
        if vt == "clip":
            got = np.stack([port.online_clip_reward(model, o, meta["text"], meta["use_crop"]) for o in obs])
        else:
            got = np.array([port.online_goal_reward(model.encode_image, o, goal, meta["use_crop"]) for o in obs])
    assert got.shape == gold.shape and got.dtype == gold.dtype
    assert np.abs(got - gold).max() <= 2e-5 * max(1.0, np.abs(gold).max())     # same fp32 algorithm, op-order noise


def test_update_rtg_matches_rollout_formula():
    from arp_b200 import online
    from oracle import port
    r = np.float32([0.731])
    for norm in (False, True):
        a = online.update_rtg(5.0, r, 100.0, reward_min=-0.25, use_normalize=norm)
        b = port.update_rtg(5.0, r, 100.0, reward_min=-0.25, use_normalize=norm)
        ref = 5.0 - ((r - (-0.25)) / 100.0 if norm else r / 100.0)              # rollout_procgen.py:147-150
        assert np.array_equal(a, b) and np.array_equal(a, ref)


# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("name", NAMES)
def test_online_reward_matches_reference(name, precision):
    from arp_b200 import online
    from arp_b200.build import build
    build()
    meta, gold = load(name)
    obs, goal, clip_sd, adapter_sd = rebuild(meta)
    vt = meta["vl_type"]
    m = online.OnlineClip(vt, vl_checkpoint=adapter_sd, clip_state_dict=clip_sd, arch=meta["arch"], precision=precision)
    fn = {"clip": online.get_torch_clip_reward, "clip_goal_conditioned": online.get_torch_clip_goal_conditioned_reward,
          "clip_ft": online.get_torch_clip_adapter_reward,
          "clip_ft_goal_conditioned": online.get_torch_clip_adapter_goal_conditioned_reward}[vt]
    second = goal if "goal" in vt else meta["text"]
    got = np.stack([np.asarray(fn(m, o, second, use_crop=meta["use_crop"])) for o in obs])
    launches = sum(e.launch_count for e in m._engines.values())
    m.close()
    assert got.shape == gold.shape and got.dtype == gold.dtype, (got.shape, got.dtype, gold.shape, gold.dtype)
    assert launches > 0
    err = np.abs(got.astype(np.float64) - gold).max()
    if "goal" in vt:
        tol = (2e-2 if precision == "bf16" else 1e-5) * np.abs(gold).max()       # feature distances
    else:
        tol = (TOL_COS_ABS if precision == "bf16" else TOL_F32_COS_ABS) * LOGIT_SCALE_RANDOM_INIT
    assert err <= tol, f"{name} [{precision}]: max abs err {err:.3e} > {tol:.3e}"


@pytest.mark.gpu
def test_online_graph_replay_is_deterministic_and_tracks_text_changes():
    from arp_b200 import online
    from arp_b200.weights import random_clip_state_dict
    sd = random_clip_state_dict("ViT-B/16", 0, "cpu")
    m = online.OnlineClip("clip", clip_state_dict=sd)
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, 256, size=(2, 64, 64, 3), dtype=np.uint8)
    t1, t2 = "the goal is to collect the coin.", ["navigate a maze to collect the yellow cheese.", "collect the coin"]
    r1 = online.get_torch_clip_reward(m, a, t1)
    r1b = online.get_torch_clip_reward(m, a, t1)                  # graph replay: same bits
    r2 = online.get_torch_clip_reward(m, a, t2)                   # new instruction: embedding + graph rebuilt
    r3 = online.get_torch_clip_reward(m, b, t1)
    r1c = online.get_torch_clip_reward(m, a, t1)
    assert np.array_equal(r1, r1b) and np.array_equal(r1, r1c)
    assert r1.shape == (1,) and r2.shape == (1,) and not np.array_equal(r1, r2) and not np.array_equal(r1, r3)
    # batch path (arp_compute_reward) and latency path agree bit for bit on the same frame
    e = m._engines[(64, 64)]
    batch = e.compute_reward(torch.from_numpy(a[None, None]).cuda()).cpu().numpy()
    assert np.array_equal(batch, r1c)
    m.close()
