"""Consumer side of the labels: drop-in for `ProcgenDataset.preprocess_rtgs` (arp_dt/data_procgen.py:132-174) and
`compute_scale` (arp_dt/utils.py:453-463) — SURVEY.md §8(f)2.

The DT data loader re-derives, from the labeler's per-frame rewards, the stacked return-to-go tokens it feeds the
policy and the statistic that sets the conditioning token at evaluation time:

    reward_min / reward_max            (:147-148)
    reward - reward_min if use_normalize   (:150-153)
    per-episode discount_cumsum(gamma=1) + deque window of num_frames   (:155-168)
    return_to_go = max(...) // 100 * 100 for CoinRun, else quantile(..., 0.9) // 100 * 100   (:171-174)
    scale = compute_scale(return_to_go)

The reference does this in Python loops over every frame at every start-up; here the arrays go through the native
library once (arp_preprocess_rtgs: min/max, shift, scan + stack; arp_quantile_f32: exact radix select of the two
order statistics np.quantile interpolates between). Results are bit-identical to the reference's numpy.

Q2: the loader reads "<key>_<vl_type>_pos_reward" but the labeler writes "<key>_<vl_type>_reward";
`reward_key()` accepts either.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import capi


def compute_scale(return_to_go):
    """arp_dt/utils.py:453-463: 10**(digits-1) if the leading digit < 5 else 10**digits (sign skipped)."""
    s = str(int(return_to_go))
    neg = not (return_to_go >= 0)
    max_digit = int(s[1] if neg else s[0])
    digits = len(s) - (1 if neg else 0)
    n = digits - 1 if max_digit < 5 else digits
    return pow(10, n)


def reward_key(store, image_key: str, vl_type: str) -> str:
    """data_procgen.py:143 reads `_pos_reward`; label_reward.py:257 writes `_reward`."""
    for suffix in ("pos_reward", "reward"):
        k = f"{image_key}_{vl_type}_{suffix}"
        if store.get(k) is not None:
            return k
    raise KeyError(f"no reward dataset for {image_key!r} / {vl_type!r}")


def _numpy_index_dtype(dtype=np.float32):
    """The float type np.quantile computes its virtual index / gamma / lerp in for a `dtype` sample: numpy >= 2 casts q to
    the sample's dtype (float32 here); numpy 1.x — the reference pins numpy==1.23.5 — keeps q a Python float, so the
    index, gamma and the interpolation are float64 and the float32 result is the rounding of that. Parity is with the
    numpy of the environment the loader runs in; `numpy_semantics` ("1" / "2") overrides the detection."""
    return dtype if int(np.__version__.split(".")[0]) >= 2 else np.float64


def linear_quantile_plan(n: int, q: float, dtype=np.float32):
    """Where np.quantile(x, q) (method 'linear') looks in the sorted sample: (k_lo, k_hi, gamma), with numpy's own
    expressions in the index dtype `dtype` (see _numpy_index_dtype)."""
    quant = np.asanyarray(q, dtype=dtype)
    virtual = np.asanyarray((n - 1) * quant)
    prev = np.floor(virtual)
    nxt = prev + 1
    if virtual >= n - 1:
        prev = nxt = np.asanyarray(n - 1, dtype=virtual.dtype)
    if virtual < 0:
        prev = nxt = np.asanyarray(0, dtype=virtual.dtype)
    gamma = np.asanyarray(virtual - prev, dtype=virtual.dtype)
    return int(prev), int(nxt), gamma


def lerp_like_numpy(a, b, t):
    """numpy's _lerp for scalars: a + (b-a)*t, or b - (b-a)*(1-t) when t >= 0.5 — in t's dtype (float32 under numpy >= 2,
    float64 under numpy 1.x where the float32 order statistics are promoted by the float64 gamma), rounded to float32."""
    wt = np.asarray(t).dtype.type
    a, b = wt(np.float32(a)), wt(np.float32(b))
    d = np.subtract(b, a)
    out = np.add(a, d * t)
    if t >= 0.5:
        out = np.subtract(b, d * (1 - t), dtype=wt, casting="unsafe")
    return np.float32(out)


@dataclass
class RtgInfo:
    rtgs: dict            # image_key -> float32 [rows, num_frames]   (data_procgen.py's self.rtgs)
    reward_min: dict
    reward_max: dict
    return_to_go: float
    scale: int


def preprocess_rtgs(rewards: dict, traj_idx, num_frames: int, env_name: str, use_normalize: bool,
                    engine: "capi.Engine | None" = None, device: int = 0, numpy_semantics: "str | None" = None) -> RtgInfo:
    """rewards: image_key -> per-frame reward (float32 [T], i.e. column -1 of the reward dataset).
    traj_idx: the loader's h5_file_traj_idx ([0] + done indices + 1, data_procgen.py:118-121)."""
    own = engine is None
    if own:
        engine = capi.Engine(device=device, max_batch=1)
    try:
        off = torch.as_tensor(np.asarray(traj_idx, dtype=np.int64))
        stacks, rmin, rmax, gmax = {}, {}, {}, {}
        for key, r in rewards.items():
            r = np.ascontiguousarray(r, dtype=np.float32)
            gs, _, stats = engine.preprocess_rtgs(torch.from_numpy(r), off, num_frames, use_normalize)
            rows = min(int(off[-1]), r.shape[0])
            stacks[key] = gs[:rows]
            rmin[key], rmax[key], gmax[key] = stats[0], stats[1], stats[3]
        if "coinrun" in env_name:
            rtg = max(gmax.values())                                            # np.max over every key's stack (:171)
        else:
            allv = torch.cat([v.reshape(-1) for v in stacks.values()]) if len(stacks) > 1 else next(iter(stacks.values()))
            n = allv.numel()
            idx_dtype = {None: _numpy_index_dtype(np.float32), "1": np.float64, "2": np.float32}[numpy_semantics]
            k_lo, k_hi, gamma = linear_quantile_plan(n, 0.9, idx_dtype)
            lo, hi = engine.order_statistics(allv, k_lo, k_hi)
            rtg = lerp_like_numpy(lo, hi, gamma)                                # np.quantile(..., 0.9) (:173)
        return_to_go = rtg // 100 * 100
        return RtgInfo({k: v.cpu().numpy() for k, v in stacks.items()}, rmin, rmax, return_to_go,
                       compute_scale(return_to_go))
    finally:
        if own:
            engine.close()


def preprocess_rtgs_from_store(store, image_keys: str, vl_type: str, num_frames: int, env_name: str,
                               use_normalize: bool, **kw) -> RtgInfo:
    """The loader's call shape: reads the reward datasets and the episode index from an open store."""
    from .label_reward import episode_index
    _, _, traj_idx = episode_index(store, "done")
    rewards = {k: np.asarray(store[reward_key(store, k, vl_type)][:, -1]).astype(np.float32)
               for k in image_keys.split(", ")}
    return preprocess_rtgs(rewards, traj_idx, num_frames, env_name, use_normalize, **kw)
