"""Seeded synthetic Procgen-shaped demonstration datasets.

Layout follows data/PPG/trajectory_recorder.py:134-189: every dataset has first dim = total steps,
second dim = num_frames sliding window left-padded with the episode's first element; `done[:, -1]`
is 1.0 only at an episode's last step. Used by tests, bench.py and the examples — there is no
network for the real demonstrations.
"""
from __future__ import annotations

import numpy as np


def episode_lengths(n_episodes: int, lo: int, hi: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(lo, hi, size=n_episodes)


def stack_frames(x: np.ndarray, num_frames: int) -> np.ndarray:
    """trajectory_recorder.py:103-115 semantics: out[i, f] = x[max(0, i - (F-1-f))]."""
    T = x.shape[0]
    idx = np.maximum(0, np.arange(T)[:, None] - (num_frames - 1 - np.arange(num_frames))[None, :])
    return x[idx]


def structured_frames(n: int, size: int, rng: np.random.Generator) -> np.ndarray:
    """Low-frequency background + a few solid sprites: exercises the resize on non-white-noise content."""
    base = rng.integers(0, 256, size=(n, 4, 4, 3)).astype(np.float32)
    rep = size // 4
    img = np.repeat(np.repeat(base, rep, axis=1), rep, axis=2)
    ramp = np.linspace(0, 40, size, dtype=np.float32)
    img = img + ramp[None, :, None, None] - ramp[None, None, :, None]
    img = np.clip(img, 0, 255).astype(np.uint8)
    for i in range(n):
        for _ in range(3):
            y, x = rng.integers(0, size - size // 8, size=2)
            img[i, y:y + size // 8, x:x + size // 8] = rng.integers(0, 256, size=3, dtype=np.uint8)
    return img


def make_dataset(n_episodes: int = 16, len_lo: int = 20, len_hi: int = 60, size: int = 64, num_frames: int = 4,
                 seed: int = 0, kind: str = "noise", stacked: bool = True, tail_rows: int = 0) -> dict:
    """Returns {"ob", "done", "reward", "act", "lengths"}.

    stacked=True  : ob is [T,F,H,W,3] with real frame stacking (reference layout).
    stacked=False : ob is [T,1,H,W,3] (only the scored frame; for large synthetic sets).
    tail_rows     : extra rows after the last `done` (the reference never labels them, label_reward.py:82-83).
    """
    rng = np.random.default_rng(seed)
    lengths = episode_lengths(n_episodes, len_lo, len_hi, seed)
    T = int(lengths.sum()) + tail_rows
    if kind == "noise":
        frames = rng.integers(0, 256, size=(T, size, size, 3), dtype=np.uint8)
    else:
        frames = structured_frames(T, size, rng)
    done = np.zeros(T, np.float32)
    done[np.cumsum(lengths) - 1] = 1.0
    reward = np.zeros(T, np.float32)
    act = rng.integers(0, 15, size=T).astype(np.int64)
    F = num_frames
    ob = np.empty((T, F, size, size, 3), np.uint8) if stacked else frames[:, None]
    done_s = np.empty((T, F), np.float32)
    rew_s = np.empty((T, F), np.float32)
    act_s = np.empty((T, F), np.int64)
    bounds = np.concatenate([[0], np.cumsum(lengths), [T]]) if tail_rows else np.concatenate([[0], np.cumsum(lengths)])
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        if hi <= lo:
            continue
        if stacked:
            ob[lo:hi] = stack_frames(frames[lo:hi], F)
        done_s[lo:hi] = stack_frames(done[lo:hi], F)
        rew_s[lo:hi] = stack_frames(reward[lo:hi], F)
        act_s[lo:hi] = stack_frames(act[lo:hi], F)
    return {"ob": ob, "done": done_s, "reward": rew_s, "act": act_s, "lengths": lengths}


def write_dataset(store, data: dict, image_key: str = "ob"):
    for k in ("done", "reward", "act"):
        store.create_dataset(k, data=data[k])
    store.create_dataset(image_key, data=data["ob"])
    return store
