"""ORACLE / TEST INFRASTRUCTURE — not part of the shipped product.

Golden vectors for the consumer side: calls the REFERENCE's unmodified
`ProcgenDataset.preprocess_rtgs` (arp_dt/data_procgen.py:132-174) — as an unbound method on a stand-in `self`
that carries exactly the attributes the method reads (h5_file, config, h5_file_traj_idx, env_name) — and
`arp_dt.utils.compute_scale`, on seeded reward arrays. Writes tests/golden/rtgs/*.npz.
Build container only (needs /root/reference):   python -m oracle.make_golden_rtgs
"""
from __future__ import annotations

import importlib
import json
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import stubs  # noqa: E402

OUT = ROOT / "tests" / "golden" / "rtgs"

CASES = {
    # env_name picks max (coinrun) vs 0.9-quantile; scale/offset make return_to_go land in different decades
    "r1_coinrun_max": dict(env_name="coinrun", keys="ob", n_eps=40, len_lo=20, len_hi=200, F=8, seed=0, mu=24.0, sd=3.0, norm=False),
    "r2_coinrun_norm": dict(env_name="coinrun", keys="ob", n_eps=25, len_lo=5, len_hi=90, F=4, seed=1, mu=-0.3, sd=1.0, norm=True),
    "r3_maze_quantile": dict(env_name="maze", keys="ob", n_eps=60, len_lo=1, len_hi=120, F=8, seed=2, mu=21.0, sd=4.0, norm=False),
    "r4_maze_quantile_norm_2keys": dict(env_name="maze_yellowline", keys="ob, ob2", n_eps=30, len_lo=2, len_hi=60, F=4, seed=3, mu=18.0, sd=6.0, norm=True),
    "r5_bigfish_negative": dict(env_name="bigfish", keys="ob", n_eps=12, len_lo=30, len_hi=50, F=8, seed=4, mu=-20.0, sd=2.0, norm=False),
}


def inputs(spec):
    rng = np.random.default_rng(spec["seed"])
    lens = rng.integers(spec["len_lo"], spec["len_hi"] + 1, size=spec["n_eps"])
    traj_idx = [0] + [int(x) for x in np.cumsum(lens)]
    T, F = traj_idx[-1], spec["F"]
    rewards = {}
    for k in spec["keys"].split(", "):
        r = (spec["mu"] + spec["sd"] * rng.standard_normal(T)).astype(np.float32)
        rewards[k] = r
    return traj_idx, rewards, T, F


def make_case(name):
    spec = dict(CASES[name])
    stubs.import_reference()
    dp = importlib.import_module("arp_dt.data_procgen")
    ut = importlib.import_module("arp_dt.utils")
    traj_idx, rewards, T, F = inputs(spec)
    vl = "clip"
    h5 = {}
    for k, r in rewards.items():
        stacked = np.zeros((T, F), np.float32)
        stacked[:, -1] = r                                     # the method reads column -1 only (:143)
        h5[f"{k}_{vl}_pos_reward"] = stacked
    cfg = SimpleNamespace(vl_type=vl, image_key=spec["keys"], use_normalize=spec["norm"], num_frames=F, scale=None)
    me = SimpleNamespace(h5_file=h5, config=cfg, h5_file_traj_idx=traj_idx, env_name=spec["env_name"])
    rtgs = dp.ProcgenDataset.preprocess_rtgs(me)               # the reference's own code, unmodified
    out = {f"rtgs_{k}": np.asarray(v, dtype=np.float32) for k, v in rtgs.items()}
    for k in rewards:
        out[f"min_{k}"] = np.asarray(me.reward_min[k])
        out[f"max_{k}"] = np.asarray(me.reward_max[k])
    out["return_to_go"] = np.asarray(me.return_to_go)
    out["scale"] = np.asarray(me.scale)
    assert me.scale == ut.compute_scale(me.return_to_go) == cfg.scale
    meta = dict(spec, name=name, numpy=np.__version__,
                generator="oracle/make_golden_rtgs.py -> /root/reference/arp_dt/data_procgen.py ProcgenDataset.preprocess_rtgs (unmodified)")
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / f"{name}.npz", meta=json.dumps(meta), **out)
    print(f"{name}: T={T} return_to_go={me.return_to_go!r} scale={me.scale} min={me.reward_min} ")


if __name__ == "__main__":
    if not stubs.reference_available():
        sys.exit("needs /root/reference (build container only)")
    for n in (sys.argv[1:] or list(CASES)):
        make_case(n)
    # compute_scale table (utils.py:453-463), including negatives and decade edges
    stubs.import_reference()
    ut = importlib.import_module("arp_dt.utils")
    xs = [0.0, 1.0, 4.0, 5.0, 9.0, 10.0, 49.0, 50.0, 99.0, 100.0, 400.0, 500.0, 1000.0, 4999.0, 5000.0, 12345.0,
          -1.0, -4.0, -5.0, -40.0, -50.0, -100.0, -499.0, -500.0, -2000.0, -7000.0]
    np.savez_compressed(OUT / "compute_scale_table.npz", x=np.asarray(xs), scale=np.asarray([ut.compute_scale(x) for x in xs]))
    print("compute_scale table:", [(x, ut.compute_scale(x)) for x in xs][:8], "...")
