"""Consumer side of the labels — SURVEY.md §8(f)2, reference ProcgenDataset.preprocess_rtgs
(arp_dt/data_procgen.py:132-174) and compute_scale (arp_dt/utils.py:453-463).

Fixtures under tests/golden/rtgs/ come from the reference's UNMODIFIED method (oracle/make_golden_rtgs.py).
Everything here is order/byte-exact work, so every comparison is bit-exact."""
import json

import numpy as np
import pytest
import torch

from _util import GOLDEN

RTGS = GOLDEN / "rtgs"
NAMES = sorted(p.stem for p in RTGS.glob("r*.npz"))


def load(name):
    z = np.load(RTGS / f"{name}.npz", allow_pickle=False)
    return json.loads(str(z["meta"])), {k: z[k] for k in z.files if k != "meta"}


def inputs(spec):
    """Same recipe as oracle/make_golden_rtgs.py::inputs."""
    rng = np.random.default_rng(spec["seed"])
    lens = rng.integers(spec["len_lo"], spec["len_hi"] + 1, size=spec["n_eps"])
    traj_idx = [0] + [int(x) for x in np.cumsum(lens)]
    T = traj_idx[-1]
    rewards = {k: (spec["mu"] + spec["sd"] * rng.standard_normal(T)).astype(np.float32) for k in spec["keys"].split(", ")}
    return traj_idx, rewards


def check(meta, gold, rtgs, rmin, rmax, rtg, scale):
    for k in meta["keys"].split(", "):
        assert rtgs[k].dtype == np.float32 and np.array_equal(rtgs[k], gold[f"rtgs_{k}"]), k
        assert np.array_equal(np.float32(rmin[k]), gold[f"min_{k}"]) and np.array_equal(np.float32(rmax[k]), gold[f"max_{k}"])
    assert float(rtg) == float(gold["return_to_go"]) and int(scale) == int(gold["scale"])


def test_fixtures_exist():
    assert len(NAMES) == 5


@pytest.mark.parametrize("name", NAMES)
def test_oracle_port_reproduces_reference_preprocess_rtgs(name):
    from oracle import port
    meta, gold = load(name)
    traj_idx, rewards = inputs(meta)
    out = port.preprocess_rtgs(rewards, traj_idx, meta["F"], meta["env_name"], meta["norm"])
    check(meta, gold, out["rtgs"], out["reward_min"], out["reward_max"], out["return_to_go"], out["scale"])


def test_compute_scale_table():
    from arp_b200.rtg_dataset import compute_scale
    from oracle import port
    z = np.load(RTGS / "compute_scale_table.npz")
    for x, s in zip(z["x"], z["scale"]):
        assert compute_scale(float(x)) == int(s) == port.compute_scale(float(x)), x


def test_quantile_plan_and_lerp_match_numpy():
    """The host half of the quantile (rank selection + interpolation) equals np.quantile given exact order statistics."""
    from arp_b200.rtg_dataset import _numpy_index_dtype, lerp_like_numpy, linear_quantile_plan
    rng = np.random.default_rng(0)
    for n in [1, 2, 3, 10, 11, 101, 6496, 100_003, 1_000_000]:
        x = (rng.standard_normal(n) * 50).astype(np.float32)
        s = np.sort(x)
        for q in (0.9, 0.5, 0.0, 1.0, 0.123):
            k_lo, k_hi, g = linear_quantile_plan(n, q, _numpy_index_dtype(np.float32))
            assert lerp_like_numpy(s[k_lo], s[k_hi], g) == np.quantile(x, q), (n, q)       # the INSTALLED numpy's rule
            # numpy 1.x (the reference pins 1.23.5): virtual index, gamma and the lerp in float64 — restated from
            # numpy/lib/function_base.py @ v1.23.5 (_compute_virtual_index, _get_gamma, _lerp), result cast to float32
            v = (n - 1) * q
            lo = min(max(int(np.floor(v)), 0), n - 1)
            hi = min(lo + 1, n - 1)
            gam = np.float64(v - np.floor(v)) if 0 <= v < n - 1 else np.float64(0.0)
            a, b = np.float64(s[lo]), np.float64(s[hi])
            want = a + (b - a) * gam if gam < 0.5 else b - (b - a) * (1 - gam)
            k_lo, k_hi, g = linear_quantile_plan(n, q, np.float64)
            assert (k_lo, k_hi) == (lo, hi) or s[k_lo] == s[lo]
            assert lerp_like_numpy(s[k_lo], s[k_hi], g) == np.float32(want), (n, q)


def test_reward_key_alias():
    from arp_b200.rtg_dataset import reward_key
    assert reward_key({"ob_clip_reward": 1}, "ob", "clip") == "ob_clip_reward"                 # what the labeler writes
    assert reward_key({"ob_clip_reward": 1, "ob_clip_pos_reward": 2}, "ob", "clip") == "ob_clip_pos_reward"
    with pytest.raises(KeyError):
        reward_key({}, "ob", "clip")


# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def eng():
    from arp_b200 import capi
    from arp_b200.build import build
    build()
    e = capi.Engine(device=0, max_batch=1)
    yield e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_preprocess_rtgs_matches_reference(eng, name):
    from arp_b200.rtg_dataset import preprocess_rtgs
    meta, gold = load(name)
    traj_idx, rewards = inputs(meta)
    info = preprocess_rtgs(rewards, traj_idx, meta["F"], meta["env_name"], meta["norm"], engine=eng)
    check(meta, gold, info.rtgs, info.reward_min, info.reward_max, info.return_to_go, info.scale)


@pytest.mark.gpu
def test_order_statistics_exact(eng):
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(1_000_003) * 30).astype(np.float32)
    x[:1000] = 7.25                                    # duplicates
    x[1000:1010] = [0.0, -0.0, np.inf, -np.inf, 1e-40, -1e-40, 3e38, -3e38, 1.0, -1.0]     # signed zeros, subnormals, extremes
    s = np.sort(x)
    t = torch.from_numpy(x)
    for k_lo, k_hi in [(0, 1), (len(x) - 2, len(x) - 1), (900_001, 900_002), (500_000, 500_000), (123, 999_999)]:
        got = eng.order_statistics(t, k_lo, k_hi)
        assert got[0] == s[k_lo] and got[1] == s[k_hi], (k_lo, k_hi, got, s[k_lo], s[k_hi])


@pytest.mark.gpu
def test_preprocess_rtgs_full_size_10k_episodes(eng):
    """BASELINE configs[4] size (10k episodes, ~1.4M frames, 11.5M stacked values) against numpy, bit-exact."""
    from arp_b200.rtg_dataset import preprocess_rtgs
    from oracle import cport
    rng = np.random.default_rng(3)
    lens = rng.integers(32, 256, size=10_000)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    r = (20 + 5 * rng.standard_normal(off[-1])).astype(np.float32)
    info = preprocess_rtgs({"ob": r}, off.tolist(), 8, "maze", True, engine=eng)
    shifted = r - r.min()
    ref = np.empty((len(r), 8), np.float32)
    for lo, hi in zip(off[:-1], off[1:]):
        ref[lo:hi] = cport.stack_outputs(cport.discount_cumsum(shifted[lo:hi]), 8)
    assert np.array_equal(info.rtgs["ob"], ref)
    assert info.reward_min["ob"] == r.min() and info.reward_max["ob"] == r.max()
    q = np.quantile(ref, 0.9)
    assert info.return_to_go == q // 100 * 100
