"""Measurement for the widened rows: per-call latency of the online reward and time of the
consumer-side statistics, each next to the reference algorithm (oracle port) on the host cores.

  python tools/next_rows_bench.py            # writes gpurun_out/next_rows.json
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import os as _os
_os.environ.setdefault("ARP_ALLOW_STANDIN_TOKENIZER", "1")   # random-init weights: the deterministic stand-in token ids
from arp_b200 import online  # noqa: E402
from arp_b200.rtg_dataset import preprocess_rtgs  # noqa: E402
from arp_b200.weights import random_clip_state_dict  # noqa: E402
from oracle import port  # noqa: E402

out = {}
text = "the goal is to collect the coin."
sd = random_clip_state_dict("ViT-B/16", 0, "cpu")
rng = np.random.default_rng(0)

# ---- (f)1 online reward: one 64x64 / 256x256 frame per call, like a rollout step ----
for size in (64, 256):
    frames = rng.integers(0, 256, size=(64, size, size, 3), dtype=np.uint8)
    m = online.OnlineClip("clip", clip_state_dict=sd)
    for i in range(5):
        online.get_torch_clip_reward(m, frames[i], text)
    torch.cuda.synchronize()
    ts = []
    for i in range(5, 64):
        t0 = time.perf_counter()
        online.get_torch_clip_reward(m, frames[i], text)
        ts.append(time.perf_counter() - t0)
    eng = m._engines[(size, size)]
    # the same kernels launched one by one (profiling mode bypasses the CUDA graph)
    tn = []
    for i in range(5, 25):
        eng.profile_begin()
        t0 = time.perf_counter()
        eng.online_reward(frames[i])
        tn.append(time.perf_counter() - t0)
        eng.profile_end()
    m.close()
    model = port.clip_shim.build("ViT-B/16", 0)
    model.load_state_dict(sd, strict=True)
    model = model.float().eval()
    port.online_clip_reward(model, frames[0], text)
    t0 = time.perf_counter()
    for i in range(1, 4):
        port.online_clip_reward(model, frames[i], text)
    cpu = (time.perf_counter() - t0) / 3
    out[f"online_clip_b16_{size}"] = {"graph_ms_median": 1e3 * float(np.median(ts)), "graph_ms_p95": 1e3 * float(np.percentile(ts, 95)),
                                      "stream_launch_ms_median": 1e3 * float(np.median(tn)),
                                      "cpu_port_ms": 1e3 * cpu, "cpu_threads": torch.get_num_threads()}
    print(size, out[f"online_clip_b16_{size}"], flush=True)

# ---- (f)2 consumer side: 10k episodes, F = 8 ----
lens = rng.integers(32, 256, size=10_000)
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
r = (20 + 5 * rng.standard_normal(off[-1])).astype(np.float32)
from arp_b200 import capi  # noqa: E402
eng = capi.Engine(device=0, max_batch=1)
preprocess_rtgs({"ob": r}, off.tolist(), 8, "maze", True, engine=eng)
t0 = time.perf_counter()
info = preprocess_rtgs({"ob": r}, off.tolist(), 8, "maze", True, engine=eng)
gpu_s = time.perf_counter() - t0
eng.close()
n_eps_cpu = 300                                   # the reference's python loops: time a slice and scale by frames
sub = int(off[n_eps_cpu])
t0 = time.perf_counter()
port.preprocess_rtgs({"ob": r[:sub]}, off[:n_eps_cpu + 1].tolist(), 8, "maze", True)
cpu_s = (time.perf_counter() - t0) * (off[-1] / sub)
out["preprocess_rtgs_10k_eps"] = {"frames": int(off[-1]), "gpu_s_incl_h2d_d2h": gpu_s, "cpu_port_s_extrapolated": cpu_s,
                                  "return_to_go": float(info.return_to_go), "scale": int(info.scale)}
print(out["preprocess_rtgs_10k_eps"])
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "next_rows.json").write_text(json.dumps(out, indent=1))
