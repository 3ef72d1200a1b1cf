"""Dev tool: the fp32 verification path against every golden fixture and the correlated-text case."""
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from _util import LOGIT_SCALE_RANDOM_INIT, golden_names, load_golden, rebuild_inputs  # noqa: E402
from test_gpu_parity import _run_product  # noqa: E402

for name in golden_names():
    meta, gold = load_golden(name)
    data, clip_sd, adapter_sd = rebuild_inputs(meta)
    with tempfile.TemporaryDirectory() as d:
        out = _run_product(Path(d), meta, data, clip_sd, adapter_sd, precision="fp32", max_batch=16)
    rk = [k for k in gold if "_pos_rtg" not in k][0]
    gk = [k for k in gold if "_pos_rtg" in k][0]
    r, r_ref = out[rk][:, -1].astype(np.float64), gold[rk][:, -1].astype(np.float64)
    print(f"{name:32s} max|dr| {np.abs(r - r_ref).max():.3e}  |dcos| {np.abs(r - r_ref).max() / LOGIT_SCALE_RANDOM_INIT:.3e}"
          f"  max|r| {np.abs(r_ref).max():.3e}  rel-to-max {np.abs(r - r_ref).max() / np.abs(r_ref).max():.3e}"
          f"  rtg rel-to-max {np.abs(out[gk] - gold[gk]).max() / np.abs(gold[gk]).max():.3e}", flush=True)
