"""The numerics model behind DESIGN.md's format choices (tools/precision_study.py: 16-bit operands, fp32 accumulation,
fp32 LayerNorm / softmax, emulated on the CPU). Pins the claims the tolerances in tests/_util.py rest on: the bf16
operand + fp32 residual pipeline sits inside TOL_COS_ABS, fp16 operands are several times closer, a bf16 residual
stream is not acceptable, and folding LayerNorm into the consumer GEMM costs next to nothing."""
import re
import subprocess
import sys
from pathlib import Path

from _util import TOL_COS_ABS

ROOT = Path(__file__).resolve().parents[1]


def test_emulated_formats_rank_as_designed():
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "precision_study.py"), "4"], capture_output=True, text=True,
                         timeout=600, check=True).stdout
    rows = {}
    for line in out.splitlines():
        m = re.match(r"(\w+)\s+(\w+)\s+(\w+)\s+(own|fold)\s+([0-9.e+-]+)\s+([0-9.e+-]+)", line)
        if m:
            rows[(m.group(1), m.group(2), m.group(3), m.group(4))] = float(m.group(5))
    today = rows[("bf16", "fp32", "bf16", "own")]
    assert today <= TOL_COS_ABS                                        # the product's format, inside the stated bar
    assert rows[("fp16", "fp32", "fp16", "own")] <= today / 4          # fp16 operands: measured 10x on the GPU
    assert rows[("fp16", "fp32", "bf16", "own")] <= today / 4          # ... and P may stay bf16
    assert rows[("bf16", "bf16", "bf16", "own")] > TOL_COS_ABS         # why the residual stream is not bf16
    assert rows[("fp16", "fp16", "fp16", "fold")] <= today             # the round-2 plan is no worse than today
    assert rows[("bf16", "fp32", "bf16", "fold")] <= 1.5 * today       # folding LayerNorm costs next to nothing
