"""Dev tool: correctness + timing of the tcgen05 / TMEM attention kernel."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from arp_b200 import capi  # noqa: E402

dev = torch.device("cuda", 0)
impls = [2]


def ref_attn(qkv, B, L):
    q, k, v = qkv.float().view(B, L, 3, 12, 64).permute(2, 0, 3, 1, 4)
    return (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, 768)


for impl in impls:
    eng = capi.Engine(device=0, max_batch=8)
    print(f"==== attention impl {impl}", flush=True)
    for B, L, scale in ((1, 197, 1.5), (5, 197, 1.5), (7, 50, 1.5), (3, 197, 4.0), (40, 197, 0.5)):
        g = torch.Generator(device=dev).manual_seed(B * L)
        qkv = (torch.randn(B * L, 2304, device=dev, generator=g) * scale).bfloat16()
        try:
            out = eng.attention(qkv, B, L)
            torch.cuda.synchronize()
            ref = ref_attn(qkv, B, L)
            err = float((out.float() - ref).abs().max() / ref.abs().max())
            ok = err < 2e-2 and bool(torch.isfinite(out.float()).all())
            per_frame = [(out.float() - ref).view(B, L, 768)[b].abs().max().item() for b in range(min(B, 3))]
        except Exception as e:  # noqa: BLE001
            err, ok, per_frame = str(e), False, []
        print(f"  {'ok ' if ok else 'BAD'} B={B} L={L} scale={scale}: relerr {err} {per_frame}", flush=True)
    B, L = 512, 197
    qkv = (torch.randn(B * L, 2304, device=dev) * 1.5).to(capi.operand_dtype())   # no conversion inside the timed loop
    for _ in range(3):
        eng.attention(qkv, B, L)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        eng.attention(qkv, B, L)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    print(f"  B=512 L=197: {ms * 1e3:.1f} us  ({4 * B * 12 * L * L * 64 / ms / 1e9:.1f} TFLOP/s algorithmic)", flush=True)
    eng.close()
