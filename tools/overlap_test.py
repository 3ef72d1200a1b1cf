"""Dev tool: do LayerNorm blocks co-run with a persistent GEMM on another stream?"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from arp_b200 import capi  # noqa: E402
dev = torch.device("cuda", 0)
eng = capi.Engine(device=0, max_batch=8)
M = 197 * 512
a = (torch.randn(M, 768, device=dev) * 0.5).bfloat16()
w = (torch.randn(3072, 768, device=dev) * 0.05).bfloat16()
b = torch.randn(3072, device=dev)
x = torch.randn(M, 768, device=dev)
g, be = torch.ones(768, device=dev), torch.zeros(768, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def gemm():
    eng.gemm(a, w, bias=b, act=1)


def ln():
    eng.layernorm(x, g, be)


def both():
    ev = torch.cuda.Event(); ev.record()
    with torch.cuda.stream(s1):
        s1.wait_event(ev); gemm()
    with torch.cuda.stream(s2):
        s2.wait_event(ev); ln(); ln()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)


tg, tl = t(gemm), t(ln)
tb = t(both)
print(f"gemm {tg:.0f} us, ln {tl:.0f} us, gemm || 2x ln {tb:.0f} us (serial would be {tg + 2 * tl:.0f})")
