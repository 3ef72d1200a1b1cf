"""Dev tool: the two residual-update GEMMs of a block (out_proj K=768, c_proj K=3072, N=768) at the bench's chunk size,
TMA reduce-add epilogue (arp_gemm_bf16 with resid = out) against the in-register epilogue that also emits the LayerNorm
row statistics (arp_resid_gemm_stats). ARP_B200_LIB selects a variant build."""
import ctypes as C
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from arp_b200 import capi  # noqa: E402

dev = torch.device("cuda", 0)
eng = capi.Engine(device=0, max_batch=8)
op = capi.operand_dtype()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
M = B * 197
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for K in (768, 3072):
    a = (torch.randn(M, K, device=dev) * 0.5).to(op)
    w = (torch.randn(768, K, device=dev) * 0.02).to(op)
    b = torch.randn(768, device=dev) * 0.1
    x = torch.randn(M, 768, device=dev).to(op)
    stats = torch.empty(M, 2, device=dev)

    def reduce():
        eng._check(eng._lib.arp_gemm_bf16(eng._h, C.c_void_p(a.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(x.data_ptr()),
                                          capi._TORCH_DT[op], M, 768, K, C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), 0, st))

    def fused():
        eng._check(eng._lib.arp_resid_gemm_stats(eng._h, C.c_void_p(a.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()),
                                                 C.c_void_p(x.data_ptr()), C.c_void_p(stats.data_ptr()), M, 768, K, st))

    for name, fn in (("reduce-add", reduce), ("in-register + stats", fused)):
        x.normal_()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"M={M} N=768 K={K} {name:>20}: {us:8.1f} us  {2.0 * M * 768 * K / us / 1e6:7.1f} TFLOP/s", flush=True)
eng.close()
