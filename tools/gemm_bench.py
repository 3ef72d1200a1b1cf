"""Dev tool: correctness + timing of the tcgen05 GEMM on the hot shapes."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from arp_b200 import capi  # noqa: E402

dev = torch.device("cuda", 0)
impls = [3]   # one implementation ships: CTA pairs (cta_group::2)
M_HOT = 197 * int(os.environ.get('GEMM_BENCH_FRAMES', '512'))


def relerr(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max())


def ref_of(a, w, b, r, act):
    ref = a.float() @ w.float().t()
    if b is not None:
        ref = ref + b
    if act == 1:
        ref = ref * torch.sigmoid(1.702 * ref)
    elif act == 2:
        ref = torch.relu(ref)
    if r is not None:
        ref = ref + r
    return ref


CASES = [  # M, N, K, act, bias, resid, out dtype
    (128, 256, 64, 0, False, False, torch.float32),
    (256, 256, 128, 0, True, False, torch.float32),
    (197, 768, 768, 0, True, False, torch.float32),
    (1, 256, 64, 0, True, True, torch.float32),
    (300, 512, 192, 0, True, True, torch.float32),
    (197 * 40, 768, 768, 0, True, True, torch.float32),
    (197 * 40, 2304, 768, 0, True, False, torch.bfloat16),
    (197 * 40, 3072, 768, 1, True, False, torch.bfloat16),
    (197 * 40, 768, 3072, 0, True, True, torch.float32),
    (100, 13312, 6656, 2, True, False, torch.bfloat16),
]
HOT = [("qkv", 2304, 768, 0, torch.bfloat16, False), ("fc_gelu", 3072, 768, 1, torch.bfloat16, False),
       ("fc_noact", 3072, 768, 0, torch.bfloat16, False), ("fc_relu", 3072, 768, 2, torch.bfloat16, False),
       ("proj_res", 768, 3072, 0, torch.float32, True), ("out_res", 768, 768, 0, torch.float32, True)]

for impl in impls:
    eng = capi.Engine(device=0, max_batch=8)
    print(f"==== impl {impl}", flush=True)
    ok_all = True
    for (M, N, K, act, bias, resid, odt) in CASES:
        g = torch.Generator(device=dev).manual_seed(M + N + K)
        a = (torch.randn(M, K, device=dev, generator=g) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev, generator=g) * 0.05).bfloat16()
        b = torch.randn(N, device=dev, generator=g) if bias else None
        r = torch.randn(M, N, device=dev, generator=g) if resid else None
        try:
            out = eng.gemm(a, w, bias=b, resid=r, act=act, out_dtype=odt)
            torch.cuda.synchronize()
            err = relerr(out, ref_of(a, w, b, r, act))
            ok = err < (1e-2 if odt == torch.bfloat16 else 3e-5)
        except Exception as e:  # noqa: BLE001
            err, ok = str(e), False
        ok_all &= ok
        print(f"  {'ok ' if ok else 'BAD'} M={M} N={N} K={K} act={act} bias={bias} resid={resid} {odt}: {err}", flush=True)
    if not ok_all:
        print("  skipping timing (failures)")
        eng.close()
        continue
    # in-place residual (the hot-path form) and timing
    for name, N, K, act, odt, res in HOT:
        op = capi.operand_dtype()                      # raw pointers below: the library's own 16-bit format
        a = (torch.randn(M_HOT, K, device=dev) * 0.5).to(op)
        w = (torch.randn(N, K, device=dev) * 0.05).to(op)
        b = torch.randn(N, device=dev)
        odt = odt if odt == torch.float32 else op
        x = torch.randn(M_HOT, N, device=dev) if res else None
        import ctypes as C
        out = x if res else torch.empty(M_HOT, N, device=dev, dtype=odt)

        def run():
            eng._check(eng._lib.arp_gemm_bf16(eng._h, C.c_void_p(a.data_ptr()), C.c_void_p(w.data_ptr()),
                                              C.c_void_p(out.data_ptr()), capi._TORCH_DT[odt], M_HOT, N, K,
                                              C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()) if res else None, act,
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        if res:
            x0 = x.clone()
            run()
            torch.cuda.synchronize()
            e = relerr(x, x0 + a.float() @ w.float().t() + b)
            print(f"  in-place residual {name}: relerr {e:.2e}")
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(20):
            run()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 20
        print(f"  {name:9s} M={M_HOT} N={N} K={K}: {ms * 1e3:8.1f} us  {2 * M_HOT * N * K / ms / 1e9:7.1f} TFLOP/s", flush=True)
    eng.close()
