"""Dev tool: cuBLAS (torch.matmul, bf16) on the hot GEMM shapes, as the library speed-of-light to compare
tools/gemm_bench.py against. Plain library GEMM: no bias / activation / residual epilogue."""
import torch
M = 197 * 512
dev = "cuda"
for name, N, K in (("qkv", 2304, 768), ("fc", 3072, 768), ("proj", 768, 3072), ("out", 768, 768)):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(a, w.t(), out=out)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(20):
        torch.matmul(a, w.t(), out=out)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 20
    print(f"cuBLAS {name:5s} M={M} N={N} K={K}: {ms * 1e3:8.1f} us  {2 * M * N * K / ms / 1e9:7.1f} TFLOP/s")
