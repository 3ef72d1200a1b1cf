import sys, torch
from pathlib import Path
sys.path.insert(0, '/root/repo')
from arp_b200 import capi
eng = capi.Engine(device=0, max_batch=8)
dev = torch.device("cuda", 0)
for B in (48, 96, 192, 512, 1024):
    L = 197
    qkv = (torch.randn(B * L, 2304, device=dev) * 1.5).bfloat16()
    for _ in range(5):
        eng.attention(qkv, B, L)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    n = max(10, 5120 // B)
    for _ in range(n):
        eng.attention(qkv, B, L)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / n
    items = B * 12
    waves = -(-items // 148)
    print(f"B={B}: {ms*1e3:.1f} us, qkv {B*L*2304*2/1e6:.0f} MB, {ms*1e3/waves:.2f} us per item-wave ({waves} waves), {ms*1e3*148/items:.2f} us per item per SM")
eng.close()
