"""Dev tool: timeline of the attention kernel's block 0. Needs a -DARP_ATTN_TRACE build of the library:
   nvcc <arp_b200.build.NVCC_FLAGS> -DARP_ATTN_TRACE -o arp_b200/_lib/libarp_b200_trace.so arp_b200/csrc/capi.cu
   ARP_B200_LIB=$PWD/arp_b200/_lib/libarp_b200_trace.so python tools/attn_trace.py"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from arp_b200 import capi  # noqa: E402

eng = capi.Engine(device=0, max_batch=8)
lib = eng._lib
B, L = 512, 197
qkv = (torch.randn(B * L, 2304, device="cuda") * 1.5).to(capi.operand_dtype())
ITEMS, EV = 10, 36
buf = np.zeros(2 * ITEMS * EV, np.int64)
for _ in range(3):
    eng.attention(qkv, B, L)
torch.cuda.synchronize()
n = lib.arp_debug_attn_trace(C.c_void_p(buf.ctypes.data), len(buf))
tr = buf.reshape(2, ITEMS, EV)
t0 = tr[tr > 0].min()
names = ["S_issue", "PV_issue", "S_ready", "pass1_done", "turn", "P_arrive", "-", "-"] + [f"pass2_end.q{q}" for q in range(4)] + [f"O_seen.q{q}" for q in range(4)] + [f"stored.q{q}" for q in range(4)] + [f"st_waited.q{q}" for q in range(4)] + ["Sx_begin", "-", "-", "Sx_commit", "PVx_begin", "-", "PVx_commit", "OBS_S_ready", "OBS_O_ready", "-", "-", "-"]
rows = [(tr[s, i, e] - t0, s, i, names[e]) for s in range(2) for i in range(ITEMS) for e in range(EV) if tr[s, i, e] > 0]
for t, s, i, nm in sorted(rows):
    if 3 <= i < 6 and nm != '-':
        print(f"{t:8d}  {'    ' * 6 * s}slot{s} item{i} {nm}")
eng.close()
