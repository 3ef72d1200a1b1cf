"""CPU study for DESIGN.md §8 item 4: how much cosine error do the operand / storage formats of the image tower cost?

Emulates the kernels' numerics in PyTorch fp32 on the CPU: every GEMM / attention operand is rounded to the studied
16-bit format, products accumulate in fp32 (what tcgen05 kind::f16 does for bf16 AND fp16 operands at the same rate),
LayerNorm and softmax run in fp32, the residual stream is kept in the studied residual format. Random-init ViT-B/16
(the only weights available offline), structured 64x64 frames, adapter-free `clip` head.

  python tools/precision_study.py [n_frames]
"""
import math
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from arp_b200.synth import structured_frames  # noqa: E402
from arp_b200.weights import random_clip_state_dict  # noqa: E402

torch.manual_seed(0)
torch.set_grad_enabled(False)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sd = {k: v.float() for k, v in random_clip_state_dict("ViT-B/16", 0, "cpu").items()}
frames = structured_frames(N, 64, np.random.default_rng(0))
x0 = torch.from_numpy(frames).permute(0, 3, 1, 2).float()
x0 = F.interpolate(x0, size=(224, 224), mode="bicubic", align_corners=False).clamp(0, 255).round() / 255.0
mean = torch.tensor((0.48145466, 0.4578275, 0.40821073)).view(1, 3, 1, 1)
std = torch.tensor((0.26862954, 0.26130258, 0.27577711)).view(1, 3, 1, 1)
x0 = (x0 - mean) / std
text = F.normalize(torch.randn(4, 512), dim=1)


def rnd(t, fmt):
    if fmt == "fp32":
        return t
    return t.to(torch.bfloat16 if fmt == "bf16" else torch.float16).float()


def tower(op: str, resid: str, p_fmt: str | None = None, fold: bool = False):
    """op = operand format of every contraction, resid = storage format of the residual stream, p_fmt = format of the
    softmax weights fed to P.V (defaults to op); fold = LayerNorm folded into the consumer GEMM:
    the operand is the rounded RAW x, the weight is rounded gamma*W, and rstd*(acc - mu*colsum) + beta W^T is applied
    after the contraction."""
    p_fmt = p_fmt or op
    mm = lambda a, w: rnd(a, op) @ rnd(w, op).t()  # noqa: E731

    def ln_mm(x, pfx, w, b):
        if not fold:
            return mm(F.layer_norm(x, (768,), sd[pfx + ".weight"], sd[pfx + ".bias"], 1e-5), w) + b
        g, beta = sd[pfx + ".weight"], sd[pfx + ".bias"]
        wg = rnd(w * g, op)
        acc = rnd(x, op) @ wg.t()
        mu = x.mean(-1, keepdim=True)
        rstd = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + 1e-5)
        return rstd * (acc - mu * wg.sum(1)) + (beta @ w.t() + b)
    W, heads = 768, 12
    patches = F.unfold(x0, 16, stride=16).transpose(1, 2)                        # [N,196,768]
    x = mm(patches, sd["visual.conv1.weight"].reshape(W, -1))
    x = torch.cat([sd["visual.class_embedding"].expand(N, 1, W), x], 1) + sd["visual.positional_embedding"]
    ln = lambda t, p: F.layer_norm(t, (W,), sd[p + ".weight"], sd[p + ".bias"], 1e-5)  # noqa: E731
    x = rnd(ln(x, "visual.ln_pre"), resid)
    for l in range(12):
        p = f"visual.transformer.resblocks.{l}."
        qkv = rnd(ln_mm(x, p + "ln_1", sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]), op)  # stored as operands
        q, k, v = (t.view(N, 197, heads, 64).transpose(1, 2) for t in qkv.chunk(3, -1))
        s = (q @ k.transpose(-1, -2)) * 0.125
        pw = torch.softmax(s, -1)
        pw = rnd(pw * 1.0, p_fmt)
        o = (pw @ v) / pw.sum(-1, keepdim=True)                                 # normalised by the rounded weights' own sum
        o = rnd(o.transpose(1, 2).reshape(N, 197, W), op)
        x = rnd(x + mm(o, sd[p + "attn.out_proj.weight"]) + sd[p + "attn.out_proj.bias"], resid)
        h = ln_mm(x, p + "ln_2", sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])
        h = rnd(h * torch.sigmoid(1.702 * h), op)
        x = rnd(x + mm(h, sd[p + "mlp.c_proj.weight"]) + sd[p + "mlp.c_proj.bias"], resid)
    f = ln(x[:, 0], "visual.ln_post") @ sd["visual.proj"]
    return F.normalize(f, dim=1) @ text.t()


ref = tower("fp32", "fp32")
print(f"{N} frames, |cos| of the reference: mean {ref.abs().mean():.3e}, max {ref.abs().max():.3e}")
print(f"{'operands':8s} {'residual':8s} {'P':5s} {'LN':5s}  max|dcos|   mean|dcos|")
for op, resid, pf, fold in (("bf16", "fp32", None, False), ("fp16", "fp32", None, False), ("fp16", "fp32", "bf16", False),
                            ("bf16", "bf16", None, False), ("bf16", "fp16", None, False), ("fp16", "fp16", None, False),
                            ("fp16", "bf16", None, False), ("bf16", "fp32", None, True), ("fp16", "fp32", None, True),
                            ("fp16", "fp16", None, True)):
    d = (tower(op, resid, pf, fold) - ref).abs()
    print(f"{op:8s} {resid:8s} {(pf or op):5s} {'fold' if fold else 'own':5s}  {d.max():.3e}   {d.mean():.3e}")
