"""Bring-up check for every native kernel against plain PyTorch on the same GPU.

Dev tool (not a test, not the bench): run under gpurun, writes gpurun_out/kernel_check.json and prints
one line per check. Each check is isolated so one failing kernel does not hide the others.
"""
from __future__ import annotations

import json
import os
import sys
import time
import traceback
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import os as _os
_os.environ.setdefault("ARP_ALLOW_STANDIN_TOKENIZER", "1")   # random-init weights: the deterministic stand-in token ids
sys.path.insert(0, str(ROOT / "oracle" / "shims"))

from arp_b200 import capi  # noqa: E402

OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
report: dict = {}
dev = torch.device("cuda", 0)


def check(name):
    def deco(fn):
        t0 = time.time()
        try:
            res = fn()
            res["ok"] = bool(res.get("ok", True))
        except Exception as e:  # noqa: BLE001
            res = {"ok": False, "error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc()[-1500:]}
        res["secs"] = round(time.time() - t0, 2)
        report[name] = res
        print(f"[{'PASS' if res['ok'] else 'FAIL'}] {name}: " + json.dumps({k: v for k, v in res.items() if k != 'trace'}), flush=True)
        (OUT / "kernel_check.json").write_text(json.dumps(report, indent=1))
        return fn
    return deco


eng = capi.Engine(device=0, patch=16, in_h=64, in_w=64, max_batch=64)


def relerr(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ---------------------------------------------------------------- GEMM
def gemm_case(M, N, K, act=0, bias=False, resid=False, out_dtype=torch.bfloat16, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    a = (torch.randn(M, K, device=dev, generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device=dev, generator=g) if bias else None
    r = torch.randn(M, N, device=dev, generator=g) if resid else None
    out = eng.gemm(a, w, bias=b, resid=r, act=act, out_dtype=out_dtype)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    if bias:
        ref = ref + b
    if act == 1:
        ref = ref * torch.sigmoid(1.702 * ref)
    elif act == 2:
        ref = torch.relu(ref)
    if resid:
        ref = ref + r
    err = relerr(out, ref)
    tol = 1e-2 if out_dtype == torch.bfloat16 else 2e-5
    return {"ok": err < tol and bool(torch.isfinite(out.float()).all()), "relerr": err, "shape": [M, N, K]}


@check("gemm_128x256x64_f32")
def _():
    return gemm_case(128, 256, 64, out_dtype=torch.float32)


@check("gemm_128x256x768_f32")
def _():
    return gemm_case(128, 256, 768, out_dtype=torch.float32)


@check("gemm_tail_M197_f32")
def _():
    return gemm_case(197, 768, 768, out_dtype=torch.float32, bias=True)


@check("gemm_multi_tile_persistent_f32")
def _():
    return gemm_case(197 * 160, 768, 768, out_dtype=torch.float32, bias=True, resid=True)


@check("gemm_qkv_bf16")
def _():
    return gemm_case(197 * 64, 2304, 768, bias=True)


@check("gemm_fc_gelu_bf16")
def _():
    return gemm_case(197 * 64, 3072, 768, act=1, bias=True)


@check("gemm_proj_k3072_resid_f32")
def _():
    return gemm_case(197 * 64, 768, 3072, bias=True, resid=True, out_dtype=torch.float32)


@check("gemm_relu_bf16_adapter_shape")
def _():
    return gemm_case(100, 13312, 6656, act=2, bias=True)


# ---------------------------------------------------------------- LayerNorm
@check("layernorm_bf16")
def _():
    x = torch.randn(1000, 768, device=dev) * 3 + 1
    g = torch.randn(768, device=dev)
    b = torch.randn(768, device=dev)
    y = eng.layernorm(x, g, b)
    ref = torch.nn.functional.layer_norm(x, (768,), g, b, 1e-5)
    err = relerr(y, ref)
    return {"ok": err < 8e-3, "relerr": err}


# ---------------------------------------------------------------- attention
def attn_case(B, L):
    qkv = (torch.randn(B * L, 2304, device=dev) * 1.5).bfloat16()
    out = eng.attention(qkv, B, L)
    q, k, v = qkv.float().view(B, L, 3, 12, 64).permute(2, 0, 3, 1, 4)
    ref = torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v
    ref = ref.permute(0, 2, 1, 3).reshape(B * L, 768)
    err = relerr(out, ref)
    return {"ok": err < 2e-2, "relerr": err}


@check("attention_L197")
def _():
    return attn_case(5, 197)


@check("attention_L50")
def _():
    return attn_case(7, 50)


# ---------------------------------------------------------------- scan
@check("scan_stack_bitexact")
def _():
    rng = np.random.default_rng(0)
    lens = rng.integers(1, 300, size=40).tolist() + [2500, 1, 2048, 2049]
    off = np.concatenate([[0], np.cumsum(lens)])
    T = int(off[-1])
    r = rng.standard_normal(T).astype(np.float32)
    F = 8
    g, rs, gs = eng.scan_only(torch.from_numpy(r), torch.from_numpy(off), F)
    g, rs, gs = g.cpu().numpy(), rs.cpu().numpy(), gs.cpu().numpy()
    ref_g = np.zeros_like(r)
    ref_rs = np.zeros((T, F), np.float32)
    ref_gs = np.zeros((T, F), np.float32)
    for e in range(len(lens)):
        lo, hi = off[e], off[e + 1]
        x = r[lo:hi]
        dc = np.zeros_like(x)
        dc[-1] = x[-1]
        for t in reversed(range(len(x) - 1)):
            dc[t] = x[t] + np.float32(1.0) * dc[t + 1]
        ref_g[lo:hi] = dc
        for i in range(len(x)):
            for f in range(F):
                j = max(0, i - (F - 1 - f))
                ref_rs[lo + i, f] = x[j]
                ref_gs[lo + i, f] = dc[j]
    return {"ok": bool((g == ref_g).all() and (rs == ref_rs).all() and (gs == ref_gs).all()),
            "rtg_mismatch": int((g != ref_g).sum()), "rs_mismatch": int((rs != ref_rs).sum()),
            "gs_mismatch": int((gs != ref_gs).sum())}


# ---------------------------------------------------------------- decode
def pil_transform(img, crop):
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToPILImage, ToTensor
    steps = [ToPILImage()]
    if crop:
        steps.append(CenterCrop(img.shape[0] // 2))
    steps += [Resize(224, interpolation=InterpolationMode.BICUBIC)]
    if not crop:
        steps.append(CenterCrop(224))
    steps += [lambda im: im.convert("RGB"), ToTensor(),
              Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))]
    return Compose(steps)(img)


def decode_case(H, crop, F=4, T=6, pre=capi.PRE_PIL_BICUBIC):
    e = capi.Engine(device=0, patch=16, in_h=H, in_w=H, use_crop=crop, preprocess=pre, max_batch=8,
                    head=capi.HEAD_ADAPTER if pre == capi.PRE_BILINEAR else capi.HEAD_CLIP)
    rng = np.random.default_rng(H + crop)
    ob = rng.integers(0, 256, size=(T, F, H, H, 3), dtype=np.uint8)
    # a smooth image too, so clamping / rounding paths differ from white noise
    yy, xx = np.mgrid[0:H, 0:H]
    ob[0, -1] = np.stack([(yy * 255 // H), (xx * 255 // H), ((yy + xx) * 255 // (2 * H))], -1).astype(np.uint8)
    out = e.decode_only(torch.from_numpy(ob).to(dev)).cpu()
    if pre == capi.PRE_PIL_BICUBIC:
        ref = torch.stack([pil_transform(ob[t, -1], crop) for t in range(T)])
        nbad = int((out != ref).sum())
        res = {"ok": nbad == 0, "mismatch": nbad, "maxabs": float((out - ref).abs().max())}
    else:
        from torchvision.transforms.functional import normalize, resize
        x = torch.from_numpy(ob[:, -1])
        if crop:
            cs = H // 2
            s = int((H - cs) / 2)
            x = x[:, s:s + cs, s:s + cs]
        x = x.permute(0, 3, 1, 2).float()
        x = resize(x, (224, 224), antialias=False) if x.shape[2] != 224 else x
        ref = normalize(x / 255.0, (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))
        d = float((out - ref).abs().max())
        res = {"ok": d < 2e-5, "maxabs": d}
    e.close()
    return res


for H in (64, 256, 128, 224):
    for crop in (False, True):
        check(f"decode_pil_bicubic_{H}_crop{int(crop)}")(lambda H=H, crop=crop: decode_case(H, crop))
for H in (64, 256, 224):
    check(f"decode_bilinear_{H}")(lambda H=H: decode_case(H, False, pre=capi.PRE_BILINEAR))
check("decode_bilinear_64_crop")(lambda: decode_case(64, True, pre=capi.PRE_BILINEAR))


# ---------------------------------------------------------------- end-to-end vs the fp32 oracle model
def e2e_case(arch, patch, H, T=24):
    import clip as clip_shim  # oracle shim
    model = clip_shim.build(arch, seed=0).to(dev).float()
    e = capi.Engine(device=0, patch=patch, in_h=H, in_w=H, max_batch=16)
    missing = e.load_state_dict(model.state_dict())
    if missing:
        return {"ok": False, "missing": missing[:5]}
    rng = np.random.default_rng(1)
    ob = rng.integers(0, 256, size=(T, 2, H, H, 3), dtype=np.uint8)
    tok = clip_shim.tokenize(["the goal is to collect the coin.", "navigate a maze to collect the yellow cheese."]).to(dev)
    with torch.no_grad():
        txt = model.encode_text(tok).float()
        txt = txt / txt.norm(dim=1, keepdim=True)
        imgs = torch.stack([pil_transform(ob[t, -1], False) for t in range(T)]).to(dev)
        f_ref = model.encode_image(imgs)
        _, lpt = model(imgs, tok)
    e.set_text(txt, float(model.logit_scale.exp()))
    obd = torch.from_numpy(ob).to(dev)
    f = e.encode_image(obd)
    r, lg = e.compute_reward(obd, want_logits=True)
    torch.cuda.synchronize()
    feat_err = relerr(f, f_ref)
    cos = torch.nn.functional.cosine_similarity(f, f_ref, dim=1).min().item()
    r_err = float((lg - lpt.t()).abs().max() / lpt.abs().max())
    e.close()
    return {"ok": feat_err < 3e-2 and cos > 0.999, "feat_relerr": feat_err, "min_cos": cos, "logit_relerr_vs_max": r_err,
            "reward_first_ok": bool(torch.equal(r, lg[:, 0]))}


@check("e2e_vitb32_64px")
def _():
    return e2e_case("ViT-B/32", 32, 64)


@check("e2e_vitb16_64px")
def _():
    return e2e_case("ViT-B/16", 16, 64)


@check("e2e_vitb16_256px")
def _():
    return e2e_case("ViT-B/16", 16, 256, T=20)


# ---------------------------------------------------------------- quick timing (not a benchmark number)
@check("timing_vitb16_256px_b256")
def _():
    import clip as clip_shim
    model = clip_shim.build("ViT-B/16", seed=0)
    e = capi.Engine(device=0, patch=16, in_h=256, in_w=256, max_batch=256)
    e.load_state_dict(model.state_dict())
    e.set_text(torch.nn.functional.normalize(torch.randn(1, 512), dim=1), 14.2857)
    T = 1024
    ob = torch.randint(0, 256, (T, 1, 256, 256, 3), dtype=torch.uint8, device=dev)
    e.compute_reward(ob)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    e.compute_reward(ob)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    fps = T / ms * 1e3
    e.close()
    return {"ok": True, "ms": ms, "frames_per_s": fps, "tflops": fps * 35.127e9 / 1e12}


nfail = sum(1 for v in report.values() if not v["ok"])
print(f"SUMMARY: {len(report) - nfail}/{len(report)} passed")
sys.exit(0)
