"""State-dict helpers: random-init CLIP / adapter weights with the reference's key names, and
checkpoint loading for `model_ckpt_dir` (label_reward.py:174-176: torch.load of the full-module
state_dict written by finetune_module/finetune.py:162-165).

Random init exists because there is no network for pretrained checkpoints: benchmarks and parity
tests use random weights of the right architecture (same shapes, PyTorch-default-like scales).
"""
from __future__ import annotations

import math

import torch

ARCH = {
    # name: (patch, vision width, vision layers, embed_dim, text width, text layers, context, vocab)
    "ViT-B/16": (16, 768, 12, 512, 512, 12, 77, 49408),
    "ViT-B/32": (32, 768, 12, 512, 512, 12, 77, 49408),
}


def _uniform(gen, shape, bound, device):
    return (torch.rand(shape, generator=gen, device=device) * 2 - 1) * bound


def _tower(sd, prefix, width, layers, gen, device, text: bool):
    for l in range(layers):
        p = f"{prefix}resblocks.{l}."
        if text:  # CLIP.initialize_parameters normal inits
            sd[p + "attn.in_proj_weight"] = torch.randn(3 * width, width, generator=gen, device=device) * width ** -0.5
            sd[p + "attn.out_proj.weight"] = torch.randn(width, width, generator=gen, device=device) * (width ** -0.5) * ((2 * layers) ** -0.5)
            sd[p + "mlp.c_fc.weight"] = torch.randn(4 * width, width, generator=gen, device=device) * (2 * width) ** -0.5
            sd[p + "mlp.c_proj.weight"] = torch.randn(width, 4 * width, generator=gen, device=device) * (width ** -0.5) * ((2 * layers) ** -0.5)
        else:     # PyTorch defaults: xavier-uniform in_proj, kaiming-uniform(a=sqrt 5) linears
            sd[p + "attn.in_proj_weight"] = _uniform(gen, (3 * width, width), math.sqrt(6.0 / (4 * width)), device)
            sd[p + "attn.out_proj.weight"] = _uniform(gen, (width, width), width ** -0.5, device)
            sd[p + "mlp.c_fc.weight"] = _uniform(gen, (4 * width, width), width ** -0.5, device)
            sd[p + "mlp.c_proj.weight"] = _uniform(gen, (width, 4 * width), (4 * width) ** -0.5, device)
        sd[p + "attn.in_proj_bias"] = torch.zeros(3 * width, device=device)
        sd[p + "attn.out_proj.bias"] = torch.zeros(width, device=device)
        sd[p + "mlp.c_fc.bias"] = _uniform(gen, (4 * width,), width ** -0.5, device)
        sd[p + "mlp.c_proj.bias"] = _uniform(gen, (width,), (4 * width) ** -0.5, device)
        for ln in ("ln_1", "ln_2"):
            sd[p + ln + ".weight"] = torch.ones(width, device=device)
            sd[p + ln + ".bias"] = torch.zeros(width, device=device)


def random_clip_state_dict(arch: str = "ViT-B/16", seed: int = 0, device="cpu", with_text: bool = True) -> dict:
    """Random CLIP weights keyed like openai/CLIP's state_dict."""
    patch, vw, vl, ed, tw, tl, ctx, vocab = ARCH[arch]
    device = torch.device(device)
    gen = torch.Generator(device=device).manual_seed(seed)
    sd: dict[str, torch.Tensor] = {}
    s = vw ** -0.5
    sd["visual.conv1.weight"] = _uniform(gen, (vw, 3, patch, patch), (3 * patch * patch) ** -0.5, device)
    sd["visual.class_embedding"] = s * torch.randn(vw, generator=gen, device=device)
    sd["visual.positional_embedding"] = s * torch.randn((224 // patch) ** 2 + 1, vw, generator=gen, device=device)
    for ln in ("visual.ln_pre", "visual.ln_post"):
        sd[ln + ".weight"] = torch.ones(vw, device=device)
        sd[ln + ".bias"] = torch.zeros(vw, device=device)
    _tower(sd, "visual.transformer.", vw, vl, gen, device, text=False)
    sd["visual.proj"] = s * torch.randn(vw, ed, generator=gen, device=device)
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07), device=device)
    if with_text:
        sd["token_embedding.weight"] = 0.02 * torch.randn(vocab, tw, generator=gen, device=device)
        sd["positional_embedding"] = 0.01 * torch.randn(ctx, tw, generator=gen, device=device)
        _tower(sd, "transformer.", tw, tl, gen, device, text=True)
        sd["ln_final.weight"] = torch.ones(tw, device=device)
        sd["ln_final.bias"] = torch.zeros(tw, device=device)
        sd["text_projection"] = tw ** -0.5 * torch.randn(tw, ed, generator=gen, device=device)
    return sd


def random_adapter_state_dict(arch: str = "ViT-B/16", seed: int = 0, device="cpu", clip_sd: dict | None = None,
                              with_inverse: bool = False, action_dim: int = 15) -> dict:
    """Random CLIPMultiscaleAdapter checkpoint (inference keys only), keyed like
    finetune_module/clip_multiscale_adapter.py's state_dict with CLIP nested under `clip_model.`.
    with_inverse adds the training-only tensors (`inverse_layer.*`, `lambda_id`; :91-96, :111) AFTER the inference
    ones, so the inference tensors are the same values either way."""
    patch, vw, vl, ed, tw, tl, ctx, vocab = ARCH[arch]
    device = torch.device(device)
    gen = torch.Generator(device=device).manual_seed(seed + 7919)
    clip_sd = clip_sd if clip_sd is not None else random_clip_state_dict(arch, seed, device)
    sd = {"clip_model." + k: v for k, v in clip_sd.items()}
    D, mid = (vl + 1) * ed, vl * tw
    sd["image_intermediate_linear.weight"] = _uniform(gen, (mid, vl * vw), (vl * vw) ** -0.5, device)
    sd["text_intermediate_linear.weight"] = _uniform(gen, (mid, tl * tw), (tl * tw) ** -0.5, device)
    for side in ("image", "text"):
        # orthogonal init is O(n^3); a scaled gaussian has the same spectrum statistics for a benchmark
        sd[f"{side}_adapter.layers.0.weight"] = torch.randn(2 * D, D, generator=gen, device=device) * math.sqrt(2.0 / (2 * D))
        sd[f"{side}_adapter.layers.0.bias"] = torch.zeros(2 * D, device=device)
        sd[f"{side}_adapter.layers.3.weight"] = torch.randn(D, 2 * D, generator=gen, device=device) * math.sqrt(1.0 / (2 * D))
        sd[f"{side}_adapter.layers.3.bias"] = torch.zeros(D, device=device)
        sd[f"{side}_residual_weight"] = torch.tensor(4.0, device=device)
    if with_inverse:
        hid = 2 * ed   # AdapterMLP(input_dim=4*output_dim*(L+1), hidden_dim=1024, output_dim=action_dim)
        sd["inverse_layer.layers.0.weight"] = torch.randn(hid, 4 * D, generator=gen, device=device) * math.sqrt(2.0 / (4 * D))
        sd["inverse_layer.layers.0.bias"] = torch.zeros(hid, device=device)
        sd["inverse_layer.layers.3.weight"] = torch.randn(action_dim, hid, generator=gen, device=device) * math.sqrt(1.0 / hid)
        sd["inverse_layer.layers.3.bias"] = torch.zeros(action_dim, device=device)
        sd["lambda_id"] = torch.tensor(math.log(1 / 0.07), device=device)
    return sd


def load_checkpoint(path, map_location="cpu") -> dict:
    """torch.load(model_ckpt_dir) (label_reward.py:175). Accepts a bare state_dict or {'state_dict': ...}."""
    obj = torch.load(path, map_location=map_location, weights_only=True)
    if isinstance(obj, dict) and "state_dict" in obj and isinstance(obj["state_dict"], dict):
        obj = obj["state_dict"]
    return obj
