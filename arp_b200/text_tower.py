"""Instruction embedding, computed ONCE per labeling run and cached on the device.

The reference re-tokenises and re-runs CLIP's text tower for every episode
(arp_dt/label_reward.py:135-141, :204-212); the result never changes, so here it is evaluated once
with plain PyTorch ops straight from the state_dict (5.96 GFLOP per instruction — not a hot path)
and handed to the native library with arp_set_text.

Restates, functionally, openai/CLIP `CLIP.encode_text` (causal 12-layer width-512 transformer, EOT
pooling by token argmax, text_projection) and, for the adapter heads,
finetune_module/clip_multiscale_adapter.py:153-177 (EOT row of every text block -> linear -> concat
with the final feature -> gated AdapterMLP -> L2 normalise).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _ln(x, sd, prefix):
    return F.layer_norm(x.float(), (x.shape[-1],), sd[prefix + ".weight"].float(), sd[prefix + ".bias"].float(), 1e-5)


def _count_layers(sd, prefix: str) -> int:
    n = 0
    while f"{prefix}resblocks.{n}.ln_1.weight" in sd:
        n += 1
    return n


@torch.no_grad()
def encode_text(sd: dict, tokens: torch.Tensor, device, prefix: str = ""):
    """Returns (text_features [n, embed_dim] un-normalised, eot_taps [n_layers][n, width])."""
    emb_w = sd[prefix + "token_embedding.weight"]
    if emb_w.device.type == "cpu":          # 49408 x 512 table (101 MB): look the few rows up where the table lives
        x0 = emb_w[tokens.to("cpu").long()].to(device).float()
    else:
        x0 = emb_w[tokens.to(emb_w.device).long()].to(device).float()
    sd = {k[len(prefix):]: v.to(device).float() for k, v in sd.items()
          if k.startswith(prefix) and torch.is_tensor(v) and
          k[len(prefix):].startswith(("transformer.", "positional_embedding", "ln_final.", "text_projection"))}
    tok = tokens.to(device).long()
    n, ctx = tok.shape
    x = x0 + sd["positional_embedding"][:ctx]
    width = x.shape[-1]
    heads = width // 64
    dh = width // heads
    causal = torch.full((ctx, ctx), float("-inf"), device=device).triu_(1)
    eot = tok.argmax(dim=-1)
    rows = torch.arange(n, device=device)
    taps = []
    for l in range(_count_layers(sd, "transformer.")):
        p = f"transformer.resblocks.{l}."
        h = _ln(x, sd, p + "ln_1")
        qkv = h @ sd[p + "attn.in_proj_weight"].t() + sd[p + "attn.in_proj_bias"]
        q, k, v = (t.view(n, ctx, heads, dh).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
        att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh) + causal, dim=-1) @ v
        att = att.transpose(1, 2).reshape(n, ctx, width)
        x = x + att @ sd[p + "attn.out_proj.weight"].t() + sd[p + "attn.out_proj.bias"]
        h = _ln(x, sd, p + "ln_2")
        h = h @ sd[p + "mlp.c_fc.weight"].t() + sd[p + "mlp.c_fc.bias"]
        h = h * torch.sigmoid(1.702 * h)
        x = x + h @ sd[p + "mlp.c_proj.weight"].t() + sd[p + "mlp.c_proj.bias"]
        taps.append(x[rows, eot])
    x = _ln(x, sd, "ln_final")
    return x[rows, eot] @ sd["text_projection"], taps


@torch.no_grad()
def clip_text_embedding(sd: dict, tokens: torch.Tensor, device, prefix: str = ""):
    """Unit-norm CLIP text features [n, 512] and exp(logit_scale) (label_reward.py:141 via CLIP.forward)."""
    feat, _ = encode_text(sd, tokens, device, prefix)
    scale = float(sd[prefix + "logit_scale"].float().exp())
    return feat / feat.norm(dim=1, keepdim=True), scale


@torch.no_grad()
def adapter_text_embedding(sd: dict, tokens: torch.Tensor, device, ensemble: bool = False):
    """CLIPMultiscaleAdapter.encode_text for 2-D token input (clip_multiscale_adapter.py:153-177).

    `sd` is the full-module checkpoint (finetune.py:164): CLIP lives under `clip_model.`.
    logit_scale is NOT in that state_dict (plain tensor attribute, :102) -> always CLIP's value.
    With ensemble=True each 512-wide scale is normalised separately (the 3-D branch of label_reward.py:217-222).
    """
    feat, taps = encode_text(sd, tokens, device, prefix="clip_model.")
    g = lambda k: sd[k].to(device).float()  # noqa: E731
    inter = torch.cat(taps, dim=-1) @ g("text_intermediate_linear.weight").t()
    tf = torch.cat([inter, feat], dim=-1)
    h = torch.relu(tf @ g("text_adapter.layers.0.weight").t() + g("text_adapter.layers.0.bias"))
    mlp = h @ g("text_adapter.layers.3.weight").t() + g("text_adapter.layers.3.bias")
    res = torch.sigmoid(g("text_residual_weight"))
    adapted = res * tf + (1.0 - res) * mlp
    if ensemble:
        n = adapted.shape[0]
        adapted = F.normalize(adapted.view(n, -1, feat.shape[-1]), dim=-1).reshape(n, -1)
    else:
        adapted = F.normalize(adapted, dim=-1)
    scale = float(g("clip_model.logit_scale").exp())
    return adapted, scale
