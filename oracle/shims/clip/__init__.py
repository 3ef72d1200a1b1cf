"""ORACLE / TEST INFRASTRUCTURE — not part of the shipped product.

Stand-in for the third-party `clip` package the reference imports but does not vendor:
    clip @ git+https://github.com/openai/CLIP.git@d50d76daa670286dd6cacf3bcd80b5e4823fc8e1
    (/root/reference/requirements.txt:17)
Reference call sites this module has to satisfy:
    arp_dt/label_reward.py:5,126,136,138,141,156,205,207
    finetune_module/clip_multiscale_adapter.py:3,59-61,102,114,119,136,160
    finetune_module/utils.py:14-18 (hook regex on module names `visual.transformer.resblocks.N`)
    arp_dt/envs/vl_reward.py:1,16,18

It restates the published architecture of openai/CLIP `clip/model.py` (ViT image tower, causal text
tower, fp32 LayerNorm, QuickGELU, nn.MultiheadAttention in LND layout, cosine logits scaled by
exp(logit_scale)) with the same parameter names, so a real CLIP state_dict loads into it.
In-tree corroboration of the architecture: arp_dt/models/openai/layers.py:9,12-13,223-250,274-336,339-370
and model.py:59-79,220-244 (the authors' Flax port of the same model).

PARITY PINNING: the reference ships no test or golden vector for this boundary, and the real package
(with its pretrained weights and BPE vocabulary) is not installable here. What pins this restatement:
  * tests/test_oracle_clip.py checks the image tower against an independent implementation,
    transformers.CLIPVisionModelWithProjection, on shared random weights (max-abs ~1e-6);
  * the reference's own files run unmodified on top of it (oracle/run_reference.py).
`tokenize` cannot be reproduced without the BPE vocabulary: it is a deterministic stand-in with the
same shape/dtype contract (SOT=49406, EOT=49407 = row max, zero padding, context 77).
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

__all__ = ["available_models", "load", "tokenize", "CLIP"]
__arp_oracle_shim__ = True   # arp_b200.tokenizer must never mistake this stand-in package for openai/CLIP

_ARCH = {
    # name: (embed_dim, image_resolution, vision_layers, vision_width, vision_patch,
    #        context_length, vocab_size, text_width, text_heads, text_layers)
    "ViT-B/32": (512, 224, 12, 768, 32, 77, 49408, 512, 8, 12),
    "ViT-B/16": (512, 224, 12, 768, 16, 77, 49408, 512, 8, 12),
}

# test knobs (set by the harness, never by the product): which architecture `load` really builds,
# and the RNG seed of the random initialisation.
OVERRIDE_ARCH: str | None = None
SEED: int = 0
_CACHE: dict = {}


def available_models():
    return list(_ARCH)


class LayerNorm(nn.LayerNorm):
    """LayerNorm evaluated in fp32 whatever the activation dtype."""

    def forward(self, x):
        return super().forward(x.float()).to(x.dtype)


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model: int, n_head: int, attn_mask=None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict(
            c_fc=nn.Linear(d_model, 4 * d_model), gelu=QuickGELU(), c_proj=nn.Linear(4 * d_model, d_model)))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask

    def forward(self, x):
        mask = None if self.attn_mask is None else self.attn_mask.to(dtype=x.dtype, device=x.device)
        h = self.ln_1(x)
        x = x + self.attn(h, h, h, need_weights=False, attn_mask=mask)[0]
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, attn_mask=None):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x):
        return self.resblocks(x)


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution, patch_size, width, layers, heads, output_dim):
        super().__init__()
        self.input_resolution, self.output_dim = input_resolution, output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        s = width ** -0.5
        self.class_embedding = nn.Parameter(s * torch.randn(width))
        self.positional_embedding = nn.Parameter(s * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(s * torch.randn(width, output_dim))

    def forward(self, x):
        x = self.conv1(x)                                  # [B, width, g, g]
        x = x.flatten(2).transpose(1, 2)                   # [B, g*g, width]
        cls = self.class_embedding.to(x.dtype).expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], dim=1) + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = self.transformer(x.transpose(0, 1)).transpose(0, 1)   # blocks run in LND layout
        x = self.ln_post(x[:, 0, :])
        return x @ self.proj if self.proj is not None else x


class CLIP(nn.Module):
    def __init__(self, embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size, context_length,
                 vocab_size, transformer_width, transformer_heads, transformer_layers):
        super().__init__()
        self.context_length, self.vocab_size = context_length, vocab_size
        self.visual = VisionTransformer(image_resolution, vision_patch_size, vision_width, vision_layers,
                                        vision_width // 64, embed_dim)
        causal = torch.full((context_length, context_length), float("-inf")).triu_(1)
        self.transformer = Transformer(transformer_width, transformer_layers, transformer_heads, attn_mask=causal)
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width))
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self._init_text_tower()

    def _init_text_tower(self):
        w, n = self.transformer.width, self.transformer.layers
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        for blk in self.transformer.resblocks:
            nn.init.normal_(blk.attn.in_proj_weight, std=w ** -0.5)
            nn.init.normal_(blk.attn.out_proj.weight, std=(w ** -0.5) * ((2 * n) ** -0.5))
            nn.init.normal_(blk.mlp.c_fc.weight, std=(2 * w) ** -0.5)
            nn.init.normal_(blk.mlp.c_proj.weight, std=(w ** -0.5) * ((2 * n) ** -0.5))
        nn.init.normal_(self.text_projection, std=w ** -0.5)

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image):
        return self.visual(image.type(self.dtype))

    def encode_text(self, text):
        x = self.token_embedding(text).type(self.dtype) + self.positional_embedding.type(self.dtype)
        x = self.transformer(x.transpose(0, 1)).transpose(0, 1)
        x = self.ln_final(x).type(self.dtype)
        eot = text.argmax(dim=-1)                           # EOT carries the largest token id
        return x[torch.arange(x.shape[0]), eot] @ self.text_projection

    def forward(self, image, text):
        img = self.encode_image(image)
        txt = self.encode_text(text)
        img = img / img.norm(dim=1, keepdim=True)
        txt = txt / txt.norm(dim=1, keepdim=True)
        logits_per_image = self.logit_scale.exp() * img @ txt.t()
        return logits_per_image, logits_per_image.t()


def build(name: str, seed: int | None = None) -> CLIP:
    """Random-init CLIP of the named architecture, deterministic in `seed` (CPU generator)."""
    seed = SEED if seed is None else seed
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        model = CLIP(*_ARCH[name])
    return model.eval()


def load(name: str, device="cpu", jit: bool = False, download_root=None):
    """clip.load(name, device) -> (model.eval(), preprocess). fp32 on CPU, fp16 weights on CUDA like upstream."""
    arch = OVERRIDE_ARCH or name
    key = (arch, SEED)
    if key not in _CACHE:
        _CACHE[key] = build(arch, SEED)
    import copy
    model = copy.deepcopy(_CACHE[key])
    dev = torch.device(device)
    if dev.type == "cpu":
        model = model.float()
    else:
        model = model.to(dev)
        for m in model.modules():                          # upstream convert_weights: Linear/Conv/MHA -> fp16
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                m.half()
            if isinstance(m, nn.MultiheadAttention):
                for p in (m.in_proj_weight, m.in_proj_bias, m.out_proj.weight, m.out_proj.bias):
                    p.data = p.data.half()
        for attr in ("text_projection",):
            getattr(model, attr).data = getattr(model, attr).data.half()
        model.visual.proj.data = model.visual.proj.data.half()
    return model.eval(), _default_preprocess(model.visual.input_resolution)


def _default_preprocess(n_px: int):
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor
    return Compose([
        Resize(n_px, interpolation=InterpolationMode.BICUBIC), CenterCrop(n_px), lambda im: im.convert("RGB"),
        ToTensor(), Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])


SOT, EOT = 49406, 49407


def tokenize(texts, context_length: int = 77, truncate: bool = False) -> torch.Tensor:
    """Deterministic stand-in for the BPE tokenizer: one id per whitespace-separated lower-cased word,
    crc32-hashed into [256, 49256). Same contract as upstream: int tensor [n, 77], SOT first, EOT (the
    row maximum, which encode_text relies on) after the last word, zero padded."""
    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), context_length, dtype=torch.int)
    for i, t in enumerate(texts):
        ids = [SOT] + [256 + zlib.crc32(w.encode("utf-8")) % 49000 for w in t.lower().split()] + [EOT]
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {t} is too long for context length {context_length}")
            ids = ids[:context_length - 1] + [EOT]
        out[i, :len(ids)] = torch.tensor(ids, dtype=torch.int)
    return out
