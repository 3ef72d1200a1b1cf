"""clip.tokenize for the host side (reference call sites: label_reward.py:136,138,205,207).

If the real openai `clip` package (pinned at d50d76d in the reference's requirements.txt:17) is
importable, its BPE tokenizer is used. It is not installable offline (needs its vocabulary file and
ftfy), so otherwise a deterministic stand-in with the same contract is used: int32 [n, 77],
SOT = 49406 first, one id per lower-cased whitespace token, EOT = 49407 (the row maximum — CLIP's
encode_text picks the EOT position with argmax) and zero padding. The stand-in only matters for
random-init experiments; the text tower is outside the per-frame hot loop either way.
"""
from __future__ import annotations

import zlib

import torch

SOT, EOT, CONTEXT = 49406, 49407, 77


def _real_tokenizer():
    try:
        import clip  # type: ignore
        if getattr(clip, "__file__", "") and "oracle" not in clip.__file__ and hasattr(clip, "simple_tokenizer"):
            return clip.tokenize
    except Exception:  # noqa: BLE001
        pass
    return None


def tokenize(texts, context_length: int = CONTEXT, truncate: bool = False) -> torch.Tensor:
    real = _real_tokenizer()
    if real is not None:
        return real(texts, context_length=context_length, truncate=truncate)
    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), context_length, dtype=torch.int32)
    for row, text in enumerate(texts):
        words = text.lower().split()
        ids = [SOT, *(256 + zlib.crc32(w.encode("utf-8")) % 49000 for w in words), EOT]
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {text} is too long for context length {context_length}")
            ids = ids[: context_length - 1] + [EOT]
        out[row, : len(ids)] = torch.tensor(ids, dtype=torch.int32)
    return out
