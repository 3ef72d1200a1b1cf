"""clip.tokenize for the host side (reference call sites: label_reward.py:136,138,205,207).

If the real openai `clip` package (pinned at d50d76d in the reference's requirements.txt:17) is importable, its BPE
tokenizer is used. It is not installable offline (needs its vocabulary file and ftfy). Without it `tokenize` REFUSES
(TokenizerUnavailable) unless the caller opts into the deterministic stand-in — `standin=True`, `tokenizer="standin"`
on the entry points, or ARP_ALLOW_STANDIN_TOKENIZER=1 (what the tests, bench.py and smoke() set): with a real
checkpoint the stand-in's ids would give a meaningless instruction embedding and plausible-looking rewards.

The stand-in has clip.tokenize's contract: int32 [n, 77], SOT = 49406 first, one id per lower-cased whitespace token,
EOT = 49407 (the row maximum — CLIP's encode_text picks the EOT position with argmax) and zero padding. It only makes
sense for random-init experiments, where both sides of a parity check use the same ids.
"""
from __future__ import annotations

import os
import zlib

import torch

SOT, EOT, CONTEXT = 49406, 49407, 77


class TokenizerUnavailable(RuntimeError):
    """openai/CLIP's BPE tokenizer is not importable and the caller did not opt into the stand-in."""


def _real_tokenizer():
    try:
        import clip  # type: ignore
    except Exception:  # noqa: BLE001
        return None
    # the oracle's shim package marks itself; a real install has the BPE module next to tokenize
    if getattr(clip, "__arp_oracle_shim__", False) or not hasattr(clip, "simple_tokenizer"):
        return None
    return clip.tokenize


def standin_allowed() -> bool:
    return os.environ.get("ARP_ALLOW_STANDIN_TOKENIZER", "0") not in ("", "0")


def resolve(tokenizer=None):
    """Entry-point argument -> callable(texts) -> int tensor [n, 77]. None: the real tokenizer, else the stand-in only
    if ARP_ALLOW_STANDIN_TOKENIZER=1; "clip": the real one or fail; "standin": the stand-in; a callable: itself."""
    if callable(tokenizer):
        return tokenizer
    if tokenizer in (None, "clip", "standin"):
        standin = {None: None, "clip": False, "standin": True}[tokenizer]
        return lambda texts: tokenize(texts, standin=standin)
    raise ValueError(f"tokenizer must be None, 'clip', 'standin' or a callable, got {tokenizer!r}")


def tokenize(texts, context_length: int = CONTEXT, truncate: bool = False, *,
             standin: "bool | None" = None) -> torch.Tensor:
    real = _real_tokenizer()
    if real is not None and standin is not True:
        return real(texts, context_length=context_length, truncate=truncate)
    if standin is None:
        standin = standin_allowed()
    if not standin:
        raise TokenizerUnavailable(
            "openai/CLIP's tokenizer (package `clip`, requirements.txt:17) is not importable. Install it, pass a "
            "tokenizer callable, or — for random-init experiments only — opt into the deterministic stand-in with "
            "tokenizer='standin' or ARP_ALLOW_STANDIN_TOKENIZER=1. With real CLIP weights the stand-in yields a "
            "meaningless instruction embedding.")
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
