"""ORACLE / TEST INFRASTRUCTURE — not part of the shipped product.

Golden vectors for the adapter fine-tuning forward: executes the REFERENCE's unmodified
`finetune_module/clip_multiscale_adapter.py` `CLIPMultiscaleAdapter.forward(batch)` (VIP-I + inverse-dynamics
losses, :179-252) and `.backward()` on a seeded batch, on the same shims / stubs as oracle/run_reference.py
(kornia's ColorJitter is the identity stub: the random train-time augmentation is the caller's business on both
sides), and writes tests/golden/finetune/*.npz:

  loss, the frozen-CLIP intermediates the hooks captured (class-token / EOT rows + encode_image / encode_text), the
  adapted features, every small gradient in full and the norm + a strided sample of every large one.

Build container only (needs /root/reference):   python -m oracle.make_golden_finetune
"""
from __future__ import annotations

import json
import re
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from arp_b200.synth import structured_frames  # noqa: E402
from arp_b200.weights import random_adapter_state_dict  # noqa: E402
from oracle import stubs  # noqa: E402
from oracle.make_golden import digest, state_digest  # noqa: E402
from oracle.run_reference import reference_adapter, reference_clip_state_dict  # noqa: E402

OUT = ROOT / "tests" / "golden" / "finetune"
MAZE = "navigate a maze to collect the yellow cheese."
COIN = "the goal is to collect the coin."

CASES = {
    "f1_text_b16_64": dict(arch="ViT-B/16", seed=0, adapter_seed=1, size=64, batch=2, goal_conditioned=False,
                           texts=[MAZE, COIN], frames_seed=21),
    "f2_goal_b16_64": dict(arch="ViT-B/16", seed=0, adapter_seed=1, size=64, batch=2, goal_conditioned=True,
                           texts=[MAZE, MAZE], frames_seed=22),
}
SAMPLE = 9973   # stride of the sample kept of every large gradient


def make_batch(spec: dict, tokenize):
    """The batch layout of finetune_module/action_finetune_data_procgen.py as forward() consumes it."""
    B = spec["batch"]
    rng = np.random.default_rng(spec["frames_seed"])
    frames = structured_frames(4 * B, spec["size"], rng)
    batch = {f"image{i}": {"ob": torch.from_numpy(frames[i * B:(i + 1) * B].copy())} for i in range(4)}
    batch["instruct"] = tokenize(spec["texts"])
    batch["r"] = torch.from_numpy(rng.integers(0, 2, size=B).astype(np.float32))
    batch["action"] = torch.from_numpy(rng.integers(0, 15, size=B).astype(np.int64))
    return batch, frames


def make_case(name: str):
    spec = dict(CASES[name])
    stubs.import_reference()
    import clip as clip_shim
    clip_sd = reference_clip_state_dict(spec["arch"], spec["seed"])
    sd = random_adapter_state_dict(spec["arch"], seed=spec["adapter_seed"], device="cpu", clip_sd=clip_sd, with_inverse=True)
    orig = torch.nn.init.orthogonal_
    torch.nn.init.orthogonal_ = lambda w, gain=1.0: w                 # the checkpoint overwrites every adapter tensor
    try:
        model = reference_adapter(spec["arch"], spec["seed"], spec["adapter_seed"])
    finally:
        torch.nn.init.orthogonal_ = orig
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("clip_model.") for k in missing), (missing, unexpected)
    model.use_vip_loss = model.use_id_loss = True                     # finetune.py:40-41 defaults
    model.goal_conditioned = spec["goal_conditioned"]
    for p in model.clip_model.parameters():                           # finetune.py:147-148
        p.requires_grad = False
    model.train()                                                     # finetune.py:155
    batch, frames = make_batch(spec, clip_shim.tokenize)

    loss = model(batch)                                               # the reference's own forward, unmodified
    loss.backward()

    out = {"loss": np.float32(loss.item())}
    with torch.no_grad():                                             # the intermediates, for the CPU half of the tests
        roles = 4 if spec["goal_conditioned"] else 3
        imgs = model.preprocess(torch.cat([batch[f"image{i}"]["ob"] for i in range(roles)]), train=True)
        feat = model.clip_model.encode_image(imgs).float()
        taps = torch.cat([v.permute(1, 0, 2)[:, 0, :].float() for k, v in model.activation.items()
                          if re.match(r"visual.transformer.resblocks.[0-9]*$", k)], dim=-1)
        out["image_taps"], out["image_feat"] = taps.numpy(), feat.numpy()
        out["adapted_image"] = model.encode_image(imgs).numpy()
        tok = batch["instruct"]
        tfeat = model.clip_model.encode_text(tok).float()
        ttaps = torch.cat([v.permute(1, 0, 2)[torch.arange(tok.shape[0]), tok.argmax(dim=-1)].float()
                           for k, v in model.activation.items() if re.match(r"transformer.resblocks.[0-9]*$", k)], dim=-1)
        out["text_taps"], out["text_feat"] = ttaps.numpy(), tfeat.numpy()
        out["adapted_text"] = model.encode_text(tok).numpy()
    grads = {}
    for k, p in model.named_parameters():
        if k.startswith("clip_model."):
            assert p.grad is None
            continue
        g = torch.zeros_like(p) if p.grad is None else p.grad
        grads[k] = float(g.norm())
        out["grad/" + k] = g.numpy() if g.numel() <= 20000 else g.reshape(-1)[::SAMPLE].numpy().copy()
    out["instruct"] = batch["instruct"].numpy()
    out["r"], out["action"] = batch["r"].numpy(), batch["action"].numpy()
    meta = dict(spec, name=name, weights_digest=state_digest(sd), frames_digest=digest([frames]), grad_norms=grads,
                sample_stride=SAMPLE, torch=torch.__version__, param_keys=sorted(k for k in model.state_dict()),
                generator="oracle/make_golden_finetune.py -> /root/reference/finetune_module/clip_multiscale_adapter.py (unmodified)")
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / f"{name}.npz", meta=json.dumps(meta), **out)
    print(f"{name}: loss {loss.item():.6f}  grad norms " + ", ".join(f"{k.split('.')[0]}={v:.3g}" for k, v in grads.items()))


if __name__ == "__main__":
    if not stubs.reference_available():
        sys.exit("needs /root/reference (build container only)")
    for n in (sys.argv[1:] or list(CASES)):
        make_case(n)
