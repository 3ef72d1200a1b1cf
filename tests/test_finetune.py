"""Adapter fine-tuning forward — SURVEY.md §8(f)4, reference finetune_module/clip_multiscale_adapter.py:179-252.

Fixtures under tests/golden/finetune/ were produced by the reference's UNMODIFIED CLIPMultiscaleAdapter.forward +
backward (oracle/make_golden_finetune.py). CPU tests pin the oracle port and the autograd half of the drop-in
(everything downstream of the frozen CLIP tower) against them; GPU tests run the whole product forward — frozen tower
in libarp_b200.so through arp_encode_taps_chw — against the same fixtures."""
import json

import numpy as np
import pytest
import torch

from _util import GOLDEN, digest, state_digest

FT = GOLDEN / "finetune"
NAMES = sorted(p.stem for p in FT.glob("*.npz"))

# bf16 tensor-core tower vs the reference's fp32 tower:
TOL_TAPS_BF16 = 1e-2      # max|Δ| / max|ref| of the class-token rows of the 12 blocks (measured 3.2e-3 .. 3.6e-3)
TOL_LOSS_BF16 = 5e-3      # absolute, on a loss of ~7 (measured 2.2e-4 golden, 8.1e-4 fresh batch)
TOL_TAPS_F32 = 1e-4       # precision="fp32" verification path
TOL_LOSS_F32 = 2e-4
TOL_CPU = 2e-5            # autograd half on CPU fp32 vs the reference's CPU fp32 (same ops, same order)


_WEIGHTS: dict = {}


def load(name):
    z = np.load(FT / f"{name}.npz", allow_pickle=False)
    return json.loads(str(z["meta"])), {k: z[k] for k in z.files if k != "meta"}


def rebuild(meta):
    """(batch, adapter state_dict incl. inverse layer) as make_golden_finetune.py built them."""
    import clip as clip_shim  # oracle shim (sys.path set by _util)
    from arp_b200.synth import structured_frames
    from arp_b200.weights import random_adapter_state_dict
    B = meta["batch"]
    rng = np.random.default_rng(meta["frames_seed"])
    frames = structured_frames(4 * B, meta["size"], rng)
    assert digest([frames]) == meta["frames_digest"]
    batch = {f"image{i}": {"ob": torch.from_numpy(frames[i * B:(i + 1) * B].copy())} for i in range(4)}
    batch["r"] = torch.from_numpy(rng.integers(0, 2, size=B).astype(np.float32))
    batch["action"] = torch.from_numpy(rng.integers(0, 15, size=B).astype(np.int64))
    key = (meta["arch"], meta["seed"], meta["adapter_seed"])
    if key not in _WEIGHTS:                                  # 476 M adapter parameters: generate once per session
        clip_sd = {k: v.clone() for k, v in clip_shim.build(meta["arch"], meta["seed"]).state_dict().items()}
        sd = random_adapter_state_dict(meta["arch"], seed=meta["adapter_seed"], device="cpu", clip_sd=clip_sd, with_inverse=True)
        assert state_digest(sd) == meta["weights_digest"]
        _WEIGHTS[key] = (sd, clip_sd)
    sd, clip_sd = _WEIGHTS[key]
    return batch, sd, clip_sd


def make_model(meta, sd, clip_sd, device=None, precision="bf16"):
    from arp_b200.finetune import CLIPMultiscaleAdapter
    m = CLIPMultiscaleAdapter(clip_state_dict=clip_sd, arch=meta["arch"], use_discrete_action=True, action_dim=15,
                              use_vip_loss=True, use_id_loss=True, goal_conditioned=meta["goal_conditioned"],
                              init="normal", precision=precision,
                              augmentation=lambda x: x)    # the goldens were made with kornia's random jitter stubbed to identity
    res = m.load_state_dict(sd, strict=True)          # the reference's checkpoint format, every key accounted for
    assert not res.missing_keys and not res.unexpected_keys
    return m.to(device) if device is not None else m


def check_grads(model, gold, meta, rtol, atol_frac):
    stride = meta["sample_stride"]
    for k, p in model.named_parameters():
        ref = gold["grad/" + k]
        g = torch.zeros_like(p) if p.grad is None else p.grad
        g = g.detach().float().cpu()
        got = g.numpy() if g.numel() <= 20000 else g.reshape(-1)[::stride].numpy()
        scale = max(float(np.abs(ref).max()), 1e-30)
        assert np.abs(got - ref).max() <= atol_frac * scale + 1e-12, (k, np.abs(got - ref).max(), scale)
        n_ref = meta["grad_norms"][k]
        assert abs(float(g.norm()) - n_ref) <= rtol * n_ref + 1e-12, (k, float(g.norm()), n_ref)


def test_finetune_fixtures_exist():
    assert NAMES == ["f1_text_b16_64", "f2_goal_b16_64"]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_port_reproduces_reference_finetune_loss(name):
    from oracle import port
    meta, gold = load(name)
    batch, sd, _ = rebuild(meta)
    batch["instruct"] = torch.from_numpy(gold["instruct"])
    ad = port.AdapterOracle(sd, meta["arch"])
    got = port.finetune_loss(ad, batch, goal_conditioned=meta["goal_conditioned"])
    assert abs(got - float(gold["loss"])) <= 1e-5 * abs(float(gold["loss"])), (got, float(gold["loss"]))


@pytest.mark.parametrize("name", NAMES)
def test_autograd_half_matches_reference_cpu(name):
    """Feed the reference's own frozen-CLIP intermediates (the tensors its hooks captured) into the drop-in's trainable
    half: adapted features, loss and every gradient must match the reference's forward/backward."""
    meta, gold = load(name)
    _, sd, clip_sd = rebuild(meta)
    m = make_model(meta, sd, clip_sd)
    assert sorted(m.state_dict()) == meta["param_keys"]                # checkpoint key set == the reference module's
    assert [k for k, _ in m.named_parameters()] == [k for k in meta["grad_norms"]]   # what the optimizer sees, in order
    assert list(m.clip_model.parameters()) == []                       # finetune.py:147-148 has nothing to freeze
    B = meta["batch"]
    ai = m.adapt_image(torch.from_numpy(gold["image_taps"]), torch.from_numpy(gold["image_feat"]))
    assert np.abs(ai.detach().numpy() - gold["adapted_image"]).max() <= TOL_CPU
    at = m.adapt_text(torch.from_numpy(gold["text_taps"]), torch.from_numpy(gold["text_feat"]))
    assert np.abs(at.detach().numpy() - gold["adapted_text"]).max() <= TOL_CPU
    loss = m.loss_from_features(torch.split(ai, B), None if meta["goal_conditioned"] else at,
                                torch.from_numpy(gold["r"]), torch.from_numpy(gold["action"]))
    assert abs(loss.item() - float(gold["loss"])) <= TOL_CPU * abs(float(gold["loss"]))
    loss.backward()
    check_grads(m, gold, meta, rtol=1e-4, atol_frac=1e-4)


def test_frozen_text_tower_matches_reference_cpu():
    meta, gold = load("f1_text_b16_64")
    _, sd, clip_sd = rebuild(meta)
    m = make_model(meta, sd, clip_sd)
    tok = torch.from_numpy(gold["instruct"])
    taps, feat = m.clip_model.text_taps(tok)
    assert np.abs(taps.numpy() - gold["text_taps"]).max() <= 1e-4 * np.abs(gold["text_taps"]).max()
    assert np.abs(feat.numpy() - gold["text_feat"]).max() <= 1e-4 * np.abs(gold["text_feat"]).max()
    assert np.abs(m.encode_text(tok).detach().numpy() - gold["adapted_text"]).max() <= TOL_CPU
    # [B, num_text, ctx] input: mean over the texts of each sample (:174-176)
    t3 = m.encode_text(torch.stack([tok, tok.flip(0)], dim=1)).detach().numpy()
    assert np.abs(t3 - (gold["adapted_text"] + gold["adapted_text"][::-1]) / 2).max() <= TOL_CPU
    assert len(m.clip_model._text_cache) == 2                           # the instruction set is cached per token row


def test_state_dict_round_trip_and_frozen_tower_needs_gpu(tmp_path):
    meta, _ = load("f1_text_b16_64")
    _, sd, clip_sd = rebuild(meta)
    m = make_model(meta, sd, clip_sd)
    out = m.state_dict()
    assert set(out) == set(sd) and all(torch.equal(out[k], sd[k]) for k in sd)
    # label_reward's loader reads the same file: torch.load(...) -> load_state_dict(strict=False) (label_reward.py:174-176)
    torch.save({k: v for k, v in out.items() if not k.startswith(("image_adapter", "text_adapter"))}, tmp_path / "c.pt")
    m2 = make_model(meta, sd, clip_sd)
    res = m2.load_state_dict(torch.load(tmp_path / "c.pt"), strict=False)
    assert all(k.startswith(("image_adapter", "text_adapter")) for k in res.missing_keys) and not res.unexpected_keys
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m.encode_image(torch.zeros(1, 3, 224, 224))


def test_preprocess_matches_oracle_and_augmentation_hook():
    from arp_b200.finetune import CLIPMultiscaleAdapter
    from oracle import port
    meta, _ = load("f1_text_b16_64")
    batch, sd, clip_sd = rebuild(meta)
    calls = []
    m = CLIPMultiscaleAdapter(clip_state_dict=clip_sd, init="normal", augmentation=lambda x: (calls.append(x.shape), x * 0.5)[1])
    x = batch["image0"]["ob"]
    assert torch.equal(m.preprocess(x), port.preprocess_bilinear(x)) and not calls
    y = m.preprocess(x, train=True)                                     # augmentation sees the resized 0..255 image (:127-129)
    assert calls == [torch.Size([2, 3, 224, 224])] and not torch.equal(y, m.preprocess(x))
    x224 = torch.zeros(1, 224, 224, 3, dtype=torch.uint8)
    assert m.preprocess(x224).shape == (1, 3, 224, 224)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("name", NAMES)
def test_finetune_forward_backward_matches_reference_gpu(built_lib, name, precision):
    meta, gold = load(name)
    batch, sd, clip_sd = rebuild(meta)
    dev = torch.device("cuda", 0)
    m = make_model(meta, sd, clip_sd, device=dev, precision=precision)
    gb = {k: ({kk: vv.to(dev) for kk, vv in v.items()} if isinstance(v, dict) else v.to(dev)) for k, v in batch.items()}
    gb["instruct"] = torch.from_numpy(gold["instruct"]).to(dev)
    roles = 4 if meta["goal_conditioned"] else 3
    imgs = m.preprocess(torch.cat([gb[f"image{i}"]["ob"] for i in range(roles)]), train=True)
    n0 = m.clip_model._get_engine(dev).launch_count
    taps, feat = m.clip_model.image_taps(imgs)
    assert m.clip_model._get_engine(dev).launch_count > n0              # the native kernels ran
    tol_t, tol_l = (TOL_TAPS_BF16, TOL_LOSS_BF16) if precision == "bf16" else (TOL_TAPS_F32, TOL_LOSS_F32)
    et = np.abs(taps.cpu().numpy() - gold["image_taps"]).max() / np.abs(gold["image_taps"]).max()
    ef = np.abs(feat.cpu().numpy() - gold["image_feat"]).max() / np.abs(gold["image_feat"]).max()
    m.train()
    loss = m(gb)
    el = abs(loss.item() - float(gold["loss"]))
    print(f"{name} [{precision}]: taps rel-to-max {et:.2e}, feat {ef:.2e}, |dloss| {el:.2e} on {float(gold['loss']):.4f}")
    assert et <= tol_t and ef <= tol_t and el <= tol_l
    loss.backward()
    if precision == "fp32":
        check_grads(m, gold, meta, rtol=2e-3, atol_frac=5e-3)
    else:                                                                # bf16 tower: direction and size of every gradient
        for k, p in m.named_parameters():
            n_ref = meta["grad_norms"][k]
            if n_ref == 0.0:
                assert p.grad is None or float(p.grad.norm()) == 0.0, k
                continue
            # the two gate scalars are differences of nearly equal terms (sigmoid(4) = 0.982): measured 5 % off, the rest < 1 %
            tol = 0.15 if k.endswith("residual_weight") else 0.03
            assert abs(float(p.grad.norm()) - n_ref) <= tol * n_ref, (k, float(p.grad.norm()), n_ref)
            ref = torch.from_numpy(gold["grad/" + k]).reshape(-1)
            g = p.grad.detach().float().cpu().reshape(-1)
            g = g if g.numel() <= 20000 else g[::meta["sample_stride"]]
            if ref.numel() > 1:
                cos = float(torch.dot(g, ref) / (g.norm() * ref.norm()))
                assert cos >= 0.98, (k, cos)          # measured 0.9927 .. 0.9999
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=0.001)   # finetune.py:149
    opt.step()
    opt.zero_grad()
    assert float(m(gb)) != float(loss)                                   # the step moved the trainable half


@pytest.mark.gpu
def test_finetune_product_matches_oracle_port_on_fresh_batch_gpu(built_lib):
    """Not a golden: a different seeded batch (128x128 frames, batch 3), product on the GPU vs the CPU port."""
    from oracle import port
    meta, gold = load("f1_text_b16_64")
    _, sd, clip_sd = rebuild(meta)
    from arp_b200.synth import structured_frames
    rng = np.random.default_rng(77)
    B = 3
    frames = structured_frames(4 * B, 128, rng)
    batch = {f"image{i}": {"ob": torch.from_numpy(frames[i * B:(i + 1) * B].copy())} for i in range(4)}
    batch["instruct"] = torch.from_numpy(np.concatenate([gold["instruct"], gold["instruct"][:1]]))
    batch["r"] = torch.from_numpy(rng.integers(0, 2, size=B).astype(np.float32))
    batch["action"] = torch.from_numpy(rng.integers(0, 15, size=B).astype(np.int64))
    want = port.finetune_loss(port.AdapterOracle(sd, meta["arch"]), batch)
    dev = torch.device("cuda", 0)
    m = make_model(meta, sd, clip_sd, device=dev)
    gb = {k: ({kk: vv.to(dev) for kk, vv in v.items()} if isinstance(v, dict) else v.to(dev)) for k, v in batch.items()}
    got = float(m(gb))
    print(f"fresh batch: product {got:.5f} vs port {want:.5f}")
    assert abs(got - want) <= TOL_LOSS_BF16
