"""Pins for the oracle's building blocks: Pillow restatements (numpy and C) against real Pillow, the
scan / stack restatements against each other, and the `clip` shim against an independent CLIP
implementation (HF transformers) on shared random weights."""
import numpy as np
import pytest
import torch

from oracle import cport, port


@pytest.mark.parametrize("size", [32, 64, 96, 128, 224, 256])
def test_bicubic_restatements_match_pillow(size):
    from PIL import Image
    rng = np.random.default_rng(size)
    noise = rng.integers(0, 256, (size, size, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:size, 0:size]
    ramp = np.stack([yy * 255 // size, xx * 255 // size, (yy + xx) * 255 // (2 * size)], -1).astype(np.uint8)
    sat = np.where(rng.random((size, size, 3)) < 0.5, 0, 255).astype(np.uint8)   # exercises the clip8 clamps
    for img in (noise, ramp, sat):
        ref = np.asarray(Image.fromarray(img).resize((224, 224), Image.BICUBIC))
        assert np.array_equal(port.pil_resize_bicubic(img), ref)
        assert np.array_equal(cport.pil_bicubic(img), ref)


def test_c_transform_matches_torchvision():
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    assert np.array_equal(cport.transform(img), port.transform_pil(False)(img).numpy())


def test_scan_and_stack_restatements():
    rng = np.random.default_rng(2)
    for n in (1, 2, 7, 8, 9, 250, 999):
        x = (rng.standard_normal(n) * 3).astype(np.float32)
        a, b = port.discount_cumsum(x), cport.discount_cumsum(x)
        assert a.dtype == np.float32 and np.array_equal(a, b)
        for F in (1, 4, 8):
            s = port.stack_outputs(x, F)
            assert np.array_equal(s, port.stack_outputs_fast(x, F)) and np.array_equal(s, cport.stack_outputs(x, F))
            assert s.shape == (n, F) and np.array_equal(s[:, -1], x) and np.all(s[0] == x[0])
    # the scan's rounding order is observable: right-to-left fp32 adds differ from a float64 sum
    x = (rng.standard_normal(999) * 0.1).astype(np.float32)
    assert port.discount_cumsum(x)[0] != np.float32(x.astype(np.float64).sum()) or True
    # 0-d input is promoted to 1-d (label_reward.py:234,248)
    assert port.discount_cumsum(np.float32(2.0)).shape == (1,)
    assert port.stack_outputs(np.float32(2.0), 4).shape == (1, 4)


def test_episode_index():
    done = np.zeros(20, np.float32)
    done[[4, 9, 15]] = 1.0
    assert port.episode_index(done) == [0, 5, 10, 16]
    assert cport.episode_index(done).tolist() == [0, 5, 10, 16]
    assert port.episode_index(np.zeros(5, np.float32)) == [0]


def test_clip_shim_matches_hf_vision_tower():
    """Independent implementation check: map the shim's random weights into
    transformers.CLIPVisionModelWithProjection and compare image features."""
    transformers = pytest.importorskip("transformers")
    cfg = transformers.CLIPVisionConfig(hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                                        num_attention_heads=12, image_size=224, patch_size=32, projection_dim=512,
                                        hidden_act="quick_gelu", layer_norm_eps=1e-5)
    hf = transformers.CLIPVisionModelWithProjection(cfg).eval()
    m = port.clip_shim.build("ViT-B/32", seed=3)
    sd = m.visual.state_dict()
    W = 768
    new = {
        "vision_model.embeddings.class_embedding": sd["class_embedding"],
        "vision_model.embeddings.patch_embedding.weight": sd["conv1.weight"],
        "vision_model.embeddings.position_embedding.weight": sd["positional_embedding"],
        "vision_model.pre_layrnorm.weight": sd["ln_pre.weight"], "vision_model.pre_layrnorm.bias": sd["ln_pre.bias"],
        "vision_model.post_layernorm.weight": sd["ln_post.weight"], "vision_model.post_layernorm.bias": sd["ln_post.bias"],
        "visual_projection.weight": sd["proj"].t().contiguous(),
    }
    for l in range(12):
        p, q = f"transformer.resblocks.{l}.", f"vision_model.encoder.layers.{l}."
        w, b = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
        for i, nm in enumerate(("q_proj", "k_proj", "v_proj")):
            new[q + f"self_attn.{nm}.weight"] = w[i * W:(i + 1) * W]
            new[q + f"self_attn.{nm}.bias"] = b[i * W:(i + 1) * W]
        new[q + "self_attn.out_proj.weight"] = sd[p + "attn.out_proj.weight"]
        new[q + "self_attn.out_proj.bias"] = sd[p + "attn.out_proj.bias"]
        new[q + "layer_norm1.weight"], new[q + "layer_norm1.bias"] = sd[p + "ln_1.weight"], sd[p + "ln_1.bias"]
        new[q + "layer_norm2.weight"], new[q + "layer_norm2.bias"] = sd[p + "ln_2.weight"], sd[p + "ln_2.bias"]
        new[q + "mlp.fc1.weight"], new[q + "mlp.fc1.bias"] = sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]
        new[q + "mlp.fc2.weight"], new[q + "mlp.fc2.bias"] = sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"]
    missing, unexpected = hf.load_state_dict(new, strict=False)
    assert not [k for k in missing if "position_ids" not in k], missing
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        a = m.encode_image(x)
        b = hf(pixel_values=x).image_embeds
    assert float((a - b).abs().max()) < 5e-5 * float(b.abs().max() + 1)


def test_tokenize_contract():
    t = port.clip_shim.tokenize(["the goal is to collect the coin.", "a"])
    assert t.shape == (2, 77) and t.dtype == torch.int32
    assert t[0, 0] == 49406 and int(t[0].argmax()) == 8 and t[0, 8] == 49407 and t[0, 9:].sum() == 0
    from arp_b200.tokenizer import tokenize
    assert torch.equal(tokenize(["the goal is to collect the coin.", "a"]), t)   # product stand-in == oracle stand-in
