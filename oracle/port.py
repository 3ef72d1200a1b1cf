"""ORACLE / TEST INFRASTRUCTURE — not part of the shipped product.

CPU restatement of the reference's reward-labeling algorithm, self-contained so that it travels to
the GPU box where /root/reference does not exist. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product (arp_b200/) never does.

Every function cites the reference lines it follows. Float work runs on torch CPU fp32 exactly as the
reference does on a CPU device; integer / byte work (episode index, Pillow resample, window stack)
is numpy. The model arithmetic is the `clip` shim (oracle/shims/clip) — see its header for how that
third-party dependency is restated and pinned.

PINNING: oracle/make_golden.py runs the reference's unmodified files in the build container and
commits their outputs under tests/golden/; tests/test_oracle_golden.py requires this port to
reproduce them (bit-exact for index/byte/scan work, fp32 round-off for the model).
"""
from __future__ import annotations

import math
import sys
from collections import deque
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

_SHIMS = str(Path(__file__).resolve().parent / "shims")
if _SHIMS not in sys.path:
    sys.path.insert(0, _SHIMS)
import clip as clip_shim  # noqa: E402  (oracle/shims/clip)

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


# ---------------------------------------------------------------------------------------------------
# a1: episode index                                                   arp_dt/label_reward.py:71-87, :267
# ---------------------------------------------------------------------------------------------------
def episode_index(done_last: np.ndarray) -> list[int]:
    """g_traj_idx = [0] + list(nonzero(done[:, -1]) + 1); episode i = rows [idx[i], min(idx[i+1], len_data))."""
    idx = list(np.nonzero(done_last)[0] + 1)
    idx.insert(0, 0)
    return [int(i) for i in idx]


# ---------------------------------------------------------------------------------------------------
# a3: Pillow bicubic resample, restated in integers    label_reward.py:109-121 -> PIL ImagingResample (8bpc)
# ---------------------------------------------------------------------------------------------------
PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float, a: float = -0.5) -> float:
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_coeffs(in_size: int, out_size: int):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the bicubic filter (support 2), full-image box."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int64)
    kk = np.zeros((out_size, ksize), np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def pil_resize_bicubic(img: np.ndarray, out_size=224) -> np.ndarray:
    """uint8 [H,W,3] -> uint8 [out_h,out_w,3] (out_size: int = square, or (out_h, out_w)): horizontal pass to uint8, then
    vertical pass to uint8 (ImagingResample)."""
    H, W, _ = img.shape
    out_h, out_w = (out_size, out_size) if isinstance(out_size, int) else out_size

    def one_pass(src: np.ndarray, in_size: int, n_out: int) -> np.ndarray:      # resample along axis 1
        bounds, kk = pil_coeffs(in_size, n_out)
        out = np.empty((src.shape[0], n_out, src.shape[2]), np.uint8)
        s = src.astype(np.int64)
        for xx in range(n_out):
            xmin, n = bounds[xx]
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(s[:, xmin:xmin + n, :], kk[xx, :n], axes=([1], [0]))
            out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
        return out

    tmp = one_pass(img, W, out_w) if W != out_w else img
    if H != out_h:
        tmp = one_pass(tmp.transpose(1, 0, 2), H, out_h).transpose(1, 0, 2)
    return np.ascontiguousarray(tmp)


def transform_restated(img: np.ndarray, use_crop: bool = False) -> np.ndarray:
    """The reference `_transform` (label_reward.py:92-121) without PIL: fp32 [3,224,224]. Non-square frames: Resize(224)
    takes the shorter side to 224 and the longer to int(224 * long / short); CenterCrop(224) (no-crop route only) takes
    the window at int(round((dim - 224) / 2.0))."""
    if use_crop:
        cs = img.shape[-2] // 2                                     # CenterCrop(image_size // 2), :96,:104
        top = int(round((img.shape[0] - cs) / 2.0))
        left = int(round((img.shape[1] - cs) / 2.0))
        img = img[top:top + cs, left:left + cs]
    H, W = img.shape[:2]
    out_h, out_w = (224, int(224 * W / H)) if W >= H else (int(224 * H / W), 224)
    u8 = pil_resize_bicubic(img, (out_h, out_w))
    if not use_crop:
        top, left = int(round((out_h - 224) / 2.0)), int(round((out_w - 224) / 2.0))
        u8 = u8[top:top + 224, left:left + 224]
    x = torch.from_numpy(np.ascontiguousarray(u8)).permute(2, 0, 1).to(torch.float32).div(255)      # ToTensor
    mean = torch.tensor(CLIP_MEAN, dtype=torch.float32).view(3, 1, 1)
    std = torch.tensor(CLIP_STD, dtype=torch.float32).view(3, 1, 1)
    return x.sub(mean).div(std).numpy()                                        # Normalize


def transform_pil(use_crop: bool, image_size: int | None = None):
    """The reference `_transform` verbatim in behaviour: the same torchvision / PIL calls (:92-121)."""
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToPILImage, ToTensor
    steps = [ToPILImage()]
    if use_crop:
        steps.append(CenterCrop(image_size // 2))
    steps.append(Resize(224, interpolation=InterpolationMode.BICUBIC))
    if not use_crop:
        steps.append(CenterCrop(224))
    steps += [lambda im: im.convert("RGB"), ToTensor(), Normalize(CLIP_MEAN, CLIP_STD)]
    return Compose(steps)


# ---------------------------------------------------------------------------------------------------
# a3': adapter preprocess                                 finetune_module/clip_multiscale_adapter.py:121-133
# ---------------------------------------------------------------------------------------------------
def preprocess_bilinear(x_u8: torch.Tensor) -> torch.Tensor:
    """[B,H,W,3] uint8 -> [B,3,224,224] fp32: float bilinear resize WITHOUT antialias (torchvision 0.12,
    requirements.txt:130), skipped if H or W is already 224, then /255 and normalise."""
    B, H, W, C = x_u8.shape
    assert C == 3
    x = x_u8.permute(0, 3, 1, 2).float()
    if H != 224 and W != 224:
        x = F.interpolate(x, size=(224, 224), mode="bilinear", align_corners=False, antialias=False)
    x = x / 255.0
    mean = torch.tensor(CLIP_MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, dtype=torch.float32).view(1, 3, 1, 1)
    return (x - mean) / std


def center_crop_np(images: np.ndarray, crop_size) -> np.ndarray:
    """label_reward.py:15-36."""
    _, H, W, _ = images.shape
    ch, cw = crop_size
    top, left = int((H - ch) / 2), int((W - cw) / 2)
    return images[:, top:top + ch, left:left + cw, :]


# ---------------------------------------------------------------------------------------------------
# a8 / a9: return-to-go scan and window re-stack                                label_reward.py:232-254
# ---------------------------------------------------------------------------------------------------
def discount_cumsum(x: np.ndarray, gamma: float = 1.0) -> np.ndarray:
    if x.ndim == 0:
        x = x[None]
    out = np.zeros_like(x)
    out[-1] = x[-1]
    for t in reversed(range(x.shape[0] - 1)):
        out[t] = x[t] + gamma * out[t + 1]
    return out


def stack_outputs(x: np.ndarray, num_frames: int) -> np.ndarray:
    if x.ndim == 0:
        x = x[None]
    win: deque = deque([], maxlen=num_frames)
    rows = []
    for i in range(len(x)):
        if i == 0:
            win.extend([x[i]] * num_frames)
        else:
            win.append(x[i])
        rows.append(list(win))
    return np.asarray(rows)


def stack_outputs_fast(x: np.ndarray, num_frames: int) -> np.ndarray:
    """Same result as stack_outputs by index arithmetic: out[i,f] = x[max(0, i-(F-1-f))] (a pure gather)."""
    T = len(x)
    idx = np.maximum(0, np.arange(T)[:, None] - (num_frames - 1 - np.arange(num_frames))[None, :])
    return x[idx]


# ---------------------------------------------------------------------------------------------------
# a5 / a7: compute_reward for model_type="clip"                                 label_reward.py:132-146
# ---------------------------------------------------------------------------------------------------
@torch.no_grad()
def compute_reward_clip(model, images: np.ndarray, text, use_crop: bool = False, reduce: str = "first",
                        preprocess: str = "pil") -> np.ndarray:
    """images uint8 [T,H,W,3]. reduce="first" is what the reference does for str AND list text (SURVEY.md Q1:
    `text` is rebound to the token tensor before the isinstance test); "mean" is envs/vl_reward.py:19-22."""
    if preprocess == "pil":
        tf = transform_pil(use_crop, images.shape[-2])
        x = torch.from_numpy(np.stack([tf(img).numpy() for img in images]))
    else:
        x = torch.from_numpy(np.stack([transform_restated(img, use_crop) for img in images]))
    tokens = clip_shim.tokenize(text if isinstance(text, list) else [text])
    _, logits_per_text = model(x, tokens)
    r = logits_per_text.mean(dim=0) if reduce == "mean" else logits_per_text[0]
    return r.cpu().numpy()


@torch.no_grad()
def compute_reward_clip_goal(model, images: np.ndarray, use_crop: bool = False) -> np.ndarray:
    """model_type="clip_goal_conditioned" (label_reward.py:152-163): -||f_t - f_T||_2, float64 via .item()."""
    tf = transform_pil(use_crop, images.shape[-2])
    x = torch.from_numpy(np.stack([tf(img).numpy() for img in images]))
    feats = model.encode_image(x)
    goal = feats[-1]
    return -1 * np.array([torch.norm(f - goal, p=2).item() for f in feats])


# ---------------------------------------------------------------------------------------------------
# a6: CLIPMultiscaleAdapter inference, restated functionally         clip_multiscale_adapter.py:135-177
# ---------------------------------------------------------------------------------------------------
class AdapterOracle:
    """encode_image / encode_text of the reference adapter from its state_dict, without importing the
    reference: forward hooks on every resblock of the shim CLIP collect the CLS / EOT rows
    (finetune_module/utils.py:6-18), exactly the tensors the reference's hooks capture."""

    def __init__(self, state_dict: dict, arch: str = "ViT-B/16"):
        self.sd = {k: v.float() for k, v in state_dict.items() if torch.is_tensor(v)}
        self.clip = clip_shim.build(arch, seed=0)
        clip_sd = {k[len("clip_model."):]: v for k, v in self.sd.items() if k.startswith("clip_model.")}
        self.clip.load_state_dict(clip_sd, strict=True)
        self.clip = self.clip.float().eval()
        self.num_clip_layers = self.clip.transformer.layers
        self.acts: dict[str, torch.Tensor] = {}
        for name, mod in self.clip.named_modules():
            if name.startswith(("visual.transformer.resblocks.", "transformer.resblocks.")) and name.split(".")[-1].isdigit():
                mod.register_forward_hook(lambda m, i, o, name=name: self.acts.__setitem__(name, o.detach()))
        self.logit_scale = self.clip.logit_scale.detach().clone()   # plain attribute, never in the checkpoint (:102)

    def _mlp(self, side: str, x):
        h = torch.relu(F.linear(x, self.sd[f"{side}_adapter.layers.0.weight"], self.sd[f"{side}_adapter.layers.0.bias"]))
        return F.linear(h, self.sd[f"{side}_adapter.layers.3.weight"], self.sd[f"{side}_adapter.layers.3.bias"])

    @torch.no_grad()
    def encode_image(self, x: torch.Tensor, normalize: str = "all") -> torch.Tensor:
        f = self.clip.encode_image(x).float()
        taps = [self.acts[f"visual.transformer.resblocks.{l}"].permute(1, 0, 2)[:, 0, :].float()
                for l in range(self.clip.visual.transformer.layers)]
        inter = F.linear(torch.cat(taps, dim=-1), self.sd["image_intermediate_linear.weight"])
        feat = torch.cat([inter, f], dim=-1)
        res = torch.sigmoid(self.sd["image_residual_weight"])
        a = res * feat + (1.0 - res) * self._mlp("image", feat)
        if normalize == "per_scale":
            return F.normalize(a.view(a.shape[0], self.num_clip_layers + 1, -1), dim=-1)      # 3-D features
        return F.normalize(a, dim=-1)

    @torch.no_grad()
    def encode_text(self, tokens: torch.Tensor, normalize: str = "all") -> torch.Tensor:
        f = self.clip.encode_text(tokens).float()
        n = tokens.shape[0]
        eot = tokens.argmax(dim=-1)
        taps = [self.acts[f"transformer.resblocks.{l}"].permute(1, 0, 2)[torch.arange(n), eot].float()
                for l in range(self.num_clip_layers)]
        inter = F.linear(torch.cat(taps, dim=-1), self.sd["text_intermediate_linear.weight"])
        feat = torch.cat([inter, f], dim=-1)
        res = torch.sigmoid(self.sd["text_residual_weight"])
        a = res * feat + (1.0 - res) * self._mlp("text", feat)
        if normalize == "per_scale":
            return F.normalize(a.view(n, self.num_clip_layers + 1, -1), dim=-1)
        return F.normalize(a, dim=-1)

    @torch.no_grad()
    def compute_reward(self, images: np.ndarray, text, use_crop: bool = False, ensemble: bool = False,
                       reduce: str = "first") -> np.ndarray:
        """label_reward.py:200-230."""
        if use_crop:
            s = images.shape[-2]
            images = center_crop_np(images, (s // 2, s // 2))
        tokens = clip_shim.tokenize(text if isinstance(text, list) else [text])
        mode = "per_scale" if ensemble else "all"
        fi = self.encode_image(preprocess_bilinear(torch.from_numpy(np.ascontiguousarray(images))), mode)
        ft = self.encode_text(tokens, mode)
        scale = self.logit_scale.exp()
        if fi.dim() == 3:                                            # :217-222
            fi = fi.reshape(len(images), -1)
            ft = ft.reshape(tokens.shape[0], -1)
            logit = (scale * (fi @ ft.T)).t() / (self.num_clip_layers + 1)
        else:
            logit = (scale * (fi @ ft.T)).t()
        r = logit.mean(dim=0) if reduce == "mean" else logit[0]
        return r.float().cpu().numpy()


def _adapter_compute_reward_goal(self, images: np.ndarray, use_crop: bool = False) -> np.ndarray:
    """label_reward.py:180-196 (`clip_*_goal_conditioned`): distance of every frame's adapted feature to the episode's
    LAST frame — POSITIVE here (no `-1 *`, unlike the plain-CLIP branch :160), float64 through `.item()`."""
    if use_crop:
        s = images.shape[-2]
        images = center_crop_np(images, (s // 2, s // 2))
    f = self.encode_image(preprocess_bilinear(torch.from_numpy(np.ascontiguousarray(images))))
    return np.array([torch.norm(x - f[-1], p=2).item() for x in f])


AdapterOracle.compute_reward_goal = _adapter_compute_reward_goal


# ---------------------------------------------------------------------------------------------------
# (f)4: adapter fine-tuning forward (value only)              clip_multiscale_adapter.py:179-252
# ---------------------------------------------------------------------------------------------------
@torch.no_grad()
def finetune_loss(adapter: "AdapterOracle", batch: dict, goal_conditioned: bool = False, use_vip_loss: bool = True,
                  use_id_loss: bool = True, gamma: float = 0.98) -> float:
    """CLIPMultiscaleAdapter.forward(batch) restated on the AdapterOracle (augmentation = identity). `adapter.sd` must
    hold `inverse_layer.layers.{0,3}.*` and `lambda_id`; discrete actions (CrossEntropyLoss, finetune.py:139)."""
    sd = adapter.sd
    total = 0.0
    for key in batch["image1"].keys():
        imgs = [preprocess_bilinear(torch.as_tensor(batch[f"image{i}"][key])) for i in range(4)]     # :199-203
        a0, a1, a2 = (adapter.encode_image(x) for x in imgs[:3])                                     # :207-211
        if goal_conditioned:                                                                          # :214-218
            a3 = adapter.encode_image(imgs[3])
            s0, s1, s2 = (-torch.linalg.norm(a3 - a, dim=-1) for a in (a0, a1, a2))
            other = a3
        else:                                                                                         # :220-224
            t = adapter.encode_text(torch.as_tensor(batch["instruct"]))
            scale = adapter.logit_scale.exp()
            s0, s1, s2 = (torch.diag(scale * (a @ t.T), 0) for a in (a0, a1, a2))
            other = t
        r = torch.as_tensor(batch["r"]).float() - 1                                                   # :227
        vip = (1 - gamma) * -s0.mean() + torch.log(1e-8 + torch.mean(torch.exp(-(r + gamma * s2 - s1))))
        x = torch.cat([a1, other, a2, other], dim=-1)                                                 # :233-244
        h = torch.relu(F.linear(x, sd["inverse_layer.layers.0.weight"], sd["inverse_layer.layers.0.bias"]))
        logits = F.linear(h, sd["inverse_layer.layers.3.weight"], sd["inverse_layer.layers.3.bias"])
        id_loss = F.cross_entropy(logits, torch.as_tensor(batch["action"]).long())
        if use_vip_loss:
            total = total + vip
        if use_id_loss:
            total = total + sd["lambda_id"] * id_loss
    return float(total)


# ---------------------------------------------------------------------------------------------------
# (f)1: online reward of a rollout                    arp_dt/envs/vl_reward.py:11-77, rollout_procgen.py:133-150
# ---------------------------------------------------------------------------------------------------
def _online_image(obs: np.ndarray) -> torch.Tensor:
    """preprocess(Image.fromarray(obs)).unsqueeze(0) with clip.load's preprocess (= _transform(224))."""
    return transform_pil(False)(np.asarray(obs)).unsqueeze(0)


def _online_crop(obs: np.ndarray) -> np.ndarray:
    """center_crop(obs[None], (obs.shape[0] // 2, obs.shape[0] // 2))[0]                      (vl_reward.py:13-14)"""
    return center_crop_np(obs[None, ...], (obs.shape[0] // 2, obs.shape[0] // 2))[0]


@torch.no_grad()
def online_clip_reward(model, obs: np.ndarray, pos_text, use_crop: bool = False) -> np.ndarray:
    """get_torch_clip_reward (vl_reward.py:11-23): list text -> mean over texts, str -> row 0; float32 [1]."""
    if use_crop:
        obs = _online_crop(obs)
    tokens = clip_shim.tokenize(pos_text)
    _, logits_per_text = model(_online_image(obs), tokens)
    r = logits_per_text.mean(axis=0) if isinstance(pos_text, list) else logits_per_text[0]
    return r.float().cpu().numpy()


@torch.no_grad()
def online_goal_reward(encode_image, obs: np.ndarray, goal_image: np.ndarray, use_crop: bool = False) -> float:
    """get_torch_clip_goal_conditioned_reward / ..._adapter_goal_conditioned_reward (vl_reward.py:26-41, :62-77).
    Quirk kept: with use_crop the goal is cropped with the ALREADY CROPPED obs' size, i.e. to H/4 (:29-30)."""
    if use_crop:
        obs = _online_crop(obs)
        goal_image = center_crop_np(goal_image[None, ...], (obs.shape[0] // 2, obs.shape[0] // 2))[0]
    f, g = encode_image(_online_image(obs)), encode_image(_online_image(goal_image))
    return -1 * torch.norm(f - g).item()


@torch.no_grad()
def online_adapter_reward(adapter: "AdapterOracle", obs: np.ndarray, pos_text, use_crop: bool = False) -> np.ndarray:
    """get_torch_clip_adapter_reward (vl_reward.py:44-59): CLIP's PIL preprocess, the adapter's encoders."""
    if use_crop:
        obs = _online_crop(obs)
    tokens = clip_shim.tokenize(pos_text)
    fi, ft = adapter.encode_image(_online_image(obs)), adapter.encode_text(tokens)
    logit = (adapter.logit_scale.exp() * (fi @ ft.T)).t()
    r = logit.mean(axis=0) if isinstance(pos_text, list) else logit[0]
    return r.float().cpu().numpy()


def update_rtg(rtg: float, clip_reward, scale: float, reward_min: float = 0.0, use_normalize: bool = False):
    """rollout_procgen.py:147-150: the return-to-go token the policy is conditioned on at the next step."""
    if use_normalize:
        return rtg - (clip_reward - reward_min) / scale
    return rtg - clip_reward / scale


# ---------------------------------------------------------------------------------------------------
# (f)2: consumer side                       arp_dt/data_procgen.py:132-174 (preprocess_rtgs), utils.py:453-463
# ---------------------------------------------------------------------------------------------------
def compute_scale(return_to_go):
    """arp_dt/utils.py:453-463."""
    if return_to_go >= 0:
        max_digit = int(str(int(return_to_go))[0])
    else:
        max_digit = int(str(int(return_to_go))[1])
    if return_to_go >= 0:
        n = len(str(int(return_to_go))) - 1 if max_digit < 5 else len(str(int(return_to_go)))
    else:
        n = len(str(int(return_to_go))) - 2 if max_digit < 5 else len(str(int(return_to_go))) - 1
    return pow(10, n)


def preprocess_rtgs(rewards: dict, traj_idx, num_frames: int, env_name: str, use_normalize: bool) -> dict:
    """ProcgenDataset.preprocess_rtgs (data_procgen.py:132-174) with the h5 reads factored out:
    rewards[image_key] = h5_file[f"{image_key}_{vl_type}_pos_reward"][:, -1].astype(np.float32) (:142-146)."""
    reward_min = {k: np.min(r) for k, r in rewards.items()}                              # :147
    reward_max = {k: np.max(r) for k, r in rewards.items()}                              # :148
    modified = {k: (r - reward_min[k]) for k, r in rewards.items()} if use_normalize else rewards   # :150-153
    out = {k: [] for k in rewards}
    for k in out:                                                                         # :155-168
        for idx in range(len(traj_idx) - 1):
            stack = deque([], maxlen=num_frames)
            ids = list(range(traj_idx[idx], traj_idx[idx + 1]))
            cs = discount_cumsum(modified[k][ids], gamma=1.0)
            for i in range(len(ids)):
                if i == 0:
                    stack.extend([cs[i]] * num_frames)
                else:
                    stack.append(cs[i])
                out[k].append(np.stack(stack))
    if "coinrun" in env_name:                                                             # :171-174
        rtg = np.max(list(out.values())) // 100 * 100
    else:
        rtg = np.quantile(list(out.values()), 0.9) // 100 * 100
    return {"rtgs": {k: np.asarray(v) for k, v in out.items()}, "reward_min": reward_min, "reward_max": reward_max,
            "return_to_go": rtg, "scale": compute_scale(rtg)}


# ---------------------------------------------------------------------------------------------------
# the whole labeler                                                          label_reward.py:44-291
# ---------------------------------------------------------------------------------------------------
def label_reward_port(data: dict, *, model=None, adapter: AdapterOracle | None = None, model_type: str = "clip",
                      text="the goal is to collect the coin.", use_crop: bool = False, reduce: str = "first",
                      image_key: str = "ob", inst_type: str = "none", preprocess: str = "pil",
                      max_episodes: int | None = None) -> dict:
    """data: {"ob": u8 [T,F,H,W,3], "done": f32 [T,F]}. Returns the two datasets the reference writes
    (plus the per-frame "reward"/"rtg" vectors and the episode index) keyed exactly like the reference."""
    done = data["done"]
    len_data, num_frames = done.shape[:2]
    idx = episode_index(done[:, -1])
    n_eps = len(idx) - 1 if max_episodes is None else min(len(idx) - 1, max_episodes)
    rs, gs, rv, gv = [], [], [], []
    for e in range(n_eps):
        traj = list(range(idx[e], min(idx[e + 1], len_data)))
        imgs = data[image_key][traj, -1]
        if model_type == "clip":
            r = compute_reward_clip(model, imgs, text, use_crop, reduce, preprocess)
        elif model_type == "clip_goal_conditioned":
            r = compute_reward_clip_goal(model, imgs, use_crop)
        elif "_goal_conditioned" in model_type:
            r = adapter.compute_reward_goal(imgs, use_crop)
        else:
            r = adapter.compute_reward(imgs, text, use_crop, ensemble="ensemble" in model_type, reduce=reduce)
        g = discount_cumsum(r)
        rv.append(r)
        gv.append(g)
        rs.append(stack_outputs(r, num_frames))
        gs.append(stack_outputs(g, num_frames))
    suffix = "" if inst_type == "none" else f"_{inst_type}"
    cat = lambda parts, shape: np.concatenate(parts) if parts else np.zeros(shape, np.float32)  # noqa: E731
    return {
        f"{image_key}_{model_type}_reward{suffix}": cat(rs, (0, num_frames)),
        f"{image_key}_{model_type}_pos_rtg{suffix}": cat(gs, (0, num_frames)),
        "reward": cat(rv, (0,)), "rtg": cat(gv, (0,)), "g_traj_idx": idx, "frames": sum(len(r) for r in rv),
    }
