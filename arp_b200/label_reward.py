"""Drop-in for the reference's offline reward labeler, `arp_dt/label_reward.py`.

Same entry point, same arguments and defaults (label_reward.py:44-60), same CLI flags (:295-312), same
`model_type` names and checkpoint format, same output datasets
    f"{img_key}_{model_type}_reward[_{inst_type}]",  f"{img_key}_{model_type}_pos_rtg[_{inst_type}]"
as float32 [T, num_frames] (:257-289) — but every frame is scored by the native sm_100a library
(include/arp_b200.h) instead of a Python loop over PyTorch calls:

    reference (per episode, :265-271)                      here (per slab of many episodes)
    -------------------------------------------------      ----------------------------------------------
    g[img_key][traj, -1]  (host gather)                    strided host pointer -> cudaMemcpy2DAsync
    [preprocess(img) for img in imgs]  (PIL, 1 thread)     fused decode kernel (Pillow-exact)
    clip.tokenize + text tower, every episode              text embedding cached once per run
    model(images, text)                                    tcgen05 ViT + fused cosine head
    discount_cumsum / stack_outputs  (python loops)        per-episode scan + window stack kernel

There is no CPU fallback: without the built library and a B200 this module raises.

Keyword-only extensions (all optional; defaults reproduce the reference):
    clip_state_dict  CLIP weights (openai/CLIP state_dict keys). The reference downloads pretrained
                     weights inside clip.load (:126); offline, pass them here or set ARP_CLIP_CHECKPOINT.
    arch             "ViT-B/16" (what the reference hard-codes) or "ViT-B/32".
    reduce           "first" (reference behaviour) or "mean" (envs/vl_reward.py semantics).
    max_batch        frames per device chunk.      slab_frames  frames per host slab.
    device           CUDA ordinal (default: LOCAL_RANK or 0).
    precision        "16bit" (default; "bf16"/"fp16" are accepted aliases): the tensor-core product path — 16-bit operands
                     in the library's operand format (fp16 by default, the reference's own CUDA format), 16-bit residual
                     stream, LayerNorm folded into the GEMMs. "fp32resid": same tensor-core kernels with an fp32 residual
                     stream and standalone LayerNorm (~2x closer to fp32, ~10 % slower). "fp32": verification path, every
                     weight, activation and contraction in fp32 (what the reference's CPU route computes; ~50x slower).
    distributed      shard episodes over torch.distributed ranks (default: on when a process group exists).
    tokenizer        None: openai/CLIP's clip.tokenize (label_reward.py:136) — if the package is missing the call REFUSES
                     unless ARP_ALLOW_STANDIN_TOKENIZER=1; "standin": the deterministic stand-in (random-init experiments
                     only); or a callable(list[str]) -> int tensor [n, 77].
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from . import capi
from .instructions import get_clip_instruct, get_clip_special_instruct
from .sharding import gather_rows, partition_episodes
from .store import open_store
from .text_tower import adapter_text_embedding, clip_text_embedding
from .tokenizer import resolve as resolve_tokenizer
from .weights import load_checkpoint


def center_crop(image, crop_size):
    """label_reward.py:15-36 (imported by envs/vl_reward.py:6 and envs/rollout_procgen.py:14): [N,H,W,C] centre crop."""
    _, H, W, _ = image.shape
    ch, cw = crop_size
    top, left = int((H - ch) / 2), int((W - cw) / 2)
    return image[:, top:top + ch, left:left + cw, :]


def episode_index(g, done_key=None):
    """label_reward.py:71-87 — (len_data, num_frames, g_traj_idx)."""
    if done_key is None:
        for cand in ("done", "rewards", "is_terminal"):
            if g.get(cand):
                done_key = cand
                break
        else:
            raise ValueError
    try:
        len_data, num_frames = g[done_key].shape[:2]
        idx = list(np.nonzero(np.asarray(g[done_key][:, -1]))[0] + 1)
        idx.insert(0, 0)
    except Exception:  # noqa: BLE001 — the reference's bare `except:` fallback for the "time" layout (:84-87)
        len_data, num_frames = g["time"].shape[:2]
        idx = list(np.where(np.asarray(g["time"][:, -1, 0]) == 1.0)[0])
        idx.append(len(g["time"]))
    return int(len_data), int(num_frames), [int(i) for i in idx]


def _head_for(model_type: str) -> tuple[int, int]:
    """model_type -> (ArpHead, ArpPreprocess), following the dispatch at label_reward.py:123-230."""
    if model_type == "clip":
        return capi.HEAD_CLIP, capi.PRE_PIL_BICUBIC
    if model_type == "clip_goal_conditioned":
        return capi.HEAD_CLIP_GOAL, capi.PRE_PIL_BICUBIC
    if model_type.startswith("clip_"):
        # Only "clip_ft" constructs a model in the reference (:166-173); other clip_* names crash there
        # (SURVEY.md Q3). They are served by the same CLIPMultiscaleAdapter weights.
        if "_goal_conditioned" in model_type:
            return capi.HEAD_ADAPTER_GOAL, capi.PRE_BILINEAR
        if "ensemble" in model_type:
            return capi.HEAD_ADAPTER_ENSEMBLE, capi.PRE_BILINEAR
        return capi.HEAD_ADAPTER, capi.PRE_BILINEAR
    raise ValueError(f"unsupported model_type {model_type!r}: the labeler only defines clip* reward models")


def _resolve_clip_weights(clip_state_dict, arch: str) -> dict:
    if clip_state_dict is not None:
        return clip_state_dict
    path = os.environ.get("ARP_CLIP_CHECKPOINT")
    if path:
        return load_checkpoint(path)
    try:  # the reference's own route, if the real package and its cached weights exist
        import clip  # type: ignore
        model, _ = clip.load(arch, device="cpu")
        return model.state_dict()
    except Exception as e:  # noqa: BLE001
        raise RuntimeError(
            "CLIP weights unavailable: pass clip_state_dict=..., set ARP_CLIP_CHECKPOINT to a state_dict file, "
            "or install openai/CLIP with its cached checkpoint (the reference downloads it in clip.load, "
            "label_reward.py:126)") from e


PRECISIONS = {"16bit": capi.PREC_16BIT, "bf16": capi.PREC_16BIT, "fp16": capi.PREC_16BIT,
              "fp32resid": capi.PREC_F32RESID, "fp32": capi.PREC_F32}


# One native handle is kept alive between label_reward() calls of a process (key = its whole configuration): creating it
# (4 GB of workspace, the pinned staging ring) and tearing it down cost ~0.25 s, a tenth of a 70k-frame labeling pass.
# Weights and the instruction embedding are uploaded again on every call. release_cached_engine() frees it.
_ENGINE_CACHE: dict = {}


def release_cached_engine():
    for e in _ENGINE_CACHE.values():
        e.close()
    _ENGINE_CACHE.clear()


def _cached_engine(**cfg):
    key = tuple(sorted(cfg.items()))
    e = _ENGINE_CACHE.get(key)
    if e is None:
        release_cached_engine()                      # a single handle: workspaces of different configurations do not add up
        e = capi.Engine(**cfg)
        _ENGINE_CACHE[key] = e
    return e


class RewardLabeler:
    """Model + cached instruction embedding on one GPU; label() scores slabs of episodes.

    This is the object form of the closures `compute_reward` / `discount_cumsum` / `stack_outputs`
    that the reference builds inside label_reward() (:132-254)."""

    def __init__(self, model_type: str, text, frame_hw: tuple[int, int], *, model_ckpt_dir=None,
                 clip_state_dict=None, arch: str = "ViT-B/16", use_crop: bool = False, reduce: str = "first",
                 max_batch: int = 1024, device: int | None = None, precision: str = "16bit", tokenizer=None):
        head, pre = _head_for(model_type)
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
        adapter = head in (capi.HEAD_ADAPTER, capi.HEAD_ADAPTER_ENSEMBLE, capi.HEAD_ADAPTER_GOAL)
        if adapter:
            assert model_ckpt_dir is not None, "specify model_ckpt_dir"  # label_reward.py:174
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.model_type, self.device_index = model_type, device
        patch = 32 if arch.endswith("/32") else 16
        self.engine = _cached_engine(device=device, patch=patch, in_h=int(frame_hw[0]), in_w=int(frame_hw[1]),
                                     use_crop=bool(use_crop), preprocess=pre, head=head,
                                     reduce=capi.REDUCE_MEAN if reduce == "mean" else capi.REDUCE_FIRST,
                                     max_batch=int(max_batch), precision=PRECISIONS[precision])
        dev = self.engine.device
        if adapter:
            sd = load_checkpoint(model_ckpt_dir) if not isinstance(model_ckpt_dir, dict) else model_ckpt_dir
            if not any(k.startswith("clip_model.") for k in sd):
                # strict=False load in the reference (:176): CLIP tensors missing from the checkpoint keep
                # the values clip.load gave them.
                base = _resolve_clip_weights(clip_state_dict, arch)
                sd = {**{"clip_model." + k: v for k, v in base.items()}, **sd}
        else:
            sd = _resolve_clip_weights(clip_state_dict, arch)
        missing = self.engine.load_state_dict(sd, strict=False)
        if missing:
            raise RuntimeError(f"checkpoint lacks {len(missing)} tensors the {model_type} path needs, e.g. {missing[:3]}")
        self.goal = self.engine.goal
        if not self.goal:
            texts = list(text) if isinstance(text, (list, tuple)) else [text]
            tokens = resolve_tokenizer(tokenizer)(texts)
            # the instruction embedding of the previous call is reused when the very same weight tensors (objects and
            # in-place version counters) and token ids come back — the cached handle keeps them referenced
            text_sd = {k: v for k, v in sd.items() if torch.is_tensor(v) and "visual." not in k}
            sig = (tokens.tolist(), head, [(k, v._version) for k, v in text_sd.items()])
            cached = getattr(self.engine, "_text_cache", None)
            if cached is not None and cached[0] == sig and all(cached[1].get(k) is v for k, v in text_sd.items()):
                emb, scale = cached[2]
            else:
                if adapter:
                    emb, scale = adapter_text_embedding(sd, tokens, dev, ensemble=head == capi.HEAD_ADAPTER_ENSEMBLE)
                else:
                    emb, scale = clip_text_embedding(sd, tokens, dev)
                self.engine._text_cache = (sig, text_sd, (emb, scale))
            self.engine.set_text(emb, scale)

    def label_slab(self, ob: np.ndarray, ep_offsets: np.ndarray, num_frames: int):
        """ob: host uint8 [T,F,H,W,3] (or [T,H,W,3]); ep_offsets: int64 [n+1] relative to the slab.
        Returns host (reward[T], rtg[T], reward_stacked[T,F], rtg_stacked[T,F])."""
        return self.engine.label_host(ob, ep_offsets, num_frames)

    def label_file(self, fd: int, file_offset: int, T: int, row_stride_bytes: int, ep_offsets: np.ndarray, num_frames: int):
        """Same, the scored frame of row t read by the library from `fd` at file_offset + t*row_stride_bytes."""
        return self.engine.label_file(fd, file_offset, T, row_stride_bytes, ep_offsets, num_frames)

    def close(self, release: bool = False):
        """The native handle stays cached for the next call of this process unless release=True."""
        if release:
            release_cached_engine()


def _slabs(ep_offsets: np.ndarray, e_lo: int, e_hi: int, slab_frames: int):
    """Group whole episodes [e_lo, e_hi) into slabs of at most ~slab_frames frames (at least one episode)."""
    e = e_lo
    while e < e_hi:
        j = e + 1
        while j < e_hi and ep_offsets[j + 1] - ep_offsets[e] <= slab_frames:
            j += 1
        yield e, j
        e = j


SIDECAR_SUFFIX = "_last"   # "<img_key>_last": uint8 [T,H,W,3], the scored (last stacked) frame of every row


def write_destacked_sidecar(store, img_key: str = "ob", rows_per_pass: int = 4096) -> str:
    """the recorder stores every row as F stacked frames in gzip chunks (1,F,H,W,3)
    (data/PPG/trajectory_recorder.py:154-162), so scoring the LAST frame of a row inflates F times more bytes than it
    uses (label_reward.py:268). This writes the scored frames once as "<img_key>_last" [T,H,W,3]; label_reward() reads
    the sidecar when it exists. Returns the sidecar key."""
    ds = store[img_key]
    T = ds.shape[0]
    key = img_key + SIDECAR_SUFFIX
    if store.get(key) is not None:
        return key
    out = store.create_dataset(key, shape=(T,) + tuple(ds.shape[2:]), dtype=np.uint8, chunks=(1,) + tuple(ds.shape[2:]),
                               compression="gzip")
    for lo in range(0, T, rows_per_pass):
        hi = min(T, lo + rows_per_pass)
        out[lo:hi] = np.asarray(ds[lo:hi, -1])
    return key


def _rows_array(ds, lo: int, hi: int, sidecar=None) -> np.ndarray:
    """Rows [lo, hi) of an image dataset as a C-contiguous host array WITHOUT touching the frames that
    are not scored when the container allows it (memory-mapped store); h5py reads only `[:, -1]`; a de-stacked
    sidecar ([T,H,W,3], see write_destacked_sidecar) is read instead of the stacked dataset when present."""
    if sidecar is not None:
        arr = getattr(sidecar, "array", None)
        if arr is not None and arr.flags.c_contiguous:
            return arr[lo:hi]
        return np.ascontiguousarray(sidecar[lo:hi])
    arr = getattr(ds, "array", None)
    if arr is not None and arr.flags.c_contiguous:
        return arr[lo:hi]                       # strided pointer goes straight to arp_label_host
    return np.ascontiguousarray(ds[lo:hi, -1])  # label_reward.py:268 — last stacked frame only


def label_reward(
    env_name,
    distribution_mode,
    num_levels,
    start_level,
    text,
    base_path,
    data_path=None,
    image_keys="ob",
    num_demonstrations=500,
    num_frames=8,
    env_type=None,
    model_type="clip",
    model_ckpt_dir=None,
    use_crop=False,
    inst_type="none",
    *,
    clip_state_dict=None,
    arch="ViT-B/16",
    reduce="first",
    max_batch=1024,
    slab_frames=16384,
    device=None,
    distributed=None,
    precision="16bit",
    tokenizer=None,
):
    image_keys = image_keys.split(", ")
    if data_path is None:
        dirname = (
            f"{env_name}_{distribution_mode}_level{start_level}to{num_levels}_num{num_demonstrations}_frame{num_frames}"
        )
        if env_type != "none":
            dirname += f"_{env_type}"
        data_path = os.path.join(base_path, dirname, "data.hdf5")

    import torch.distributed as dist
    if distributed is None:
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    rank = dist.get_rank() if distributed else 0
    world = dist.get_world_size() if distributed else 1

    import time
    _t = [("start", time.perf_counter())]
    _mark = lambda name: _t.append((name, time.perf_counter()))  # noqa: E731 — ARP_TIMING=1 prints the phase breakdown
    # Every rank READS through its own read-only handle; rank 0 reopens the container for writing only after all
    # readers have closed (HDF5 file locking refuses a writer next to readers, and a reader next to the writer).
    g = open_store(data_path, "r")
    labeler = None
    failure = None
    results = {}
    try:
        len_data, num_frames, g_traj_idx = episode_index(g)  # num_frames comes from the file (Q4)
        n_eps = len(g_traj_idx) - 1
        off = np.minimum(np.asarray(g_traj_idx, dtype=np.int64), len_data)  # min(idx[i+1], len_data) (:267)
        shards = partition_episodes(off, world)
        e_lo, e_hi = shards[rank]
        target_keys = [f"{model_type}_reward", f"{model_type}_pos_rtg"]
        if inst_type != "none":
            target_keys = [f"{x}_{inst_type}" for x in target_keys]
        _mark("open + episode index")
        try:
            H, W = g[image_keys[0]].shape[-3:-1]
            if use_crop:
                print(f"image_size: {g[image_keys[0]].shape[-2]}")  # label_reward.py:104-105
            labeler = RewardLabeler(model_type, text, (int(H), int(W)), model_ckpt_dir=model_ckpt_dir,
                                    clip_state_dict=clip_state_dict, arch=arch, use_crop=use_crop, reduce=reduce,
                                    max_batch=max_batch, device=device, precision=precision, tokenizer=tokenizer)
            _mark("engine + weights + text tower")
            for img_key in image_keys:
                ds = g[img_key]
                side = g.get(img_key + SIDECAR_SUFFIX)
                if side is not None and (side.shape[0] != ds.shape[0] or tuple(side.shape[1:]) != tuple(ds.shape[2:])):
                    side = None                      # stale or foreign sidecar: fall back to the stacked dataset
                results[img_key] = _label_rows(labeler, ds, side, off, e_lo, e_hi, num_frames, slab_frames)
            _mark("label (decode + encoder + head + scan, H2D / D2H)")
        except Exception as e:  # noqa: BLE001 — reported to every rank below, then re-raised
            failure = e
        if distributed:
            # a rank that failed must not leave the others waiting in the gather: agree on success first
            backend = dist.get_backend()
            cdev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
            flag = torch.tensor([0 if failure is None else 1], device=cdev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if int(flag) and failure is None:
                failure = RuntimeError("label_reward failed on another rank; nothing was written")
        if failure is not None:
            release_cached_engine()                  # do not keep a handle that may be in an error state
            raise failure
        if distributed:
            rows = [int(off[b] - off[a]) for a, b in shards]
            for img_key in image_keys:
                rs, gs = results[img_key]
                both = torch.from_numpy(np.stack([rs, gs], axis=1)).to(cdev)  # [n, 2, F]: the path's one collective
                full = gather_rows(both, rows, dst=0)
                if rank == 0:
                    full = full.cpu().numpy()
                    results[img_key] = (np.ascontiguousarray(full[:, 0]), np.ascontiguousarray(full[:, 1]))
    finally:
        if labeler is not None:
            labeler.close()
        g.close()
    _mark("gather + close")
    if distributed:
        dist.barrier()                                   # every reader has closed
    if rank == 0:
        g = open_store(data_path, "a")
        try:
            for img_key in image_keys:
                _write_labels(g, img_key, target_keys, results[img_key], off, n_eps, len_data, num_frames)
        finally:
            g.close()
    _mark("write labels")
    if distributed:
        dist.barrier()
    if os.environ.get("ARP_TIMING", "0") not in ("", "0") and rank == 0:
        print("[arp_b200] label_reward phases: " + ", ".join(f"{b[0]} {b[1] - a[1]:.3f} s" for a, b in zip(_t[:-1], _t[1:]))
              + f"; total {_t[-1][1] - _t[0][1]:.3f} s", flush=True)


def _label_rows(labeler, ds, side, off, e_lo: int, e_hi: int, num_frames: int, slab_frames: int):
    """(reward_stacked, rtg_stacked) for episodes [e_lo, e_hi). A memory-mapped container is handed to the library in ONE
    call (a strided pointer into the file mapping; the native stager overlaps page-cache reads, PCIe and compute). A
    container that has to be read through its API (h5py: chunked, gzip) is read slab by slab on a reader thread, one slab
    ahead of the GPU."""
    empty = np.zeros((0, num_frames), np.float64 if labeler.goal else np.float32)
    lo_all, hi_all = int(off[e_lo]), int(off[e_hi])
    if hi_all <= lo_all:
        return empty, empty

    def finish(r, rs, gs, rel_off):
        return _goal_float64(r, rel_off, num_frames) if labeler.goal else (rs, gs)

    src = side if side is not None else ds
    rel = off[e_lo:e_hi + 1] - lo_all
    fsrc = src.file_source() if hasattr(src, "file_source") and hasattr(labeler, "label_file") else None
    if fsrc is not None:
        # rows contiguous in a file: the library preads the scored frame of every row itself (arp_label_file)
        path, base = fsrc
        frame = int(np.prod(src.shape[-3:]))
        stride = frame * (src.shape[1] if len(src.shape) == 5 else 1)
        first = base + lo_all * stride + (stride - frame)            # last stacked frame of row lo_all
        fd = os.open(path, os.O_RDONLY)
        try:
            r, _, rs, gs = labeler.label_file(fd, first, hi_all - lo_all, stride, rel, num_frames)
        finally:
            os.close(fd)
        return finish(r, rs, gs, rel)
    arr = getattr(src, "array", None)
    if arr is not None and arr.flags.c_contiguous:
        r, _, rs, gs = labeler.label_slab(arr[lo_all:hi_all], rel, num_frames)
        return finish(r, rs, gs, rel)

    import threading
    slabs = [(a, b) for a, b in _slabs(off, e_lo, e_hi, slab_frames) if off[b] > off[a]]
    box = {}

    def read(i):
        a, b = slabs[i]
        try:
            box[i] = _rows_array(ds, int(off[a]), int(off[b]), side)
        except Exception as e:  # noqa: BLE001 — re-raised on the consumer side
            box[i] = e

    parts_r, parts_g = [], []
    th = threading.Thread(target=read, args=(0,))
    th.start()
    for i, (a, b) in enumerate(slabs):
        th.join()
        rows = box.pop(i)
        if isinstance(rows, Exception):
            raise rows
        if i + 1 < len(slabs):
            th = threading.Thread(target=read, args=(i + 1,))
            th.start()
        rel = off[a:b + 1] - off[a]
        r, _, rs, gs = labeler.label_slab(rows, rel, num_frames)
        rs, gs = finish(r, rs, gs, rel)
        parts_r.append(rs)
        parts_g.append(gs)
    return (np.concatenate(parts_r) if parts_r else empty), (np.concatenate(parts_g) if parts_g else empty)


def _goal_float64(r: np.ndarray, ep_off: np.ndarray, num_frames: int):
    """Goal-conditioned rewards are a float64 array in the reference (np.array of .item() values,
    label_reward.py:160-162,193-195), so its discount_cumsum / stack_outputs run in float64 and the
    datasets are float64. The per-frame distances come from the GPU in fp32 (exactly what .item()
    widens); the O(T) float64 scan is done here, sequentially right-to-left like :252-253."""
    r64 = r.astype(np.float64)
    g64 = np.empty_like(r64)
    for lo, hi in zip(ep_off[:-1], ep_off[1:]):
        if hi > lo:
            g64[lo:hi] = np.cumsum(r64[lo:hi][::-1])[::-1]
    F = num_frames
    idx = np.arange(len(r64))
    start = np.repeat(ep_off[:-1], np.diff(ep_off))
    win = np.maximum(start[:, None], idx[:, None] - (F - 1 - np.arange(F))[None, :])
    return r64[win], g64[win]


def _write_labels(g, img_key, target_keys, data, off, n_eps, len_data, num_frames):
    """label_reward.py:260-289. `data` rows are the labeled rows [off[0], off[n_eps]) in order. When the key does not
    exist the reference creates the dataset from the first episode (gzip, chunks (1,F), maxshape (len_data,F)) and
    APPENDS every later episode (:273-286) — one dataset of off[n_eps]-off[0] rows, whatever off[0] is (it is non-zero
    only in the "time" layout, :84-87). When the key exists it assigns in place at the true row indices (:288-289)."""
    if n_eps <= 0:
        return
    first, last = int(off[0]), int(off[n_eps])
    for _key, arr in zip(target_keys, data):
        key = f"{img_key}_{_key}"
        arr = arr[:last - first]
        existing = g.get(key)
        if not existing:
            g.create_dataset(key, compression="gzip", chunks=(1, num_frames), maxshape=(len_data, num_frames), data=arr)
        else:
            existing[first:last] = arr


def main():
    parser = argparse.ArgumentParser(description="Process rollout training arguments.")
    parser.add_argument("--env_name", type=str, default="coinrun")
    parser.add_argument("--env_type", type=str, default="none")
    parser.add_argument("--num_levels", type=int, default=500)
    parser.add_argument("--start_level", type=int, default=0)
    parser.add_argument("--distribution_mode", type=str, default="hard")
    parser.add_argument("--image_keys", type=str, default="ob")
    parser.add_argument("--data_path", type=str, default=None)
    parser.add_argument("--base_path", type=str, default="./demonstrations")
    parser.add_argument("--num_demonstrations", type=int, default=500)
    parser.add_argument("--save_type", type=str, default="npy", choices=["npy", "hdf5"])  # parsed, unused (:306)
    parser.add_argument("--num_frames", type=int, default=8)
    parser.add_argument("--model_type", type=str, default="clip")
    parser.add_argument("--model_ckpt_dir", type=str, default=None)
    parser.add_argument("--use_crop", type=bool, default=False)  # any non-empty string is True, as in :311
    parser.add_argument("--inst_type", type=str, default="none")
    # extensions
    parser.add_argument("--arch", type=str, default="ViT-B/16")
    parser.add_argument("--reduce", type=str, default="first", choices=["first", "mean"])
    parser.add_argument("--max_batch", type=int, default=1024)
    parser.add_argument("--precision", type=str, default="16bit", choices=sorted(PRECISIONS))
    parser.add_argument("--tokenizer", type=str, default=None, choices=["clip", "standin"])
    args = parser.parse_args()

    env_name = f"{args.env_name}" if args.env_type == "none" else f"{args.env_name}_{args.env_type}"
    if args.inst_type != "none":
        text = get_clip_special_instruct(env_name, args.inst_type)
    else:
        text = get_clip_instruct(env_name)
    print(f"[INFO] env_name: {env_name}\t instruction: {text}")

    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")

    label_reward(
        env_name=args.env_name, env_type=args.env_type, distribution_mode=args.distribution_mode,
        image_keys=args.image_keys, data_path=args.data_path, text=text, num_levels=args.num_levels,
        start_level=args.start_level, num_demonstrations=args.num_demonstrations, num_frames=args.num_frames,
        base_path=args.base_path, model_type=args.model_type, model_ckpt_dir=args.model_ckpt_dir,
        use_crop=args.use_crop, inst_type=args.inst_type, arch=args.arch, reduce=args.reduce,
        max_batch=args.max_batch, precision=args.precision, tokenizer=args.tokenizer,
    )


if __name__ == "__main__":
    main()
