"""ORACLE / TEST INFRASTRUCTURE — ctypes view of oracle/_build/liboracle.so (the plain-C restatement)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "liboracle.so"


def build() -> Path:
    subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
    return LIB


def _lib():
    if not LIB.exists():
        build()
    lib = C.CDLL(str(LIB))
    lib.oracle_pil_bicubic_rgb.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.oracle_normalize_chw.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.oracle_discount_cumsum.argtypes = [C.c_void_p, C.c_int64, C.c_float, C.c_void_p]
    lib.oracle_stack_outputs.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    lib.oracle_episode_index.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    lib.oracle_episode_index.restype = C.c_int64
    return lib


def pil_bicubic(img: np.ndarray, out: int = 224) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    dst = np.empty((out, out, 3), np.uint8)
    assert _lib().oracle_pil_bicubic_rgb(img.ctypes.data, img.shape[0], img.shape[1], dst.ctypes.data, out) == 0
    return dst


def transform(img: np.ndarray, out: int = 224) -> np.ndarray:
    u8 = pil_bicubic(img, out)
    chw = np.empty((3, out, out), np.float32)
    _lib().oracle_normalize_chw(u8.ctypes.data, out, chw.ctypes.data)
    return chw


def discount_cumsum(x: np.ndarray, gamma: float = 1.0) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    _lib().oracle_discount_cumsum(x.ctypes.data, x.size, gamma, out.ctypes.data)
    return out


def stack_outputs(x: np.ndarray, num_frames: int) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty((x.size, num_frames), np.float32)
    _lib().oracle_stack_outputs(x.ctypes.data, x.size, num_frames, out.ctypes.data)
    return out


def episode_index(done_last: np.ndarray) -> np.ndarray:
    d = np.ascontiguousarray(done_last, np.float32)
    idx = np.empty(d.size + 1, np.int64)
    n = _lib().oracle_episode_index(d.ctypes.data, d.size, idx.ctypes.data, idx.size)
    return idx[:n]
