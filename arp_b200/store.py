"""Dataset container access for the labeler.

The reference reads and writes one HDF5 file through h5py (label_reward.py:69, layout written by
data/PPG/trajectory_recorder.py:134-189). h5py is used when it is importable.
It is not part of this image, so the same subset of the h5py API is also provided over a plain
directory of .npy files (`NpyStore`): datasets are memory-mapped, so a [T,F,H,W,3] uint8 `ob` is
never loaded whole, and the labeler can hand strided host pointers straight to the native library.
Real HDF5 chunk/gzip I/O is a "next" row of the scope table.
"""
from __future__ import annotations

import json
import os
from pathlib import Path

import numpy as np


class NpyDataset:
    """np.memmap-backed dataset with the h5py.Dataset calls label_reward needs."""

    def __init__(self, path: Path, mode: str, maxshape=None):
        self._path, self._mode, self.maxshape = Path(path), mode, maxshape
        self._a = np.load(self._path, mmap_mode="r+" if mode != "r" else "r")

    shape = property(lambda self: self._a.shape)
    dtype = property(lambda self: self._a.dtype)

    def __len__(self):
        return self._a.shape[0]

    def __bool__(self):
        return True

    def __getitem__(self, idx):
        return self._a[idx]

    def __setitem__(self, idx, value):
        self._a[idx] = value

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._a) if dtype is None else np.asarray(self._a, dtype=dtype)

    @property
    def array(self) -> np.ndarray:
        """The underlying (memory-mapped) ndarray, C-contiguous — lets the caller pass raw pointers."""
        return self._a

    def file_source(self):
        """(path, byte offset of element [0, ...]) when the rows sit contiguously in the backing file — lets the labeler
        hand the FILE to the native library (pread into pinned memory) instead of touching the mapping page by page."""
        off = getattr(self._a, "offset", None)
        if off is None or not self._a.flags.c_contiguous:
            return None
        return self._path, int(off)

    def resize(self, size, axis=None):
        new_shape = list(self._a.shape)
        if axis is None:
            new_shape = list(size)
        else:
            new_shape[axis] = int(size)
        if self.maxshape is not None:
            for n, m in zip(new_shape, self.maxshape):
                if m is not None and n > m:
                    raise ValueError(f"resize to {tuple(new_shape)} exceeds maxshape {self.maxshape}")
        old = np.array(self._a)
        del self._a
        out = np.zeros(new_shape, dtype=old.dtype)
        sl = tuple(slice(0, min(a, b)) for a, b in zip(old.shape, new_shape))
        out[sl] = old[sl]
        np.save(self._path, out)
        self._a = np.load(self._path, mmap_mode="r+")

    def flush(self):
        if hasattr(self._a, "flush"):
            self._a.flush()


class NpyStore:
    """Directory of <key>.npy files + attrs.json, with the h5py.File subset used by the labeler."""

    def __init__(self, path, mode: str = "r"):
        self.path, self.mode = Path(path), mode
        if mode in ("w",) and self.path.exists():
            for p in self.path.glob("*.npy"):
                p.unlink()
        if mode != "r":
            self.path.mkdir(parents=True, exist_ok=True)
        elif not self.path.is_dir():
            raise FileNotFoundError(f"{self.path} is not an NpyStore directory")
        self._open: dict[str, NpyDataset] = {}
        attrs = self.path / "attrs.json"
        self.attrs = json.loads(attrs.read_text()) if attrs.exists() else {}

    def _file(self, key: str) -> Path:
        return self.path / (key.replace("/", "__") + ".npy")

    def keys(self):
        return sorted(p.stem.replace("__", "/") for p in self.path.glob("*.npy"))

    def __contains__(self, key):
        return self._file(key).exists()

    def get(self, key, default=None):
        return self[key] if key in self else default

    def __getitem__(self, key) -> NpyDataset:
        if key not in self._open:
            if key not in self:
                raise KeyError(key)
            self._open[key] = NpyDataset(self._file(key), self.mode)
        return self._open[key]

    def create_dataset(self, key, shape=None, dtype=None, data=None, compression=None, chunks=None, maxshape=None, **_):
        if self.mode == "r":
            raise OSError("store opened read-only")
        if key in self:
            raise ValueError(f"Unable to create dataset (name already exists): {key}")
        arr = np.zeros(shape, dtype=dtype or np.float32) if data is None else np.asarray(data, dtype=dtype)
        np.save(self._file(key), arr)
        ds = NpyDataset(self._file(key), self.mode, maxshape=maxshape)
        self._open[key] = ds
        return ds

    def close(self):
        for ds in self._open.values():
            ds.flush()
        self._open.clear()
        if self.mode != "r" and self.attrs:
            (self.path / "attrs.json").write_text(json.dumps(self.attrs))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def open_store(path, mode: str = "a"):
    """h5py.File(path, mode) when the path is an HDF5 file and h5py exists; NpyStore for a directory."""
    p = Path(path)
    if p.is_dir() or (not p.exists() and p.suffix not in (".hdf5", ".h5")):
        return NpyStore(p, mode)
    try:
        import h5py  # type: ignore
    except ImportError as e:
        raise ImportError(f"{p} looks like an HDF5 file but h5py is not installed; convert it to an NpyStore "
                          "directory or install h5py") from e
    return h5py.File(os.fspath(p), mode)
