"""ORACLE / TEST INFRASTRUCTURE — not part of the shipped product.

In-memory stand-in for the subset of h5py that arp_dt/label_reward.py uses (h5py is not installable
here): File(path, mode) keyed by path in a process-global registry so a test can build a dataset,
run the reference on it, and read the labels back. Semantics honoured (reference line in brackets):
  File.get(key) -> Dataset | None [71-75, 264]      truthiness of a Dataset = len(shape) > 0 [71, 275]
  ds.shape [81, 104]   ds[:, -1] [82]   ds[list_of_rows, -1] [268]   ds[list] = arr [289]
  ds[-n:] = arr [287]  ds.resize(n, axis=0) [286]
  create_dataset(key, compression=, chunks=, maxshape=, data=) [277-283]     close() [291]
"""
from __future__ import annotations

import numpy as np

_STORE: dict[str, dict[str, "Dataset"]] = {}


class Dataset:
    def __init__(self, data, maxshape=None, chunks=None, compression=None):
        self._a = np.array(data)
        self.maxshape, self.chunks, self.compression = maxshape, chunks, compression

    shape = property(lambda self: self._a.shape)
    dtype = property(lambda self: self._a.dtype)
    ndim = property(lambda self: self._a.ndim)

    def __len__(self):
        return self._a.shape[0]

    def __bool__(self):  # h5py: a dataset object is truthy when it is a valid (open) handle
        return True

    def __getitem__(self, idx):
        return self._a[idx]

    def __setitem__(self, idx, value):
        self._a[idx] = value

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    def resize(self, size, axis=None):
        if axis is None:
            new_shape = tuple(size)
        else:
            new_shape = list(self._a.shape)
            new_shape[axis] = int(size)
            new_shape = tuple(new_shape)
        if self.maxshape is not None:
            for n, m in zip(new_shape, self.maxshape):
                if m is not None and n > m:
                    raise ValueError(f"resize to {new_shape} exceeds maxshape {self.maxshape}")
        out = np.zeros(new_shape, dtype=self._a.dtype)
        sl = tuple(slice(0, min(a, b)) for a, b in zip(self._a.shape, new_shape))
        out[sl] = self._a[sl]
        self._a = out


class File:
    def __init__(self, path, mode="r"):
        self.path, self.mode = str(path), mode
        if mode in ("w",):
            _STORE[self.path] = {}
        self._d = _STORE.setdefault(self.path, {})
        self.attrs = {}

    def get(self, key, default=None):
        return self._d.get(key, default)

    def __getitem__(self, key):
        return self._d[key]

    def __contains__(self, key):
        return key in self._d

    def keys(self):
        return self._d.keys()

    def create_dataset(self, key, shape=None, dtype=None, data=None, compression=None, chunks=None, maxshape=None,
                       **_):
        if key in self._d:
            raise ValueError(f"Unable to create dataset (name already exists): {key}")
        if data is None:
            data = np.zeros(shape, dtype=dtype or np.float32)
        ds = Dataset(np.asarray(data, dtype=dtype) if dtype else data, maxshape=maxshape, chunks=chunks,
                     compression=compression)
        self._d[key] = ds
        return ds

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def reset():
    _STORE.clear()
