"""TEST-ONLY stand-in for h5py (not installed in this image; SURVEY.md §7 "Missing deps"): just enough of the File / Dataset
surface for the reference's own utils/file_utils.save_hdf5 (resize-append), datasets/dataset_h5.Whole_Slide_Bag_FP (coords +
attrs) and the re-read in extract_features_fp.py:248-255.  Backed by one pickle per file."""
import os
import pickle

import numpy as np


class Dataset:
    def __init__(self, data, attrs=None):
        self.data = data
        self.attrs = attrs if attrs is not None else {}

    @property
    def shape(self):
        return self.data.shape

    @property
    def dtype(self):
        return self.data.dtype

    def __len__(self):
        return len(self.data)

    def __getitem__(self, key):
        return self.data[key]

    def __setitem__(self, key, value):
        self.data[key] = value

    def resize(self, size, axis=0):
        shape = list(self.data.shape)
        shape[axis] = size
        new = np.zeros(shape, dtype=self.data.dtype)
        n = min(size, self.data.shape[axis])
        sl = [slice(None)] * self.data.ndim
        sl[axis] = slice(0, n)
        new[tuple(sl)] = self.data[tuple(sl)]
        self.data = new


class File:
    def __init__(self, path, mode="r"):
        self.path, self.mode = path, mode
        self.sets = {}
        if mode in ("r", "a", "r+") and os.path.exists(path):
            with open(path, "rb") as f:
                raw = pickle.load(f)
            self.sets = {k: Dataset(v[0], v[1]) for k, v in raw.items()}
        elif mode == "r":
            raise OSError(f"Unable to open file {path}")

    def __contains__(self, key):
        return key in self.sets

    def __getitem__(self, key):
        return self.sets[key]

    def keys(self):
        return self.sets.keys()

    def create_dataset(self, name, shape=None, maxshape=None, chunks=None, dtype=None, data=None):
        arr = np.array(data) if data is not None else np.zeros(shape, dtype=dtype)
        self.sets[name] = Dataset(arr)
        return self.sets[name]

    def close(self):
        if self.mode != "r":
            with open(self.path, "wb") as f:
                pickle.dump({k: (d.data, dict(d.attrs)) for k, d in self.sets.items()}, f)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
