"""Feature sink: a slide's region embeddings in the on-disk layout the reference's training / evaluation scripts read.

Reference: extract_features_fp.py:159-171 appends every region's [1,192] row to `<feat_dir>/h5_files/<slide>.h5` through
utils/file_utils.save_hdf5 (one HDF5 resize + chunk (1,192) write per region), then :248-255 re-reads the whole file and
torch.saves the [N,192] tensor to `<feat_dir>/pt_files/<slide>.pt`; main.py / eval.py read the .pt
(datasets/dataset_generic.py:505-528: `features = torch.load(.../pt_files/<slide_id>.pt)`), create_heatmaps / sampling
read `features` + `coords` from the .h5.  At B200 rates (thousands of regions per second per node) the per-region HDF5
append is thousands of syscalls a second, so a slide is written ONCE here: the .pt straight from the [N,192] tensor the
pipeline produced, and the .h5 (datasets `features` [N,192] float32 and `coords` [N,2], same names, same chunk shape) when
h5py is importable.
"""
import os

import numpy as np
import torch


def _h5py():
    try:
        import h5py
        return h5py
    except ImportError:
        return None


def save_slide_features(feat_dir, slide_id, features, coords=None, attrs=None, write_h5=None):
    """features [N,192] (any device), coords [N,2] int (region origins in level-0 pixels) or None.
    write_h5: True = require h5py, False = .pt only, None = write the .h5 when h5py is available.
    Returns {"pt": path, "h5": path or None}."""
    feats = features.detach().to("cpu", torch.float32).contiguous()
    if feats.dim() != 2:
        raise ValueError(f"features must be [N, F], got {tuple(feats.shape)}")
    os.makedirs(os.path.join(feat_dir, "pt_files"), exist_ok=True)
    pt_path = os.path.join(feat_dir, "pt_files", f"{slide_id}.pt")
    tmp = pt_path + ".tmp"
    torch.save(feats, tmp)                                   # what extract_features_fp.py:255 saves: the bare tensor
    os.replace(tmp, pt_path)                                 # a reader never sees a half-written bag
    h5_path = None
    h5py = _h5py() if write_h5 is not False else None
    if write_h5 and h5py is None:
        raise RuntimeError("write_h5=True needs h5py, which is not installed")
    if h5py is not None:
        os.makedirs(os.path.join(feat_dir, "h5_files"), exist_ok=True)
        h5_path = os.path.join(feat_dir, "h5_files", f"{slide_id}.h5")
        arr = feats.numpy()
        with h5py.File(h5_path, "w") as f:
            f.create_dataset("features", shape=arr.shape, maxshape=(None,) + arr.shape[1:], chunks=(1,) + arr.shape[1:],
                             dtype=arr.dtype)[:] = arr
            if coords is not None:
                c = np.asarray(coords.cpu() if torch.is_tensor(coords) else coords)
                if c.shape[0] != arr.shape[0]:
                    raise ValueError("coords and features disagree on the number of regions")
                d = f.create_dataset("coords", shape=c.shape, maxshape=(None,) + c.shape[1:], chunks=(1,) + c.shape[1:],
                                     dtype=c.dtype)
                d[:] = c
                for k, v in (attrs or {}).items():
                    d.attrs[k] = v
    return {"pt": pt_path, "h5": h5_path}


def load_slide_features(feat_dir, slide_id):
    """The read of Generic_MIL_Dataset.__getitem__ (datasets/dataset_generic.py:505-512, use_h5 = False)."""
    return torch.load(os.path.join(feat_dir, "pt_files", f"{slide_id}.pt"))


def save_slide_set(feat_dir, slide_ids, bags, coords=None, write_h5=None):
    """One call per slide set: `bags` is a list of [n_i,192] tensors (e.g. ShardedSlideSet results split at its bag offsets)."""
    return [save_slide_features(feat_dir, s, b, None if coords is None else coords[i], write_h5=write_h5)
            for i, (s, b) in enumerate(zip(slide_ids, bags))]
