"""Runs the REFERENCE's own modules (byte-compiled into oracle/_ref by oracle/build_ref.py) on the CPU — TEST
INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by the product package.

The glue that cannot be imported from the reference is restated here and cites it line by line:
  * HIPT_4K.forward                    HIPT_4K/hipt_4k.py:63-76   (unfold, rearrange, 256-patch minibatches, .cpu() bounce,
                                                                  reshape/transpose to the [1,384,w,h] grid, ViT-4K)
  * HIPT_4K.prepare_img_tensor         HIPT_4K/hipt_4k.py:308-330 (CenterCrop to multiples of 256)
  * eval_transforms                    HIPT_4K/hipt_model_utils.py:113-118 (ToTensor + Normalize(0.5, 0.5))
  * per-slide pooling with every fold  eval.py -> utils/eval_utils.py:62-100 (model(data) per slide, per fold)
"""
import contextlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys

import torch

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_TOP = ("HIPT_4K", "models", "utils")
_cache = {}


def available():
    return os.path.exists(os.path.join(_REF, "HIPT_4K", "vision_transformer.refbin"))


class _RefFinder(importlib.abc.MetaPathFinder):
    """Imports the reference's byte-compiled modules from oracle/_ref/<package>/<module>.refbin (the bytes of a sourceless
    .pyc under an extension the GPU-box snapshot keeps); directories are packages."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] not in _TOP:
            return None
        rel = os.path.join(_REF, *fullname.split("."))
        if os.path.isfile(rel + ".refbin"):
            loader = importlib.machinery.SourcelessFileLoader(fullname, rel + ".refbin")
            return importlib.util.spec_from_file_location(fullname, rel + ".refbin", loader=loader)
        if os.path.isdir(rel):
            spec = importlib.machinery.ModuleSpec(fullname, None, is_package=True)
            spec.submodule_search_locations = [rel]
            return spec
        return None


@contextlib.contextmanager
def _isolated_imports():
    """The reference's top-level package names (HIPT_4K, models, utils) are also the names of this repository's import
    shims: import the reference under a private finder / sys.modules view and restore the caller's afterwards."""
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in _TOP}
    finder = _RefFinder()
    sys.meta_path.insert(0, finder)
    try:
        yield
    finally:
        sys.meta_path.remove(finder)
        for k in [k for k in sys.modules if k.split(".")[0] in _TOP]:
            del sys.modules[k]
        sys.modules.update(saved)


def modules():
    """(vision_transformer, vision_transformer4k, model_clam) modules of the reference."""
    if not _cache:
        if not available():
            raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
        with _isolated_imports():
            import HIPT_4K.vision_transformer as vits
            import HIPT_4K.vision_transformer4k as vits4k
            import models.model_clam as mclam
        _cache.update(vits=vits, vits4k=vits4k, mclam=mclam)
    return _cache["vits"], _cache["vits4k"], _cache["mclam"]


def build_models(seed=0, clam_seeds=(10, 11, 12, 13, 14), size_arg="hipt_smaller"):
    """Random-init reference modules at the seeds the GPU arm uses (same RNG consumption: same weights)."""
    vits, vits4k, mclam = modules()
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(sys.stderr):                        # vision_transformer4k.py prints "# of Patches"
        m256 = vits.vit_small(patch_size=16, num_classes=0).eval()
        m4k = vits4k.vit4k_xs(num_classes=0).eval()
    folds = []
    for s in clam_seeds:
        torch.manual_seed(s)
        folds.append(mclam.CLAM_SB(size_arg=size_arg, dropout=0.0, n_classes=2).eval())
    for m in (m256, m4k):
        for p in m.parameters():
            p.requires_grad = False
    return m256, m4k, folds


def eval_transforms_u8(region_u8):
    """[.., 3, W, H] uint8 -> fp32 in [-1, 1] (hipt_model_utils.py:113-118)."""
    return (region_u8.float() / 255.0 - 0.5) / 0.5


def hipt4k_forward(m256, m4k, x):
    """hipt_4k.py:63-76 on CPU: x [1,3,W,H] fp32 normalised -> [1,192]."""
    from einops import rearrange
    b, c, w, h = x.shape
    cw, ch = w - w % 256, h - h % 256                                   # prepare_img_tensor :308-330 (CenterCrop)
    top, left = int(round((w - cw) / 2.0)), int(round((h - ch) / 2.0))
    batch_256 = x[:, :, top:top + cw, left:left + ch]
    w_256, h_256 = w // 256, h // 256
    batch_256 = batch_256.unfold(2, 256, 256).unfold(3, 256, 256)       # :64
    batch_256 = rearrange(batch_256, 'b c p1 p2 w h -> (b p1 p2) c w h')  # :65 (201 MB copy at 4096 x 4096)
    feats = []
    for mini_bs in range(0, batch_256.shape[0], 256):                   # :68-70
        feats.append(m256(batch_256[mini_bs:mini_bs + 256]).detach().cpu())
    feats = torch.vstack(feats)                                         # :72
    grid = feats.reshape(w_256, h_256, 384).transpose(0, 1).transpose(0, 2).unsqueeze(dim=0)   # :73
    return m4k.forward(grid)                                            # :75


@torch.no_grad()
def slide_forward(m256, m4k, folds, regions_u8):
    """One slide the way the reference's scripts would: every region through eval_transforms + HIPT_4K.forward, then every
    fold's CLAM_SB over the bag.  Returns (features [R,192], [ (logits, Y_prob, Y_hat, A_raw) per fold ])."""
    feats = torch.cat([hipt4k_forward(m256, m4k, eval_transforms_u8(regions_u8[i:i + 1])) for i in range(regions_u8.shape[0])])
    outs = [f(feats)[:4] for f in folds]
    return feats, outs
