"""Weight loaders and input transforms of the HIPT_4K path (interface of HIPT_4K/hipt_model_utils.py in the reference).

`get_vit256` / `get_vit4k` follow hipt_model_utils.py:39-73 / :76-110: build the architecture, freeze it, put it in eval
mode on the CPU, read the checkpoint with map_location='cpu', take its 'teacher' entry when present, strip the
`module.` and `backbone.` prefixes, load non-strictly, and fail with an AssertionError when the file is missing.
The reference file itself cannot be imported (TabError at :72), so this is a re-statement of its behaviour, not of its
text; plotting / webdataset helpers that live in the same reference file are not part of the hot path and are omitted.
"""
import os

import numpy as np
import torch

from . import vision_transformer as vits
from . import vision_transformer4k as vits4k


def _load_frozen(model, pretrained_weights, checkpoint_key="teacher"):
    for p in model.parameters():
        p.requires_grad = False
    model.eval()
    model.to(torch.device("cpu"))
    assert os.path.isfile(pretrained_weights), "pretrained weights not available at {}".format(pretrained_weights)
    state_dict = torch.load(pretrained_weights, map_location="cpu")
    if checkpoint_key is not None and checkpoint_key in state_dict:
        print(f"Take key {checkpoint_key} in provided checkpoint dict")
        state_dict = state_dict[checkpoint_key]
    state_dict = {k.replace("module.", "").replace("backbone.", ""): v for k, v in state_dict.items()}
    msg = model.load_state_dict(state_dict, strict=False)
    print("Pretrained weights found at {} and loaded with msg: {}".format(pretrained_weights, msg))
    return model


def get_vit256(pretrained_weights, arch="vit_small", device=torch.device("cuda:0")):
    """Builds the ViT-256 model (frozen, eval, on CPU — the caller moves it, as HIPT_4K.__init__ does)."""
    return _load_frozen(vits.__dict__[arch](patch_size=16, num_classes=0), pretrained_weights)


def get_vit4k(pretrained_weights, arch="vit4k_xs", device=torch.device("cuda:1")):
    """Builds the ViT-4K model (frozen, eval, on CPU)."""
    return _load_frozen(vits4k.__dict__[arch](num_classes=0), pretrained_weights)


HIPT_MEAN = (0.5, 0.5, 0.5)
HIPT_STD = (0.5, 0.5, 0.5)


class _ToNormalizedTensor:
    """ToTensor + Normalize(mean, std) for PIL images / HxWx3 uint8 arrays (hipt_model_utils.py:113-118) without
    importing torchvision at module import time."""

    def __init__(self, mean, std):
        self.mean = torch.tensor(mean, dtype=torch.float32).view(3, 1, 1)
        self.std = torch.tensor(std, dtype=torch.float32).view(3, 1, 1)

    def __call__(self, img):
        arr = np.asarray(img)
        if arr.ndim == 2:
            arr = arr[:, :, None]
        t = torch.from_numpy(np.ascontiguousarray(arr)).permute(2, 0, 1).contiguous()   # CHW, as torchvision's ToTensor
        t = t.float().div(255.0) if t.dtype == torch.uint8 else t.float()
        return (t - self.mean) / self.std


def eval_transforms():
    return _ToNormalizedTensor(HIPT_MEAN, HIPT_STD)


def roll_batch2img(batch: torch.Tensor, w: int, h: int, patch_size=256):
    """[B,3,ps,ps] patches (row-major w x h grid) -> one [3, w*ps, h*ps] uint8 HWC image (hipt_model_utils.py:121-133)."""
    b = batch.reshape(w, h, 3, patch_size, patch_size)
    img = b.permute(2, 0, 3, 1, 4).reshape(3, w * patch_size, h * patch_size).unsqueeze(0)
    return tensorbatch2im(img)[0]


def tensorbatch2im(input_image, imtype=np.uint8):
    """Undo the (0.5, 0.5) normalisation of a [B,3,H,W] batch into [B,H,W,3] uint8 (hipt_model_utils.py:136-154)."""
    if isinstance(input_image, np.ndarray):
        return input_image
    x = input_image.detach().cpu().float().numpy()
    x = (np.transpose(x, (0, 2, 3, 1)) + 1) / 2.0 * 255.0
    return x.astype(imtype)
