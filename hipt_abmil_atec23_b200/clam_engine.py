"""Host side of the fused CLAM_SB pooling kernel (hb_clam_sb_forward in include/hipt_b200.h).

`forward_single` serves CLAM_SB.forward (one bag, one model — models/model_clam.py:147-191); `forward_bags` is the
batched form used for slide sets and fold ensembles: many ragged bags and up to 8 weight sets in one launch, the
feature matrix read from HBM once.
"""
import ctypes as C

import torch

from . import _lib


def _gate_module(model):
    return model.attention_net[-1]


def _weights(model, device):
    """The 10 fp32 tensors of one CLAM_SB in the order hb_clam_sb_forward documents."""
    g = _gate_module(model)
    ts = [model.attention_net[0].weight, model.attention_net[0].bias,
          g.attention_a[0].weight, g.attention_a[0].bias, g.attention_b[0].weight, g.attention_b[0].bias,
          g.attention_c.weight, g.attention_c.bias, model.classifiers.weight, model.classifiers.bias]
    out = []
    for t in ts:
        t = t.detach()
        if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(device=device, dtype=torch.float32).contiguous()
        out.append(t)
    return out


def forward_bags(models, feats, bag_offsets, max_bag_len=None, want=("logits", "y_prob", "y_hat", "m")):
    """models: list of CLAM_SB (same size_arg / n_classes); feats [total, L0] fp32 CUDA; bag_offsets int32 [n_bags+1]
    (CUDA or CPU).  Returns dict with a_raw [n_models, total] and the requested per-bag outputs [n_models, n_bags, ...]."""
    _lib.require_cuda(feats, "features")
    lib = _lib.load()
    dev = feats.device
    if feats.dtype != torch.float32 or not feats.is_contiguous():
        feats = feats.float().contiguous()
    total, L0 = feats.shape
    n_models = len(models)
    m0 = models[0]
    L1 = m0.attention_net[0].out_features
    D = _gate_module(m0).attention_c.in_features
    Cc = m0.classifiers.out_features
    if _gate_module(m0).attention_c.out_features != 1:
        raise RuntimeError("the fused kernel implements the single-branch (CLAM_SB) head")
    if max_bag_len is None:
        off_cpu = bag_offsets.cpu()
        lens = off_cpu[1:] - off_cpu[:-1]
        max_bag_len = int(lens.max().item()) if lens.numel() else 0
    offs = bag_offsets.to(device=dev, dtype=torch.int32).contiguous()
    n_bags = offs.numel() - 1
    with torch.cuda.device(dev):
        _lib.device_check()
        keep = []
        ptrs = []
        for m in models:
            w = _weights(m, dev)
            keep.append(w)
            ptrs += [t.data_ptr() for t in w]
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        a_raw = torch.empty((n_models, total), dtype=torch.float32, device=dev)
        m_out = torch.empty((n_models, n_bags, L1), dtype=torch.float32, device=dev) if "m" in want else None
        logits = torch.empty((n_models, n_bags, Cc), dtype=torch.float32, device=dev) if "logits" in want else None
        y_prob = torch.empty((n_models, n_bags, Cc), dtype=torch.float32, device=dev) if "y_prob" in want else None
        y_hat = torch.empty((n_models, n_bags), dtype=torch.int64, device=dev) if "y_hat" in want else None
        ws_bytes = lib.hb_clam_workspace_bytes(max_bag_len, n_bags, n_models, L1)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        _lib.check(lib.hb_clam_sb_forward(_lib.ptr(feats), _lib.ptr(offs), n_bags, total, max_bag_len, arr, n_models,
                                          L0, L1, D, Cc, _lib.ptr(a_raw), _lib.ptr(m_out), _lib.ptr(logits),
                                          _lib.ptr(y_prob), _lib.ptr(y_hat), _lib.ptr(ws), ws.numel(),
                                          _lib.stream_ptr()))
    return {"a_raw": a_raw, "m": m_out, "logits": logits, "y_prob": y_prob, "y_hat": y_hat}


def forward_single(model, h, attention_only=False):
    """One bag through one model with the reference's return shapes: logits [1,C], Y_prob [1,C], Y_hat [1,1] int64,
    A_raw [1,N], M [1,L1]  (A_raw alone when attention_only)."""
    N = h.shape[0]
    offs = torch.tensor([0, N], dtype=torch.int32)
    want = () if attention_only else ("logits", "y_prob", "y_hat", "m")
    r = forward_bags([model], h, offs, max_bag_len=N, want=want)
    if attention_only:
        return r["a_raw"]
    return r["logits"][0], r["y_prob"][0], r["y_hat"][0].view(1, 1), r["a_raw"], r["m"][0]
