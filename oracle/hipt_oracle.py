"""CPU fp32 ORACLE for the HIPT_4K + CLAM_SB hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional PyTorch on the CPU, the arithmetic the reference performs on the hot path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it; the product
package (hipt_abmil_atec23_b200/) never does, and fails loudly without its CUDA library.

Parity pinning: the reference has no tests, golden vectors or fixtures for this path, so the oracle
is pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/make_golden.py imports the reference's own modules from
/root/reference (HIPT_4K.vision_transformer, HIPT_4K.vision_transformer4k, models.model_clam, models.model_mil) in the
build container, runs them at fixed seeds and stores the results under tests/golden/; tests/test_oracle_golden.py checks
every function below against those files.  HIPT_4K/hipt_4k.py and HIPT_4K/hipt_model_utils.py cannot be imported
(TabError at hipt_model_utils.py:72, missing h5py/matplotlib/skimage/webdataset), so `hipt4k_forward`,
`prepare_img_tensor` and `eval_transforms_u8` restate hipt_4k.py:63-76, :308-330 and hipt_model_utils.py:113-118 and are
cross-checked in make_golden.py against the reference's second statement of the same forward
(HIPT_4K/attention_visualization_utils.py:424-441 hipt_forward_pass).

All state dicts use the reference's key names.
"""
import hashlib
import math

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------- ViT building blocks
def interpolate_pos_encoding(pos_embed, n_tokens, w0, h0):
    """HIPT_4K/vision_transformer.py:213-233 and vision_transformer4k.py:201-221.

    pos_embed [1, 1+N, dim]; n_tokens = tokens excluding CLS; (w0, h0) = token grid (w // patch_size for ViT-256,
    w // 1 for ViT-4K).  Bicubic resize of the sqrt(N) x sqrt(N) grid with scale_factor ((w0+0.1)/sqrt(N), ...)."""
    N = pos_embed.shape[1] - 1
    if n_tokens == N and w0 == h0:
        return pos_embed
    dim = pos_embed.shape[-1]
    cls_pos = pos_embed[:, 0]
    patch_pos = pos_embed[:, 1:]
    s = int(math.sqrt(N))
    wf, hf = w0 + 0.1, h0 + 0.1
    patch_pos = F.interpolate(patch_pos.reshape(1, s, s, dim).permute(0, 3, 1, 2),
                              scale_factor=(wf / math.sqrt(N), hf / math.sqrt(N)), mode="bicubic")
    assert int(wf) == patch_pos.shape[-2] and int(hf) == patch_pos.shape[-1]
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((cls_pos.unsqueeze(0), patch_pos), dim=1)


def attention(sd, pre, x, heads):
    """Attention.forward, vision_transformer.py:119-131 (returns only x; the attn matrix is not used by forward)."""
    B, N, C = x.shape
    hd = C // heads
    qkv = F.linear(x, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"]).reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    att = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
    att = att.softmax(dim=-1)
    y = (att @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(y, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def block(sd, pre, x, heads, eps=1e-6):
    """Block.forward, vision_transformer.py:146-152 (DropPath = Identity, Dropout p = 0)."""
    C = x.shape[-1]
    y = F.layer_norm(x, (C,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps)
    x = x + attention(sd, pre + "attn.", y, heads)
    y = F.layer_norm(x, (C,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps)
    y = F.linear(y, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])      # Mlp.forward :98-104
    y = F.gelu(y)                                                               # nn.GELU() = exact erf
    y = F.linear(y, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
    return x + y


def last_selfattention(sd, tokens, heads):
    """get_last_selfattention, vision_transformer.py:255-262 / vision_transformer4k.py:248-255: blocks 0..depth-2 as usual,
    then the last block's attention probabilities (Block.forward(return_attention=True) :147-149) [B, heads, N, N]."""
    n = _depth(sd)
    t = tokens
    for i in range(n - 1):
        t = block(sd, f"blocks.{i}.", t, heads)
    pre = f"blocks.{n - 1}."
    C = t.shape[-1]
    y = F.layer_norm(t, (C,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-6)
    B, N, _ = y.shape
    hd = C // heads
    qkv = F.linear(y, sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"]).reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    return ((qkv[0] @ qkv[1].transpose(-2, -1)) * (hd ** -0.5)).softmax(dim=-1)


def _depth(sd):
    return 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))


def vit256_tokens(sd, x):
    """prepare_tokens, vision_transformer.py:235-246: patch-embed conv (as an explicit (c,i,j)-ordered GEMM), CLS, pos."""
    B, nc, w, h = x.shape
    W = sd["patch_embed.proj.weight"]
    ps = W.shape[-1]
    cols = F.unfold(x, kernel_size=ps, stride=ps)                     # [B, nc*ps*ps, T], K order (c, i, j)
    tok = cols.transpose(1, 2) @ W.reshape(W.shape[0], -1).t() + sd["patch_embed.proj.bias"]
    cls = sd["cls_token"].expand(B, -1, -1)
    tok = torch.cat((cls, tok), dim=1)
    return tok + interpolate_pos_encoding(sd["pos_embed"], tok.shape[1] - 1, w // ps, h // ps)


def vit256_forward(sd, x, heads=6, return_tokens=False, depth_limit=None):
    """VisionTransformer.forward, vision_transformer.py:248-253: [B,3,256,256] -> CLS [B,384]."""
    t = vit256_tokens(sd, x)
    n = _depth(sd) if depth_limit is None else depth_limit
    for i in range(n):
        t = block(sd, f"blocks.{i}.", t, heads)
    if return_tokens:
        return t
    t = F.layer_norm(t, (t.shape[-1],), sd["norm.weight"], sd["norm.bias"], 1e-6)
    return t[:, 0]


def vit4k_tokens(sd, grid):
    """VisionTransformer4K.prepare_tokens, vision_transformer4k.py:223-239: [B,384,w,h] -> [B,1+w*h,192]."""
    B, _, w, h = grid.shape
    t = grid.flatten(2, 3).transpose(1, 2)
    t = F.gelu(F.linear(t, sd["phi.0.weight"], sd["phi.0.bias"]))
    cls = sd["cls_token"].expand(B, -1, -1)
    t = torch.cat((cls, t), dim=1)
    return t + interpolate_pos_encoding(sd["pos_embed"], t.shape[1] - 1, w // 1, h // 1)


def vit4k_forward(sd, grid, heads=6):
    """VisionTransformer4K.forward, vision_transformer4k.py:241-246."""
    t = vit4k_tokens(sd, grid)
    for i in range(_depth(sd)):
        t = block(sd, f"blocks.{i}.", t, heads)
    t = F.layer_norm(t, (t.shape[-1],), sd["norm.weight"], sd["norm.bias"], 1e-6)
    return t[:, 0]


# ---------------------------------------------------------------------------------------------- HIPT_4K glue
def eval_transforms_u8(img_u8):
    """hipt_model_utils.py:113-118 on an already-decoded uint8 [.., 3, H, W] tensor: ToTensor (/255) then
    Normalize(mean=0.5, std=0.5)."""
    return (img_u8.float() / 255.0 - 0.5) / 0.5


def prepare_img_tensor(img, patch_size=256):
    """hipt_4k.py:308-330: centre-crop so both spatial dims are multiples of patch_size (torchvision CenterCrop
    offsets: int(round((size - crop) / 2)))."""
    b, c, w, h = img.shape
    cw, ch = w - w % patch_size, h - h % patch_size
    top = int(round((w - cw) / 2.0))
    left = int(round((h - ch) / 2.0))
    return img[:, :, top:top + cw, left:left + ch], w // patch_size, h // patch_size


def unfold_region(img):
    """hipt_4k.py:64-65: unfold(2,256,256).unfold(3,256,256) + rearrange 'b c p1 p2 w h -> (b p1 p2) c w h'."""
    b, c = img.shape[:2]
    p = img.unfold(2, 256, 256).unfold(3, 256, 256)                    # [b, c, p1, p2, 256, 256]
    return p.permute(0, 2, 3, 1, 4, 5).reshape(-1, c, 256, 256)


def hipt4k_forward(sd256, sd4k, x, return_cls256=False):
    """HIPT_4K.forward, hipt_4k.py:48-76: [1,3,W,H] normalised fp32 -> [1,192]."""
    img, w_256, h_256 = prepare_img_tensor(x)
    batch = unfold_region(img)
    feats = []
    for i in range(0, batch.shape[0], 256):                           # :68-70 minibatches of 256
        feats.append(vit256_forward(sd256, batch[i:i + 256]))
    cls256 = torch.vstack(feats)
    grid = cls256.reshape(w_256, h_256, 384).transpose(0, 1).transpose(0, 2).unsqueeze(0)   # :73
    out = vit4k_forward(sd4k, grid)
    return (out, cls256) if return_cls256 else out


# ---------------------------------------------------------------------------------------------- CLAM / MIL heads
def _gate_prefix(sd):
    for k in sd:
        if k.endswith("attention_a.0.weight"):
            return k[: -len("attention_a.0.weight")]
    raise KeyError("no gated attention in state dict")


def clam_sb_forward(sd, h, attention_only=False, return_features=False):
    """CLAM_SB.forward (eval mode, instance_eval=False), models/model_clam.py:147-191 with Attn_Net_Gated :59-64."""
    g = _gate_prefix(sd)
    h1 = F.relu(F.linear(h, sd["attention_net.0.weight"], sd["attention_net.0.bias"]))
    a = torch.tanh(F.linear(h1, sd[g + "attention_a.0.weight"], sd[g + "attention_a.0.bias"]))
    b = torch.sigmoid(F.linear(h1, sd[g + "attention_b.0.weight"], sd[g + "attention_b.0.bias"]))
    A = F.linear(a * b, sd[g + "attention_c.weight"], sd[g + "attention_c.bias"])
    A = A.transpose(1, 0)
    if attention_only:
        return A
    A_raw = A
    A = F.softmax(A, dim=1)
    M = A @ h1
    logits = F.linear(M, sd["classifiers.weight"], sd["classifiers.bias"])
    Y_hat = torch.topk(logits, 1, dim=1)[1]
    Y_prob = F.softmax(logits, dim=1)
    res = {"features": M} if return_features else {}
    return logits, Y_prob, Y_hat, A_raw, res


def mil_fc_forward(sd, h, top_k=1):
    """MIL_fc.forward (return_features=False), models/model_mil.py:26-43; classifier = Linear, ReLU, [Dropout], Linear."""
    last = max(int(k.split(".")[1]) for k in sd if k.startswith("classifier."))
    z = F.relu(F.linear(h, sd["classifier.0.weight"], sd["classifier.0.bias"]))
    logits = F.linear(z, sd[f"classifier.{last}.weight"], sd[f"classifier.{last}.bias"])
    y_probs = F.softmax(logits, dim=1)
    idx = torch.topk(y_probs[:, 1], top_k, dim=0)[1].view(1,)
    top = torch.index_select(logits, 0, idx)
    return top, F.softmax(top, dim=1), torch.topk(top, 1, dim=1)[1], y_probs, {}


# ---------------------------------------------------------------------------------------------- seeded inputs (§8d)
def synthetic_region_u8(seed=1, size=4096):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (1, 3, size, size), dtype=torch.uint8, generator=g)


def sd_digest(sd):
    """Order-independent fingerprint of a state dict (float64 sums + a sha1 of a few raw tensors)."""
    h = hashlib.sha1()
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].detach().double()
        tot += float(v.sum()) + 0.5 * float((v * v).sum())
        if v.numel() <= 4096:
            h.update(sd[k].detach().contiguous().numpy().tobytes())
    return {"sum": tot, "sha1_small": h.hexdigest(), "n": len(sd)}


# ---------------------------------------------------------------------------------------------------- CLAM_SB training mode
def clam_sb_forward_train(sd, h, masks=None, label=None, instance_eval=False, k_sample=8, subtyping=False,
                          n_classes=2, inst_loss=None):
    """CLAM_SB.forward in TRAINING mode with explicit dropout masks and the instance-clustering branch, differentiable
    w.r.t. the tensors of `sd` (models/model_clam.py:147-191):
      masks = (m1 [N,L1], ma [N,D], mb [N,D]) holding 0 or 1/(1-p) — what nn.Dropout multiplies by after the ReLU (:84-85)
              and after Tanh / Sigmoid inside Attn_Net_Gated (:50-52, applied BEFORE a * b at :62); None = no dropout
      instance_eval: inst_eval :116-132 for the label's class (top-k / bottom-k of the SOFTMAXED scores, index_select on the
              post-dropout h), inst_eval_out :135-145 for the other classes when subtyping, loss averaged over the
              instance classifiers only when subtyping (:170-171)
    Returns (logits, Y_prob, Y_hat, A_raw, results_dict)."""
    g = _gate_prefix(sd)
    h1 = F.relu(F.linear(h, sd["attention_net.0.weight"], sd["attention_net.0.bias"]))
    if masks is not None:
        h1 = h1 * masks[0]
    a = torch.tanh(F.linear(h1, sd[g + "attention_a.0.weight"], sd[g + "attention_a.0.bias"]))
    b = torch.sigmoid(F.linear(h1, sd[g + "attention_b.0.weight"], sd[g + "attention_b.0.bias"]))
    if masks is not None:
        a, b = a * masks[1], b * masks[2]
    A = F.linear(a * b, sd[g + "attention_c.weight"], sd[g + "attention_c.bias"]).transpose(1, 0)
    A_raw = A
    A = F.softmax(A, dim=1)
    res = {}
    if instance_eval:
        loss_fn = inst_loss or torch.nn.CrossEntropyLoss()
        total, preds_all, targets_all = 0.0, [], []
        onehot = F.one_hot(label, num_classes=n_classes).squeeze()
        for i in range(n_classes):
            W, bi = sd[f"instance_classifiers.{i}.weight"], sd[f"instance_classifiers.{i}.bias"]
            if onehot[i].item() == 1:
                top_p = torch.index_select(h1, 0, torch.topk(A, k_sample)[1][-1])
                top_n = torch.index_select(h1, 0, torch.topk(-A, k_sample, dim=1)[1][-1])
                targets = torch.cat([torch.ones(k_sample, dtype=torch.long), torch.zeros(k_sample, dtype=torch.long)])
                lg = F.linear(torch.cat([top_p, top_n]), W, bi)
            elif subtyping:
                top_p = torch.index_select(h1, 0, torch.topk(A, k_sample)[1][-1])
                targets = torch.zeros(k_sample, dtype=torch.long)
                lg = F.linear(top_p, W, bi)
            else:
                continue
            total = total + loss_fn(lg, targets)
            preds_all.append(torch.topk(lg, 1, dim=1)[1].squeeze(1))
            targets_all.append(targets)
        if subtyping:
            total = total / n_classes
        res = {"instance_loss": total, "inst_labels": torch.cat(targets_all), "inst_preds": torch.cat(preds_all)}
    M = A @ h1
    logits = F.linear(M, sd["classifiers.weight"], sd["classifiers.bias"])
    return logits, F.softmax(logits, dim=1), torch.topk(logits, 1, dim=1)[1], A_raw, res
