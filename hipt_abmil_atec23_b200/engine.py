"""Per-(module, device) execution state of a ViT on the CUDA path: bf16-packed weights, workspace, plan.

The nn.Module owns the fp32 master parameters (identical keys and shapes to the reference); this file derives what
libhipt_b200 needs from them — bf16 copies of the GEMM weights (nn.Linear.weight is already the [N, K] K-major operand
the tcgen05 kernel wants), the interpolated positional tables, the patch-embed weights with the input normalisation
folded in — and owns the workspace and the hb_vit_plan.  State is rebuilt when any parameter storage or version changes
(load_state_dict, .to(), optimiser steps).
"""
import ctypes as C
import threading

import torch

from . import _lib

_lock = threading.Lock()

import os

# default capacities (token sequences per launch)
# Sequences (256 x 256 patches) per ViT-256 launch sequence.  1024 = four 4096x4096 regions = 1,028 CTA-pair tiles = 13.9 waves
# over the 74 SM pairs (99 % full) and a 1.84 GB workspace.  Measured on one box, 16 regions, serpentine tile order
# (tools/exp_group_size.py, profiles/r02d_group_size.jsonl): 221 -> 274.5, 294 -> 278.0, 512 -> 285.0, 1024 -> 290.8,
# 2048 -> 292.9, 4096 -> 293.8 regions/s: fewer launch tails beat the smaller L2 footprint of short groups.
VIT256_MAX_PATCHES = int(os.environ.get("HB_VIT256_MAX_PATCHES", "1024"))
VIT4K_MAX_REGIONS = 64


def _signature(module):
    return tuple((p.data_ptr(), p._version) for p in module.parameters())


def get_engine(module, kind, device):
    """Return (building if needed) the engine of `module` on `device`."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"{type(module).__name__} runs only on CUDA (B200): there is no CPU path; got device '{device}'")
    key = device.index if device.index is not None else torch.cuda.current_device()
    with _lock:
        cache = module.__dict__.setdefault("_engines", {})
        eng = cache.get(key)
        sig = _signature(module)
        if eng is None or eng.signature != sig:
            first = next(module.parameters())
            if first.device.type != "cuda" or (first.device.index or 0) != key:
                raise RuntimeError(f"{type(module).__name__} parameters live on {first.device}, input on {device}")
            eng = VitEngine(module, kind, torch.device("cuda", key))
            eng.signature = sig
            cache[key] = eng
        return eng


class VitEngine:
    def __init__(self, module, kind, device, max_seqs=None):
        self.kind = kind
        self.device = device
        self.lib = _lib.load()
        with torch.cuda.device(device):
            _lib.device_check()
        self.dim = module.embed_dim
        self.heads = module.num_heads
        self.depth = len(module.blocks)
        self.mlp_dim = module.blocks[0].mlp.fc1.out_features
        self.eps = float(module.norm.eps)
        if kind == "vit256":
            self.seq_len = 257
            self.max_seqs = max_seqs or VIT256_MAX_PATCHES
        else:
            self.seq_len = None                      # 1 + w*h, per call
            self.max_seqs = max_seqs or VIT4K_MAX_REGIONS
        self.max_rows = self.max_seqs * 257
        self._module = module
        # one workspace per (module, device): concurrent callers (threads of nn.DataParallel sharing a device, user threads)
        # take turns; launches go to the caller's current stream, and the next caller's stream waits for the previous one
        self._run_lock = threading.RLock()
        self._last_stream = None
        self._pos_cache = {}
        self._embed_cache = {}
        self._pack(module)
        self._build_plan()

    # -------------------------------------------------------------------------------------------------- weights
    def _pack(self, m):
        dev = self.device
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        bf16 = lambda t: t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
        w = [f32(m.cls_token).reshape(-1), f32(m.norm.weight), f32(m.norm.bias)]

        def fold(norm, lin):
            """LayerNorm folded into the Linear that consumes it (hb_gemm_lnfold_bf16): bf16(W * gamma), the row sums c
            of that ROUNDED weight (so the mean term cancels exactly against the MMA), and d = W beta + bias."""
            W = lin.weight.detach().double().cpu()
            g = norm.weight.detach().double().cpu()
            be = norm.bias.detach().double().cpu()
            b = lin.bias.detach().double().cpu() if lin.bias is not None else torch.zeros(W.shape[0], dtype=torch.float64)
            wg = (W * g[None, :]).to(torch.bfloat16)
            c = wg.double().sum(dim=1)
            d = W @ be + b
            return wg.to(dev).contiguous(), c.float().to(dev).contiguous(), d.float().to(dev).contiguous()

        for blk in m.blocks:
            qw, qc, qd = fold(blk.norm1, blk.attn.qkv)
            fw, fc, fd = fold(blk.norm2, blk.mlp.fc1)
            w += [qw, qc, qd, bf16(blk.attn.proj.weight), f32(blk.attn.proj.bias),
                  fw, fc, fd, bf16(blk.mlp.fc2.weight * 0.5), f32(blk.mlp.fc2.bias)]      # fc1 epilogue emits 2 * GELU
        self.weights = w                              # keeps the device copies alive
        if self.kind == "vit4k":
            self.phi_w = bf16(m.phi[0].weight)
            self.phi_b = f32(m.phi[0].bias)
            self.in_dim = m.phi[0].in_features

    def _build_plan(self):
        cfg = _lib.HbVitConfig(self.dim, self.heads, self.depth, self.mlp_dim, self.max_rows, self.eps)
        nbytes = self.lib.hb_vit_workspace_bytes(C.byref(cfg))
        with torch.cuda.device(self.device):
            raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
        off = (-raw.data_ptr()) % 1024
        self._ws_raw = raw
        self.workspace_bytes = nbytes
        arr = (C.c_void_p * len(self.weights))(*[t.data_ptr() for t in self.weights])
        plan = C.c_void_p()
        _lib.check(self.lib.hb_vit_plan_create(C.byref(cfg), arr, len(self.weights), C.c_void_p(raw.data_ptr() + off),
                                               nbytes, C.byref(plan)))
        self.plan = plan

    def __del__(self):
        plan = getattr(self, "plan", None)
        if plan is not None and plan.value:
            try:
                self.lib.hb_vit_plan_destroy(plan)
            except Exception:
                pass
            self.plan = None

    # ---------------------------------------------------------------------------------------- derived tables
    def pos_table(self, w0, h0):
        """[1 + w0*h0, dim] fp32 cls+pos rows for a w0 x h0 token grid (cached per grid shape)."""
        key = (w0, h0)
        t = self._pos_cache.get(key)
        if t is None:
            from .vision_transformer import interpolate_pos_table
            t = interpolate_pos_table(self._module.pos_embed, w0 * h0, w0, h0).to(self.device)
            self._pos_cache[key] = t
        return t

    def embed_weights(self, mean=None, std=None):
        """patch_embed.proj as the [384, 768] bf16 GEMM operand.  With (mean, std) given, the input is RAW uint8 pixel
        values p and ToTensor + Normalize ((p/255 - mean_c)/std_c, hipt_model_utils.py:113-118 or
        datasets/wsi_dataset.py:12-16) is folded in:  W'[:, c] = W[:, c] / (255 std_c),  b' = b - sum_c W[:, c] mean_c/std_c."""
        key = None if mean is None else (tuple(float(v) for v in mean), tuple(float(v) for v in std))
        ent = self._embed_cache.get(key)
        if ent is None:
            proj = self._module.patch_embed.proj
            W = proj.weight.detach().double().cpu()                 # [384, 3, 16, 16], K order (c, i, j)
            b = proj.bias.detach().double().cpu()
            if key is not None:
                m = torch.tensor(key[0], dtype=torch.float64).view(1, 3, 1, 1)
                s = torch.tensor(key[1], dtype=torch.float64).view(1, 3, 1, 1)
                b = b - (W * (m / s)).sum(dim=(1, 2, 3))
                W = W / (255.0 * s)
            ent = (W.reshape(W.shape[0], -1).to(torch.bfloat16).to(self.device).contiguous(),
                   b.float().to(self.device).contiguous())
            self._embed_cache[key] = ent
        return ent

    def embed_weights_f16(self, mean, std):
        """Operands of the fused uint8 patch embed (hb_vit256_forward_u8): fp16(W / std_c) as [384, 768] in K order (c, i, j),
        bias b - sum_k fp16(W / std)[k] mean_c (the ROUNDED weights, so the mean term cancels exactly against the MMA), and the
        accumulator scale 1/255 (ToTensor).  fp16 keeps 11 significant bits of the weights (bf16: 8)."""
        key = ("f16", tuple(float(v) for v in mean), tuple(float(v) for v in std))
        ent = self._embed_cache.get(key)
        if ent is None:
            proj = self._module.patch_embed.proj
            W = proj.weight.detach().double().cpu()
            b = proj.bias.detach().double().cpu()
            m = torch.tensor(key[1], dtype=torch.float64).view(1, 3, 1, 1)
            s = torch.tensor(key[2], dtype=torch.float64).view(1, 3, 1, 1)
            wr = (W / s).to(torch.float16)
            if not torch.isfinite(wr).all():
                raise RuntimeError("patch-embed weights overflow fp16 after folding the normalisation")
            bias = b - (wr.double() * m).sum(dim=(1, 2, 3))
            ent = (wr.reshape(W.shape[0], -1).to(self.device).contiguous(), bias.float().to(self.device).contiguous(), 1.0 / 255.0)
            self._embed_cache[key] = ent
        return ent

    # ------------------------------------------------------------------------------------------------ forwards
    def _order_streams(self):
        """The workspace is shared: work queued on another stream by the previous caller must finish first."""
        cur = torch.cuda.current_stream(self.device)
        if self._last_stream is not None and self._last_stream != cur:
            cur.wait_stream(self._last_stream)
        self._last_stream = cur

    def _set_cls_attention(self, buf):
        _lib.check(self.lib.hb_vit_plan_set_cls_attention(self.plan, _lib.ptr(buf)))

    def forward_patches(self, image, patch_begin=0, n_patches=None, mean=None, std=None, want_f32=True, out_bf16=None,
                        cls_attn=None):
        """ViT-256 over patches of `image` (region [3,H,W], region batch [R,3,H,W] or patch batch [B,3,256,256]; fp32
        normalised, or uint8 with mean/std).  Returns (cls_f32 [n, dim] or None, cls_bf16 [n, dim]).
        cls_attn: optional fp32 [n, heads, 257] CUDA tensor that receives the softmax row of the CLS query of the last block
        (get_last_selfattention(...)[:, :, 0, :], vision_transformer.py:255-262) from the fused attention launch."""
        assert self.kind == "vit256"
        _lib.require_cuda(image, "image")
        ps, cs, rp, gc, total, ppi, istride = _lib.image_layout(image)
        n = total - patch_begin if n_patches is None else n_patches
        is_f32 = image.dtype == torch.float32
        if not is_f32 and mean is None:
            raise RuntimeError("uint8 input needs the (mean, std) of the normalisation to fold into the patch embed")
        # uint8 regions: unfold + normalise + patch embed are one tensor-core kernel reading the bytes from HBM
        fused = (not is_f32 and gc > 0 and self.dim == 384 and os.environ.get("HB_EMBED_UNFUSED") != "1"
                 and cs % 16 == 0 and rp % 16 == 0 and istride % 16 == 0 and image.data_ptr() % 16 == 0)
        if fused:
            ew, eb, escale = self.embed_weights_f16(mean, std)
        else:
            ew, eb = self.embed_weights(None if is_f32 else mean, None if is_f32 else std)
        pos = self.pos_table(16, 16)
        with self._run_lock, torch.cuda.device(self.device):
            self._order_streams()
            cls_f32 = torch.empty((n, self.dim), dtype=torch.float32, device=self.device) if want_f32 else None
            cls_bf16 = out_bf16 if out_bf16 is not None else torch.empty((n, self.dim), dtype=torch.bfloat16,
                                                                         device=self.device)
            if cls_attn is not None:
                assert cls_attn.is_cuda and cls_attn.dtype == torch.float32 and cls_attn.is_contiguous()
                assert tuple(cls_attn.shape) == (n, self.heads, 257)
            done = 0
            while done < n:                         # minibatches of the plan capacity (hipt_4k.py:68-70)
                cur = min(self.max_seqs, n - done)
                if cls_attn is not None:
                    self._set_cls_attention(cls_attn[done:])
                if fused:
                    _lib.check(self.lib.hb_vit256_forward_u8(
                        self.plan, _lib.ptr(image), cs, rp, gc, ppi // gc, istride, total // ppi, patch_begin + done, cur,
                        _lib.ptr(ew), _lib.ptr(eb), escale, _lib.ptr(pos),
                        _lib.ptr(cls_f32[done:] if cls_f32 is not None else None), _lib.ptr(cls_bf16[done:]), _lib.stream_ptr()))
                    done += cur
                    continue
                _lib.check(self.lib.hb_vit256_forward(
                    self.plan, _lib.ptr(image), int(is_f32), ps, cs, rp, gc, ppi, istride, patch_begin + done, cur, _lib.ptr(ew),
                    _lib.ptr(eb), _lib.ptr(pos), _lib.ptr(cls_f32[done:] if cls_f32 is not None else None),
                    _lib.ptr(cls_bf16[done:]), _lib.stream_ptr()))
                done += cur
            if cls_attn is not None:
                self._set_cls_attention(None)
        return cls_f32, cls_bf16

    def forward_grid(self, cls256_bf16, n_regions, w0, h0, out=None, cls_attn=None):
        """ViT-4K over n_regions grids of w0*h0 ViT-256 CLS tokens: [n_regions*w0*h0, in_dim] bf16 -> [n_regions, dim]
        (written into `out` [>= n_regions, dim] fp32 when given: the slide-set pass pools straight out of that buffer)."""
        assert self.kind == "vit4k"
        _lib.require_cuda(cls256_bf16, "cls256")
        T = w0 * h0
        assert cls256_bf16.dtype == torch.bfloat16 and cls256_bf16.is_contiguous()
        assert cls256_bf16.shape[0] >= n_regions * T and cls256_bf16.shape[1] == self.in_dim
        pos = self.pos_table(w0, h0)
        cap = max(1, self.max_rows // (T + 1))
        with self._run_lock, torch.cuda.device(self.device):
            self._order_streams()
            if out is None:
                out = torch.empty((n_regions, self.dim), dtype=torch.float32, device=self.device)
            else:
                assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.shape[1] == self.dim
                assert out.shape[0] >= n_regions
            if cls_attn is not None:                  # [n_regions, heads, T + 1] fp32: CLS-query softmax row of the last block
                assert cls_attn.is_cuda and cls_attn.dtype == torch.float32 and cls_attn.is_contiguous()
                assert tuple(cls_attn.shape) == (n_regions, self.heads, T + 1)
            done = 0
            while done < n_regions:
                cur = min(cap, n_regions - done)
                if cls_attn is not None:
                    self._set_cls_attention(cls_attn[done:])
                _lib.check(self.lib.hb_vit4k_forward(
                    self.plan, _lib.ptr(cls256_bf16[done * T:]), cur, T, self.in_dim, _lib.ptr(self.phi_w),
                    _lib.ptr(self.phi_b), _lib.ptr(pos), _lib.ptr(out[done:]), _lib.stream_ptr()))
                done += cur
            if cls_attn is not None:
                self._set_cls_attention(None)
        return out[:n_regions]

    # ------------------------------------------------------------------------------- attention maps (heatmap helpers)
    def last_selfattention(self, run_prefix, n_seq, seq_len):
        """Attention probabilities of the LAST block, [n_seq, heads, seq_len, seq_len] fp32 (get_last_selfattention,
        vision_transformer.py:255-262 / vision_transformer4k.py:248-255; consumer: hipt_heatmap_utils.py:328-335).
        `run_prefix()` launches the forward with the plan limited to blocks 0 .. depth-2, which leaves the residual stream
        entering the last block in the workspace; norm1, the qkv Linear and the softmax of that one block are torch ops on
        its rows (this is the visualisation path, SURVEY.md §8f rank 4, not the extraction hot path)."""
        import torch.nn.functional as F
        blk = self._module.blocks[-1]
        with self._run_lock:                          # the depth limit is plan state: no other forward may interleave
            if self.depth > 1:
                self.set_depth_limit(self.depth - 1)
            try:
                run_prefix()
            finally:
                self.set_depth_limit(0)
            return self._last_block_attention(blk, n_seq, seq_len, F)

    def _last_block_attention(self, blk, n_seq, seq_len, F):
        if self.depth == 1:
            raise NotImplementedError("attention-map export needs at least two blocks")
        x = self.buffer(1, n_seq * seq_len, self.dim, torch.bfloat16).float().view(n_seq, seq_len, self.dim)
        y = F.layer_norm(x, (self.dim,), blk.norm1.weight.float(), blk.norm1.bias.float(), self.eps)
        qkv = F.linear(y, blk.attn.qkv.weight.float(), blk.attn.qkv.bias.float() if blk.attn.qkv.bias is not None else None)
        hd = self.dim // self.heads
        qkv = qkv.reshape(n_seq, seq_len, 3, self.heads, hd).permute(2, 0, 3, 1, 4)
        att = (qkv[0] @ qkv[1].transpose(-2, -1)) * (hd ** -0.5)
        return att.softmax(dim=-1)

    # -------------------------------------------------------------------------------------------- test hooks
    def set_depth_limit(self, n):
        _lib.check(self.lib.hb_vit_plan_set_depth_limit(self.plan, n))

    def buffer(self, which, rows, cols, dtype):
        """View of a workspace buffer (1 bf16 residual stream, 2 qkv, 3 attention out, 4 hidden) as [rows, cols]."""
        p, nb = C.c_void_p(), C.c_size_t()
        _lib.check(self.lib.hb_vit_plan_buffer(self.plan, which, C.byref(p), C.byref(nb)))
        off = p.value - self._ws_raw.data_ptr()
        esz = torch.empty((), dtype=dtype).element_size()
        return self._ws_raw[off:off + rows * cols * esz].view(dtype).view(rows, cols)
