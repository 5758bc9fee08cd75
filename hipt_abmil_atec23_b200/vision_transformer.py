"""ViT-256 (DINO ViT-S/16) with the reference's module tree and state_dict keys, executed by libhipt_b200.

Mirrors the public surface of HIPT_4K/vision_transformer.py in the reference (VisionTransformer :173-272, vit_tiny /
vit_small / vit_base :275-293): same constructor arguments, same parameter names and shapes (150 tensors for vit_small),
same random initialisation order, same `forward(x) -> CLS [B, embed_dim]`.  The nn.Linear / nn.LayerNorm / nn.Conv2d
children only HOLD the fp32 master parameters; `forward` hands a bf16-packed copy of them to the CUDA plan in
engine.py.  There is no PyTorch or CPU execution path: a CPU input raises.
"""
import math
from functools import partial

import torch
import torch.nn as nn

from . import engine


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)


class Attention(nn.Module):
    def __init__(self, dim, num_heads, qkv_bias):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads, qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.img_size = img_size
        self.patch_size = patch_size
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


def init_vit_weights(model):
    """Same RNG consumption as the reference: trunc_normal(pos_embed), trunc_normal(cls_token), then every nn.Linear in
    module.apply order gets trunc_normal(std=.02) weights and zero bias; LayerNorm = (1, 0).  The Conv2d keeps torch's
    default init (vision_transformer.py:200-211)."""
    nn.init.trunc_normal_(model.pos_embed, std=.02)
    nn.init.trunc_normal_(model.cls_token, std=.02)

    def _init(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    model.apply(_init)


def interpolate_pos_table(pos_embed, n_tokens, w0, h0):
    """[1, 1+N, dim] learned table -> [1+n_tokens, dim] for a w0 x h0 token grid, bicubic with the reference's
    scale_factor ((w0+0.1)/sqrt(N), (h0+0.1)/sqrt(N)) (vision_transformer.py:213-233).  Pure weight preprocessing, done
    once per grid shape in fp32 on the CPU."""
    pe = pos_embed.detach().float().cpu()
    N = pe.shape[1] - 1
    if n_tokens == N and w0 == h0:
        return pe[0].contiguous()
    dim = pe.shape[-1]
    s = int(math.sqrt(N))
    grid = pe[:, 1:].reshape(1, s, s, dim).permute(0, 3, 1, 2)
    grid = nn.functional.interpolate(grid, scale_factor=((w0 + 0.1) / math.sqrt(N), (h0 + 0.1) / math.sqrt(N)),
                                     mode="bicubic")
    if (int(w0 + 0.1), int(h0 + 0.1)) != tuple(grid.shape[-2:]):
        raise AssertionError("interpolated positional grid has the wrong shape")
    grid = grid.permute(0, 2, 3, 1).reshape(-1, dim)
    return torch.cat((pe[0, :1], grid), dim=0).contiguous()


class VisionTransformer(nn.Module):
    """Vision Transformer; forward runs on the B200 CUDA path only."""

    def __init__(self, img_size=[224], patch_size=16, in_chans=3, num_classes=0, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., norm_layer=nn.LayerNorm, **kwargs):
        super().__init__()
        if drop_rate or attn_drop_rate or drop_path_rate or qk_scale:
            raise NotImplementedError("the CUDA path implements the frozen eval configuration (no dropout / drop-path)")
        self.num_features = self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.patch_embed = PatchEmbed(img_size[0], patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        self.blocks = nn.ModuleList(
            [Block(embed_dim, num_heads, mlp_ratio, qkv_bias, norm_layer) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        init_vit_weights(self)
        self._engines = {}

    # ------------------------------------------------------------------------------------------- CUDA execution
    def _engine(self, device):
        return engine.get_engine(self, "vit256", device)

    def forward(self, x):
        """x: [B, 3, 256, 256] fp32 (already normalised) on a CUDA device -> final-LayerNorm CLS tokens [B, embed_dim]."""
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 256, 256):
            raise RuntimeError(f"ViT-256 CUDA path expects [B,3,256,256] inputs, got {tuple(x.shape)}")
        return self._engine(x.device).forward_patches(x)[0]

    def interpolate_pos_encoding(self, x, w, h):
        npatch = x.shape[1] - 1
        ps = self.patch_embed.patch_size
        return interpolate_pos_table(self.pos_embed, npatch, w // ps, h // ps).unsqueeze(0).to(x.device)

    @torch.no_grad()
    def get_last_selfattention(self, x):
        """[B, 3, 256, 256] -> attention probabilities of the last block [B, heads, 257, 257] (vision_transformer.py:255-262)."""
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 256, 256):
            raise RuntimeError(f"ViT-256 CUDA path expects [B,3,256,256] inputs, got {tuple(x.shape)}")
        eng = self._engine(x.device)
        outs = []
        for b0 in range(0, x.shape[0], eng.max_seqs):
            xb = x[b0:b0 + eng.max_seqs].contiguous()
            outs.append(eng.last_selfattention(lambda: eng.forward_patches(xb), xb.shape[0], 257))
        return torch.cat(outs)

    def get_intermediate_layers(self, x, n=1):
        raise NotImplementedError("intermediate-layer export is outside the accelerated hot path")


def vit_tiny(patch_size=16, **kwargs):
    return VisionTransformer(patch_size=patch_size, embed_dim=192, depth=12, num_heads=3, mlp_ratio=4, qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def vit_small(patch_size=16, **kwargs):
    return VisionTransformer(patch_size=patch_size, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4, qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def vit_base(patch_size=16, **kwargs):
    return VisionTransformer(patch_size=patch_size, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
