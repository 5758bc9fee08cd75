"""ViT-4K (region-level aggregator over the 16x16 grid of ViT-256 CLS tokens), executed by libhipt_b200.

Mirrors HIPT_4K/vision_transformer4k.py in the reference (VisionTransformer4K :161-265, vit4k_xs :267-272): same
constructor arguments, the same 78 state_dict tensors (`phi.0.*`, `cls_token`, `pos_embed`, `blocks.i.*`, `norm.*`) and
the same initialisation order.  `forward(x)` takes the [B, 384, w, h] CLS grid and returns the [B, 192] region token.
"""
from functools import partial

import torch
import torch.nn as nn

from . import engine
from .vision_transformer import Block, init_vit_weights, interpolate_pos_table


class VisionTransformer4K(nn.Module):
    def __init__(self, num_classes=0, img_size=[224], input_embed_dim=384, output_embed_dim=192, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., norm_layer=nn.LayerNorm, num_prototypes=64, **kwargs):
        super().__init__()
        if drop_rate or attn_drop_rate or drop_path_rate or qk_scale:
            raise NotImplementedError("the CUDA path implements the frozen eval configuration (no dropout / drop-path)")
        embed_dim = output_embed_dim
        self.num_features = self.embed_dim = embed_dim
        self.input_embed_dim = input_embed_dim
        self.num_heads = num_heads
        # Sequential(Linear, GELU, Dropout) in the reference: indices kept so the key stays `phi.0.*`
        self.phi = nn.Sequential(nn.Linear(input_embed_dim, output_embed_dim), nn.GELU(), nn.Dropout(p=drop_rate))
        num_patches = int(img_size[0] // 16) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.blocks = nn.ModuleList(
            [Block(embed_dim, num_heads, mlp_ratio, qkv_bias, norm_layer) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        init_vit_weights(self)
        self._engines = {}
        self.mpp_feature = None

    def _engine(self, device):
        return engine.get_engine(self, "vit4k", device)

    def forward(self, x):
        """x: [B, 384, w, h] fp32 grid of ViT-256 CLS tokens on a CUDA device -> [B, 192]."""
        if x.dim() != 4 or x.shape[1] != self.input_embed_dim:
            raise RuntimeError(f"ViT-4K CUDA path expects [B,{self.input_embed_dim},w,h] inputs, got {tuple(x.shape)}")
        self.mpp_feature = x                                    # kept for parity with prepare_tokens (:225)
        B, C, w, h = x.shape
        tokens = x.flatten(2, 3).transpose(1, 2).reshape(B * w * h, C)          # [B*T, 384], token t = row-major (w,h)
        return self._engine(x.device).forward_grid(tokens.to(torch.bfloat16).contiguous(), B, w, h)

    def interpolate_pos_encoding(self, x, w, h):
        return interpolate_pos_table(self.pos_embed, x.shape[1] - 1, w, h).unsqueeze(0).to(x.device)

    @torch.no_grad()
    def get_last_selfattention(self, x):
        """[B, 384, w, h] -> attention probabilities of the last block [B, heads, 1+w*h, 1+w*h] (vision_transformer4k.py:248-255)."""
        if x.dim() != 4 or x.shape[1] != self.input_embed_dim:
            raise RuntimeError(f"ViT-4K CUDA path expects [B,{self.input_embed_dim},w,h] inputs, got {tuple(x.shape)}")
        B, C, w, h = x.shape
        eng = self._engine(x.device)
        cap = max(1, eng.max_rows // (w * h + 1))
        outs = []
        for b0 in range(0, B, cap):
            xb = x[b0:b0 + cap]
            n = xb.shape[0]
            tokens = xb.flatten(2, 3).transpose(1, 2).reshape(n * w * h, C).to(torch.bfloat16).contiguous()
            outs.append(eng.last_selfattention(lambda: eng.forward_grid(tokens, n, w, h), n, w * h + 1))
        return torch.cat(outs)

    def get_intermediate_layers(self, x, n=1):
        raise NotImplementedError("intermediate-layer export is outside the accelerated hot path")


def vit4k_xs(patch_size=16, **kwargs):
    return VisionTransformer4K(patch_size=patch_size, input_embed_dim=384, output_embed_dim=192, depth=6, num_heads=6,
                               mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def count_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
