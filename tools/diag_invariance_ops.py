"""Which kernel's output for a row depends on the rows around it?  Each op runs on M = 2 x 65,792 rows and on the second
half alone; the second half must come out bit-identical."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hipt_abmil_atec23_b200 import _lib as L
H = 256 * 257
M = 2 * H
g = torch.Generator().manual_seed(0)
def r(shape, s=1.0): return (torch.randn(shape, generator=g) * s).cuda()
x = r((M, 384)); xb = x.bfloat16()
stats = torch.stack([x.sum(1), (x * x).sum(1)], 1).contiguous()
w_qkv = r((1152, 384), 0.05).bfloat16(); c = r((1152,), 0.1); d = r((1152,), 0.1)
o_full = L.gemm_lnfold_bf16(xb, w_qkv, c, d, stats, 1e-6)
o_half = L.gemm_lnfold_bf16(xb[H:].contiguous(), w_qkv, c, d, stats[H:].contiguous(), 1e-6)
print("lnfold qkv        :", torch.equal(o_full[H:], o_half))
qkv = o_full
a_full = L.attention(qkv, 512, 257, 6, 64, 0.125)
a_half = L.attention(qkv[H:].contiguous(), 256, 257, 6, 64, 0.125)
print("attention         :", torch.equal(a_full[H:], a_half), (a_full[H:].float() - a_half.float()).abs().max().item())
w_p = r((384, 384), 0.05).bfloat16(); b_p = r((384,), 0.1)
att = r((M, 384)).bfloat16()
res_full = xb.clone(); res_half = xb[H:].clone()
of, sf = L.gemm_resid_bf16(att, w_p, b_p, res_full)
oh, sh = L.gemm_resid_bf16(att[H:].contiguous(), w_p, b_p, res_half)
print("proj resid        :", torch.equal(of[H:], oh), "stats", torch.equal(sf[:, H:], sh))
w1 = r((1536, 384), 0.05).bfloat16(); c1 = r((1536,), 0.1); d1 = r((1536,), 0.1)
w2 = r((384, 1536), 0.02).bfloat16(); b2 = r((384,), 0.1)
planes = torch.zeros((6, M, 2), device="cuda"); planes[0] = stats
xf = xb.clone(); xh = xb[H:].clone()
s_full = L.mlp_fused_bf16(xf, w1, c1, d1, w2, b2, planes)
s_half = L.mlp_fused_bf16(xh, w1, c1, d1, w2, b2, planes[:, H:].contiguous())
print("mlp fused         :", torch.equal(xf[H:], xh), (xf[H:].float() - xh.float()).abs().max().item(), "stats", torch.equal(s_full[:, H:], s_half))
# patch-embed tokens GEMM
T = 256
a768 = r((2 * 65536, 768)).bfloat16(); we = r((384, 768), 0.03).bfloat16(); be = r((384,), 0.1); pos = r((257, 384))
out_f = torch.zeros((M, 384), device="cuda"); out_h = torch.zeros((H, 384), device="cuda")
L.gemm_bf16(a768, we, be, L.HB_EPI_TOKENS_F32, out=out_f, tok_table=pos, tokens_per_seq=T)
L.gemm_bf16(a768[65536:].contiguous(), we, be, L.HB_EPI_TOKENS_F32, out=out_h, tok_table=pos, tokens_per_seq=T)
print("tokens embed      :", torch.equal(out_f[H:], out_h))
