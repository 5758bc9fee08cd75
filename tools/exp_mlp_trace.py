import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from hipt_abmil_atec23_b200 import _lib as L
M = 256 * 257
g = torch.Generator().manual_seed(0)
def r(shape, s=1.0): return (torch.randn(shape, generator=g) * s).cuda()
x = r((M, 384)); xb = x.bfloat16()
planes = torch.zeros((6, M, 2), device="cuda"); planes[0] = torch.stack([x.sum(1), (x * x).sum(1)], 1)
w1 = r((1536, 384), 0.05).bfloat16(); c1 = r((1536,), 0.1); d1 = r((1536,), 0.1)
w2 = r((384, 1536), 0.02).bfloat16(); b2 = r((384,), 0.1)
for _ in range(3): L.mlp_fused_bf16(xb, w1, c1, d1, w2, b2, planes)
torch.cuda.synchronize()
buf = (C.c_longlong * 8192)()
assert L.load().hb_exp_read_mlp_trace(buf) == 0
t = np.array(buf[:], dtype=np.int64)
t0 = t[0]
m = t[:4096].reshape(32, 32, 4) - t0
e = t[4096:].reshape(32, 128) - t0
for ti in range(4):
    print(f"== tile {ti}")
    for j in range(26):
        print(f" j{j:2d} S: wait {m[ti,j,0]:7d} issue {m[ti,j,1]:7d} | O: wait {m[ti,j,2]:7d} issue {m[ti,j,3]:7d}")
    for k in range(6):
        print(f"   WG0 chunk {4*k:2d}: s_wait {e[ti,k*8]:7d} got {e[ti,k*8+1]:7d} math_done {e[ti,k*8+2]:7d} h_free {e[ti,k*8+3]:7d} published {e[ti,k*8+4]:7d}")
    print(f"   WG0 gelu loop end {e[ti,64]:7d} o_full {e[ti,65]:7d} tile epilogue done {e[ti,66]:7d}")
