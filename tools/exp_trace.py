"""Dump the in-kernel event trace (HB_EXP_TRACE build) of CTA 0 for one fc1 / qkv launch."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hipt_abmil_atec23_b200 import _lib as L
which = sys.argv[1] if len(sys.argv) > 1 else "fc1"
M = 256 * 257
g = torch.Generator().manual_seed(0)
def r(shape, s=1.0): return (torch.randn(shape, generator=g) * s).cuda()
x = r((M, 384)); xb = x.bfloat16()
stats = torch.stack([x.sum(1), (x * x).sum(1)], 1).contiguous()
N = 1536 if which == "fc1" else 1152
w = r((N, 384), 0.05).bfloat16(); b = r((N,), 0.1); c = r((N,), 0.1)
f = lambda: L.gemm_lnfold_bf16(xb, w, c, b, stats, 1e-6, gelu=(2 if which == "fc1" else 0))
for _ in range(3): f()
torch.cuda.synchronize()
f(); torch.cuda.synchronize()
buf = (C.c_longlong * 4096)()
lib = L.load()
assert lib.hb_exp_read_trace(buf) == 0
import numpy as np
t = np.array(buf[:], dtype=np.int64).reshape(4, 128, 8)
t0 = t[0, 0, 0]
names = ["MMA", "EPIa", "EPIb", "PROD"]
for it in range(22):
    print(f"it {it:2d} | MMA wait {t[0,it,0]-t0:7d} ->{t[0,it,1]-t0:7d} kb0 {t[0,it,2]-t0:7d} last {t[0,it,3]-t0:7d}"
          f" | EPI0 wait {t[1,it,0]-t0:7d} full {t[1,it,1]-t0:7d} ld {t[1,it,2]-t0:7d} math {t[1,it,3]-t0:7d} drain {t[1,it,4]-t0:7d} bar {t[1,it,5]-t0:7d} sts {t[1,it,6]-t0:7d} st {t[1,it,7]-t0:7d}"
          f" | EPI1 full {t[2,it,1]-t0:7d} math {t[2,it,3]-t0:7d} st {t[2,it,7]-t0:7d} | PROD {t[3,it,0]-t0:7d} {t[3,it,1]-t0:7d}")
