"""Dump the in-kernel event trace (HB_EXP_TRACE build) of CTA 0 of the tcgen05 attention kernel: cycles per phase of the
softmax warpgroups (warp 4 / 8) and the epilogue warpgroup (warp 12).
HB_LIB_PATH=hipt_abmil_atec23_b200/lib/exp_NAME.so python tools/exp_at2_trace.py [n_seq]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from hipt_abmil_atec23_b200 import _lib as L
NSEQ = int(sys.argv[1]) if len(sys.argv) > 1 else 512
qkv = (torch.randn((NSEQ * 257, 1152), generator=torch.Generator().manual_seed(0))).cuda().bfloat16()
f = lambda: L.attention(qkv, NSEQ, 257, 6, 64, 0.125)
for _ in range(3): f()
torch.cuda.synchronize()
f(); torch.cuda.synchronize()
buf = (C.c_longlong * (4 * 64 * 16))()
lib = L.load()
assert lib.hb_exp_read_at2_trace(buf) == 0
t = np.array(buf[:], dtype=np.int64).reshape(4, 64, 16)
t0 = t[0, 0, 0]
names = ["kvq+dot", "sA", "unitA", "sB", "unitB", "stats"]
for w in (0, 1):
    print(f"--- softmax warpgroup {w} (one warp): item start, then deltas [{' '.join(names)}], item total")
    for j in range(2, 14):
        r = t[w, j]
        d = [int(r[k] - r[k - 1]) for k in range(1, 7)]
        print(f"item {j:2d} start {int(r[0] - t0):8d} | " + " ".join(f"{x:6d}" for x in d) + f" | total {int(t[w, j + 1, 0] - r[0]):6d}")
print("--- epilogue warpgroup (warp 12): absolute [O_a tile0 drained, tile0 stored, O_a tile1 drained, tile1 stored]")
for j in range(2, 14):
    r = t[2, j]
    print(f"item {j:2d} " + " ".join(f"{int(r[k] - t0):8d}" for k in range(4)) + f" | item period {int(t[2, j + 1, 0] - r[0]):6d}")
