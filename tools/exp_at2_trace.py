"""Dump the in-kernel event trace (HB_EXP_TRACE build) of CTA 0 of the tcgen05 attention kernel: cycles relative to the
first stamp, per warpgroup tile.  HB_LIB_PATH=hipt_abmil_atec23_b200/lib/exp_NAME.so python tools/exp_at2_trace.py [n_seq]"""
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
names = ["start", "kvq", "sA", "pA", "sB", "pB", "oA", "ldA", "oB", "ldB", "bar", "x"]
for w in (0, 1):
    print(f"--- softmax warpgroup {w}: per tile, start then deltas between successive stamps [{' '.join(names[1:11])}]")
    for j in range(2, 14):
        r = t[w, j]
        d = [int(r[k] - r[k - 1]) for k in range(1, 11)]
        print(f"tile {j:2d} start {int(r[0] - t0):8d} | " + " ".join(f"{x:6d}" for x in d) + f" | tile total {int(t[w, j + 1, 0] - r[0]):6d}")
    print(f"    unit A inner: [ld done, max done, exp done] relative to the sA stamp")
    for j in range(2, 8):
        r = t[w, j]
        print(f"tile {j:2d} " + " ".join(f"{int(r[k] - r[2]):6d}" for k in (11, 12, 13)) + f"   (pA {int(r[3]-r[2])})")
    print(f"--- MMA warp {w}: [wait pA, issue PV A, wait pB, issue PV B, wait freeA, issue S A', wait freeB, issue S B'] absolute; then observed completion [S A, S B, O A, O B] (warpgroup 0 only)")
    for j in range(2, 10):
        r = t[2 + w, j]
        print(f"tile {j:2d} " + " ".join(f"{int(r[k] - t0):8d}" for k in range(8)) + " | " + " ".join(f"{int(r[k] - t0):8d}" for k in range(8, 12)))
