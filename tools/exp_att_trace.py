"""Dump the in-kernel event trace of CTA 0 of attention_tc_kernel (needs an HB_EXP_TRACE build: tools/build_exp.sh TRACE -DHB_EXP_TRACE,
then HB_LIB_PATH=hipt_abmil_atec23_b200/lib/exp_TRACE.so python tools/exp_att_trace.py)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from hipt_abmil_atec23_b200 import _lib as L
NSEQ = int(sys.argv[1]) if len(sys.argv) > 1 else 512
M = NSEQ * 257
qkv = (torch.randn((M, 1152), generator=torch.Generator().manual_seed(0))).cuda().bfloat16()
for _ in range(3): L.attention(qkv, NSEQ, 257, 6, 64, 0.125)
torch.cuda.synchronize()
buf = (C.c_longlong * 8192)()
assert L.load().hb_exp_read_att_trace(buf) == 0
t = np.array(buf[:], dtype=np.int64).reshape(16, 32, 16)
t = t - t[0, 0, 0]
for it in range(3, 8):
    print(f"== item {it}")
    print("  PROD   k_wait %6d k_go %6d v_wait %6d v_go %6d" % tuple(t[0, it, :4]))
    print("  TAIL   start %6d k_full %6d qk_done %6d v_ok %6d done %6d" % tuple(t[3, it, :5]))
    for x in range(2):
        n = 2 * it + x
        print(f"  MMA t{x}  start %6d kq %6d ofree %6d v %6d | atoms " % tuple(t[1, n, :4]) + " ".join("%6d" % v for v in t[1, n, 4:9]))
        print(f"  SM  t{x}  start %6d kq %6d dot %6d s_full %6d max %6d bar %6d | P " % tuple(t[4, n, :6]) + " ".join("%6d" % v for v in t[4, n, 6:10])
              + " | o_wait %6d o_full %6d stored %6d" % tuple(t[4, n, 10:13]))
