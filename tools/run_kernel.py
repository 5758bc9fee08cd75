"""Run one kernel a few times on seeded data (for ncu captures): python tools/run_kernel.py attention|gemm_qkv|gemm_fc1|gemm_fc2|gemm_proj|layernorm"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hipt_abmil_atec23_b200 import _lib as L

which = sys.argv[1] if len(sys.argv) > 1 else "attention"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
NSEQ = int(sys.argv[3]) if len(sys.argv) > 3 else 256
M = NSEQ * 257
g = torch.Generator().manual_seed(0)
def r(shape, s=1.0): return (torch.randn(shape, generator=g) * s).cuda()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if which == "attention":
    qkv = r((M, 1152)).bfloat16()
    f = lambda: L.attention(qkv, NSEQ, 257, 6, 64, 0.125)
elif which.startswith("gemm_"):
    N, K, epi = {"gemm_qkv": (1152, 384, 0), "gemm_fc1": (1536, 384, 5), "gemm_fc2": (384, 1536, 2), "gemm_proj": (384, 384, 2)}[which]
    a = r((M, K)).bfloat16(); w = r((N, K), 0.05).bfloat16(); b = r((N,), 0.1)
    out = torch.zeros((M, N), device="cuda", dtype=torch.float32 if epi == 2 else torch.bfloat16)
    f = lambda: L.gemm_bf16(a, w, b, epi, out=out)
elif which == "layernorm":
    x = r((M, 384)); gm = r((384,)); bt = r((384,))
    f = lambda: L.layernorm(x, gm, bt, 1e-6, M, 384)
for _ in range(3): f()
torch.cuda.synchronize()
ev0.record()
for _ in range(reps): f()
ev1.record(); torch.cuda.synchronize()
print(which, "avg us", 1000 * ev0.elapsed_time(ev1) / reps)
