"""Time the four ViT-256 block GEMMs as the plan launches them (LN-folded qkv / fc1+GELU, residual proj / fc2) on one
region's rows; HB_LIB_PATH selects an experimental build.  Not part of the product path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hipt_abmil_atec23_b200 import _lib as L

M = int(os.environ.get("EXP_M", 256 * 257))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
g = torch.Generator().manual_seed(0)
def r(shape, s=1.0): return (torch.randn(shape, generator=g) * s).cuda()
x = r((M, 384)); xb = x.bfloat16()
stats = torch.stack([x.sum(1), (x * x).sum(1)], 1).contiguous()
hid = r((M, 1536)).bfloat16(); att = r((M, 384)).bfloat16()
out = {}
def timeit(name, f, flops):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    us = 1000 * e0.elapsed_time(e1) / reps
    print(f"{name}: {us:.1f} us  {flops / us / 1e6:.0f} TFLOP/s", flush=True)
only = os.environ.get("EXP_ONLY")
for name, N, K, kind in (("qkv", 1152, 384, "ln"), ("fc1", 1536, 384, "lng"), ("proj", 384, 384, "res"), ("fc2", 384, 1536, "res")):
    if only and name not in only.split(","): continue
    w = r((N, K), 0.05).bfloat16(); b = r((N,), 0.1); c = r((N,), 0.1)
    if kind in ("ln", "lng"):
        f = lambda: L.gemm_lnfold_bf16(xb, w, c, b, stats, 1e-6, gelu=(2 if kind == "lng" else 0))
    else:
        a = att if K == 384 else hid
        f = lambda: L.gemm_resid_bf16(a, w, b, xb)
    timeit(name, f, 2.0 * M * N * K)
if not only or "mlp" in only.split(","):
    w1 = r((1536, 384), 0.05).bfloat16(); c1 = r((1536,), 0.1); d1 = r((1536,), 0.1)
    w2 = r((384, 1536), 0.02).bfloat16(); b2 = r((384,), 0.1)
    planes = torch.zeros((6, M, 2), device="cuda"); planes[0] = stats
    timeit("mlp_fused", lambda: L.mlp_fused_bf16(xb, w1, c1, d1, w2, b2, planes), 4.0 * M * 1536 * 384)
