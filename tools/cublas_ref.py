"""cuBLAS (torch.matmul bf16) on the four ViT-256 GEMM shapes, for context next to our kernels (not part of the product path)."""
import torch
M = 256 * 257
for name, N, K in (("qkv", 1152, 384), ("proj", 384, 384), ("fc1", 1536, 384), ("fc2", 384, 1536), ("big", 8192, 8192)):
    m = 8192 if name == "big" else M
    a = torch.randn(m, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    for _ in range(5): torch.matmul(a, w.t())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): torch.matmul(a, w.t())
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 20
    print(f"{name}: {us:.1f} us  {2*m*N*K/us/1e6:.0f} TFLOP/s")
