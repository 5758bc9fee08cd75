"""Power and clock of each hot kernel run alone in a tight loop (NVML samples every 50 ms for ~2.5 s per kernel):
energy per launch = mean power x mean launch time.  Under sw_power_cap the step is energy-limited, so this says which kernel's
ENERGY (not time) to cut.  python tools/power_probe.py"""
import json, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pynvml
from hipt_abmil_atec23_b200 import _lib as L

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
NSEQ = 512
M = NSEQ * 257
g = torch.Generator().manual_seed(0)
def r(shape, s=1.0): return (torch.randn(shape, generator=g) * s).cuda()

def sample(stop, out):
    while not stop.is_set():
        out.append((pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
        time.sleep(0.05)

def probe(name, fn, secs=2.5):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    stop, samples = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, samples)); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(50): fn()
        n += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    us = 1000 * e0.elapsed_time(e1) / n
    s = samples[len(samples) // 3:]
    p = sum(x[0] for x in s) / len(s); c = sum(x[1] for x in s) / len(s)
    print(json.dumps({"kernel": name, "us_per_launch": round(us, 1), "power_w": round(p, 1), "sm_mhz": round(c), "mJ_per_launch": round(p * us / 1000, 2)}))

idle = []
stop = threading.Event(); th = threading.Thread(target=sample, args=(stop, idle)); th.start(); time.sleep(1.0); stop.set(); th.join()
print(json.dumps({"kernel": "idle", "power_w": round(sum(x[0] for x in idle) / len(idle), 1)}))
qkv = r((M, 1152)).bfloat16()
probe("attention (2 regions)", lambda: L.attention(qkv, NSEQ, 257, 6, 64, 0.125))
a384 = r((M, 384)).bfloat16()
def gemm(N, K, epi, A):
    w = r((N, K), 0.05).bfloat16(); b = r((N,), 0.1)
    out = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16)
    return lambda: L.gemm_bf16(A, w, b, epi, out=out)
probe("qkv-shaped GEMM 384->1152 bias epilogue", gemm(1152, 384, L.HB_EPI_BIAS_BF16, a384))
probe("proj-shaped GEMM 384->384 bias epilogue", gemm(384, 384, L.HB_EPI_BIAS_BF16, a384))
# fused MLP
xb = r((M, 384)).bfloat16()
w1 = r((1536, 384), 0.05).bfloat16(); w2 = r((384, 1536), 0.03).bfloat16()
c1 = r((1536,), 0.1); d1 = r((1536,), 0.1); b2 = r((384,), 0.1)
stats = torch.zeros((6, M, 2), device="cuda")
stats[:, :, 1] = 64.0
probe("fused MLP (2 regions)", lambda: L.mlp_fused_bf16(xb, w1, c1, d1, w2, b2, stats))
a8k = r((8192, 8192)).bfloat16(); b8k = r((8192, 8192)).bfloat16()
probe("torch.matmul bf16 8192^3 (cuBLAS)", lambda: torch.matmul(a8k, b8k))
big = torch.empty(1 << 29, dtype=torch.bfloat16, device="cuda"); big2 = torch.empty_like(big)
probe("copy 1 GiB (HBM bound)", lambda: big2.copy_(big))
