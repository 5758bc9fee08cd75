"""How much of the step is launch gaps?  Eager forward_regions_u8 over 16 regions vs the same captured in one CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.common import seeded_modules
from hipt_abmil_atec23_b200.hipt_4k import HIPT_4K
DEV = torch.device("cuda:0")
m256, m4k = seeded_modules(0)
hipt = HIPT_4K.from_modules(m256, m4k, DEV, DEV)
regs = torch.randint(0, 256, (16, 3, 4096, 4096), dtype=torch.uint8, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
def timeit(f, n=6):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
eager = timeit(lambda: hipt.forward_regions_u8(regs))
ref = hipt.forward_regions_u8(regs).clone()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    hipt.forward_regions_u8(regs)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    out = hipt.forward_regions_u8(regs)
graphed = timeit(lambda: g.replay())
print(f"eager {eager:.3f} ms/step ({16e3/eager:.1f} regions/s)  graph {graphed:.3f} ms/step ({16e3/graphed:.1f} regions/s)  equal {torch.equal(out, ref)}")
