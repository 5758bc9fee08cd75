import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.common import seeded_modules
from hipt_abmil_atec23_b200.hipt_4k import HIPT_4K
DEV = torch.device("cuda:0")
m256, m4k = seeded_modules(0)
hipt = HIPT_4K.from_modules(m256, m4k, DEV, DEV)
gen = torch.Generator(device=DEV).manual_seed(1003)
regs = torch.randint(0, 256, (2, 3, 4096, 4096), dtype=torch.uint8, device=DEV, generator=gen)
e = hipt.model256._engine(DEV)
H = 256 * 257
for depth in (1, 2):
    e.set_depth_limit(depth)
    e.forward_patches(regs, mean=(0.5,)*3, std=(0.5,)*3, want_f32=False)
    torch.cuda.synchronize()
    pair = {k: e.buffer(k, 2 * H, c, torch.bfloat16).clone() for k, c in ((1, 384), (2, 1152), (3, 384))}
    hidp = e.buffer(4, 2 * 65536, 768, torch.bfloat16).clone()
    e.forward_patches(regs[1:2], mean=(0.5,)*3, std=(0.5,)*3, want_f32=False)
    torch.cuda.synchronize()
    alone = {k: e.buffer(k, H, c, torch.bfloat16).clone() for k, c in ((1, 384), (2, 1152), (3, 384))}
    hida = e.buffer(4, 65536, 768, torch.bfloat16).clone()
    print("depth", depth, "im2col", torch.equal(hidp[65536:], hida))
    for k, name in ((2, "qkv"), (3, "att"), (1, "xb")):
        a, b = pair[k][H:], alone[k]
        ne = (a != b).any(dim=1)
        print("  ", name, torch.equal(a, b), "rows differing", int(ne.sum()), "first", ne.nonzero()[:5].flatten().tolist(),
              "tokens", (ne.nonzero()[:8].flatten() % 257).tolist())
e.set_depth_limit(12)
