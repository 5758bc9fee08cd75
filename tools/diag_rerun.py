"""Which kernel is not run-to-run deterministic inside the model: run ViT-256 over two full-size regions several times per
depth limit and compare the plan's qkv / att / xb buffers bit for bit with the first run."""
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
H = 2 * 256 * 257
for depth in (1, 2, 3, 6, 11):
    e.set_depth_limit(depth)
    first = None
    for rep in range(6):
        e.forward_patches(regs, mean=(0.5,)*3, std=(0.5,)*3, want_f32=False)
        torch.cuda.synchronize()
        cur = {k: e.buffer(k, H, c, torch.bfloat16).clone() for k, c in ((1, 384), (2, 1152), (3, 384))}
        if first is None:
            first = cur
            continue
        for k, name in ((2, "qkv"), (3, "att"), (1, "xb")):
            a, b = first[k], cur[k]
            if not torch.equal(a, b):
                ne = (a != b)
                rows = ne.any(dim=1).nonzero().flatten()
                cols = ne.any(dim=0).nonzero().flatten()
                print(f"depth {depth} rep {rep} {name}: {len(rows)} rows differ, tokens {sorted(set((rows % 257).tolist()))[:12]}, "
                      f"seqs {sorted(set((rows // 257).tolist()))[:12]}, cols {cols[:4].tolist()}..{cols[-4:].tolist()} ({len(cols)}), "
                      f"max |d| {(a.float() - b.float()).abs().max().item():.3e}")
    print("depth", depth, "done")
e.set_depth_limit(12)
