import sys, torch
sys.path.insert(0, "/root/repo")
from tests.common import seeded_modules
from hipt_abmil_atec23_b200.hipt_4k import HIPT_4K
DEV = torch.device("cuda:0")
m256, m4k = seeded_modules(0)
hipt = HIPT_4K.from_modules(m256, m4k, DEV, DEV)
gen = torch.Generator(device=DEV).manual_seed(1003)
regs = torch.randint(0, 256, (3, 3, 4096, 4096), dtype=torch.uint8, device=DEV, generator=gen)
o_t, c_t = hipt.forward_regions_u8(regs, return_cls256=True)
outs = [hipt.forward_regions_u8(regs[i:i + 1], return_cls256=True) for i in range(3)]
o_a = torch.cat([o[0] for o in outs]); c_a = torch.cat([o[1] for o in outs])
print("cls256 equal:", torch.equal(c_t, c_a), "max diff", (c_t.float() - c_a.float()).abs().max().item(), "rows differing", int((c_t != c_a).any(dim=1).sum()))
for i in range(3):
    print(" region", i, "cls equal", torch.equal(c_t[i*256:(i+1)*256], c_a[i*256:(i+1)*256]))
print("out equal:", torch.equal(o_t, o_a), (o_t - o_a).abs().max().item())
# ViT-4K alone on identical cls input, batched vs single
e4 = hipt.model4k._engine(DEV)
b = e4.forward_grid(c_t.contiguous(), 3, 16, 16)
s = torch.cat([e4.forward_grid(c_t[i*256:(i+1)*256].contiguous(), 1, 16, 16) for i in range(3)])
print("vit4k batched vs single equal:", torch.equal(b, s), (b - s).abs().max().item())
# depth-limited ViT-256: which block first differs
e = hipt.model256._engine(DEV)
for depth in (0, 1, 2, 11, 12):
    e.set_depth_limit(depth)
    _, ct = e.forward_patches(regs[:2], mean=(0.5,)*3, std=(0.5,)*3, want_f32=False)
    _, ca = e.forward_patches(regs[1:2], mean=(0.5,)*3, std=(0.5,)*3, want_f32=False)
    print("depth", depth, "region1 in pair vs alone equal:", torch.equal(ct[256:], ca))
e.set_depth_limit(12)
