import sys, torch
sys.path.insert(0, "/root/repo")
from tests.common import seeded_modules
from hipt_abmil_atec23_b200.hipt_4k import HIPT_4K
from hipt_abmil_atec23_b200.hipt_model_utils import HIPT_MEAN, HIPT_STD
DEV = torch.device("cuda:0")
m256, m4k = seeded_modules(0)
hipt = HIPT_4K.from_modules(m256, m4k, DEV, DEV)
regs = torch.randint(0, 256, (2, 3, 4096, 4096), dtype=torch.uint8, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
e = hipt.model256._engine(DEV)
e.set_depth_limit(1)
for _ in range(3):
    e.forward_patches(regs, mean=HIPT_MEAN, std=HIPT_STD, want_f32=False)
torch.cuda.synchronize()
print("ok")
