"""Accuracy of the CLAM score kernels against an fp64 evaluation of the oracle (HB_CLAM_TC=0 selects the CUDA-core path)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.common import seeded_clam
from oracle import hipt_oracle as O
from hipt_abmil_atec23_b200 import clam_engine
DEV = torch.device("cuda:0")
gen = torch.Generator().manual_seed(4)
lens = [300, 5000, 129, 128, 1, 20000]
offs = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
feats = torch.randn(sum(lens), 192, generator=gen) * 3.0
SIZE = sys.argv[1] if len(sys.argv) > 1 else "hipt_smaller"
FOLDS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
models = [seeded_clam(SIZE, 10 + i) for i in range(FOLDS)]
r = clam_engine.forward_bags([m.to(DEV) for m in models], feats.to(DEV), offs)
ea = el = 0.0
for mi, m in enumerate(models):
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    for b, n in enumerate(lens):
        bag = feats[offs[b]:offs[b + 1]]
        logits, y_prob, y_hat, a_raw, _ = O.clam_sb_forward({k: v.double() for k, v in sd.items()}, bag.double())
        ea = max(ea, (r["a_raw"][mi, offs[b]:offs[b + 1]].cpu().double() - a_raw[0]).abs().max().item())
        el = max(el, (r["logits"][mi, b].cpu().double() - logits[0]).abs().max().item())
print(f"{SIZE} x {FOLDS} folds, {'CUDA-core' if os.environ.get('HB_CLAM_TC') == '0' else 'tensor-core (split-TF32)'} path vs fp64 oracle: max |dA_raw| {ea:.3e}  max |dlogits| {el:.3e}")
