"""Seeded model construction shared by the tests (SURVEY.md §8d seeds)."""
import torch


def seeded_modules(seed=0):
    from hipt_abmil_atec23_b200 import vision_transformer as vits
    from hipt_abmil_atec23_b200 import vision_transformer4k as vits4k
    torch.manual_seed(seed)
    m256 = vits.vit_small(patch_size=16, num_classes=0).eval()
    m4k = vits4k.vit4k_xs(num_classes=0).eval()
    return m256, m4k


def seeded_vits(seed=0):
    m256, m4k = seeded_modules(seed)
    return ({k: v.detach() for k, v in m256.state_dict().items()},
            {k: v.detach() for k, v in m4k.state_dict().items()})


def seeded_clam(size_arg="hipt_smaller", seed=2, dropout=0.0, n_classes=2):
    from hipt_abmil_atec23_b200.model_clam import CLAM_SB
    torch.manual_seed(seed)
    return CLAM_SB(size_arg=size_arg, dropout=dropout, n_classes=n_classes).eval()
