"""Two CUDA devices in ONE process (skipped on a single-GPU box): the reference's default HIPT_4K placement
(device256 = cuda:0, device4k = cuda:1, hipt_4k.py:36-46) and the nn.DataParallel wrap extract_features_fp.py:217-218 applies
whenever torch.cuda.device_count() > 1.  Regression for ADVICE r1: the dynamic-shared-memory opt-in is per device."""
import pytest
import torch
import torch.nn as nn

from oracle import hipt_oracle as O
from tests.common import seeded_modules

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two CUDA devices")]


def _region(seed=41, w=512, h=768):
    px = torch.randint(0, 256, (1, 3, w, h), dtype=torch.uint8, generator=torch.Generator().manual_seed(seed))
    return O.eval_transforms_u8(px)


def test_vit256_on_cuda0_and_vit4k_on_cuda1_in_one_process():
    from hipt_abmil_atec23_b200.hipt_4k import HIPT_4K
    x = _region()
    m256, m4k = seeded_modules(0)
    one = HIPT_4K.from_modules(m256, m4k, torch.device("cuda:0"), torch.device("cuda:0"))
    ref = one(x.to("cuda:0")).cpu()
    m256b, m4kb = seeded_modules(0)
    two = HIPT_4K.from_modules(m256b, m4kb, torch.device("cuda:0"), torch.device("cuda:1"))
    out = two(x.to("cuda:0"))
    assert out.device == torch.device("cuda:1")
    assert torch.equal(out.cpu(), ref)                                  # same kernels, same bits, whichever device runs them
    # and the other way round: every kernel's first launch on cuda:1 after it has already run on cuda:0
    m256c, m4kc = seeded_modules(0)
    swapped = HIPT_4K.from_modules(m256c, m4kc, torch.device("cuda:1"), torch.device("cuda:0"))
    assert torch.equal(swapped(x.to("cuda:1")).cpu(), ref)
    from hipt_abmil_atec23_b200.model_clam import CLAM_SB
    torch.manual_seed(2)
    clam = CLAM_SB(size_arg="hipt_big", n_classes=2).eval()
    bag = torch.randn(300, 192, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        a = clam.to("cuda:0")(bag.to("cuda:0"))[0].cpu()
        b = clam.to("cuda:1")(bag.to("cuda:1"))[0].cpu()
    assert torch.equal(a, b)


def test_data_parallel_wrap_of_the_extraction_script():
    """model = nn.DataParallel(model); model.eval(); features = model(batch) with batch [1,3,W,H] (batch_size 1 in the
    documented command line, docs/README.md:47)."""
    from hipt_abmil_atec23_b200.hipt_4k import HIPT_4K
    x = _region(42)
    m256, m4k = seeded_modules(0)
    model = HIPT_4K.from_modules(m256, m4k, torch.device("cuda:0"), torch.device("cuda:0"))
    ref = model(x.to("cuda:0")).cpu()
    dp = nn.DataParallel(model)
    dp.eval()
    with torch.no_grad():
        out = dp(x.to("cuda:0"))
    assert out.shape == (1, 192) and torch.equal(out.cpu(), ref)
