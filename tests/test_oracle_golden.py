"""The CPU oracle against golden outputs of the reference's own modules (tests/golden/, made by oracle/make_golden.py).

These run without a GPU.  They pin (1) that seeded construction through THIS repo's module classes reproduces the
reference's random-init weights bit for bit, and (2) that every oracle function agrees with the reference outputs.
"""
import os

import pytest
import torch

from oracle import hipt_oracle as O

GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD_DIR, "hipt_reference_outputs.pt"), map_location="cpu")


@pytest.fixture(scope="module")
def vits():
    from tests.common import seeded_vits
    return seeded_vits()


def test_seeded_weights_match_reference(gold, vits):
    sd256, sd4k = vits
    d256, d4k = O.sd_digest(sd256), O.sd_digest(sd4k)
    assert d256["n"] == 150 and d4k["n"] == 78
    assert d256["sha1_small"] == gold["digest256"]["sha1_small"] and d256["sum"] == gold["digest256"]["sum"]
    assert d4k["sha1_small"] == gold["digest4k"]["sha1_small"] and d4k["sum"] == gold["digest4k"]["sum"]


def test_vit256_blocks(gold, vits):
    sd256, _ = vits
    g = gold["vit256_small"]
    px = torch.randint(0, 256, (2, 3, 256, 256), dtype=torch.uint8,
                       generator=torch.Generator().manual_seed(g["pixels_seed"]))
    x = O.eval_transforms_u8(px)
    with torch.no_grad():
        for depth in (0, 1, 6, 12):
            t = O.vit256_forward(sd256, x, return_tokens=True, depth_limit=depth)
            ref = g["tokens_first3_per_block"][depth]
            assert (t[:, :3] - ref).abs().max().item() < 2e-4 * max(1.0, ref.abs().max().item()), depth
        cls = O.vit256_forward(sd256, x)
    assert (cls - g["cls"]).abs().max().item() < 1e-4


def test_mini_region_with_crop_and_nonsquare_grid(gold, vits):
    sd256, sd4k = vits
    g = gold["mini_region"]
    reg = torch.randint(0, 256, g["shape"], dtype=torch.uint8, generator=torch.Generator().manual_seed(g["pixels_seed"]))
    with torch.no_grad():
        out, cls = O.hipt4k_forward(sd256, sd4k, O.eval_transforms_u8(reg), return_cls256=True)
    assert (g["w_256"], g["h_256"]) == (2, 3)
    assert (cls - g["cls256"]).abs().max().item() < 1e-4
    assert (out - g["out"]).abs().max().item() < 1e-4


def test_config1_vit4k_on_golden_cls(gold, vits):
    """ViT-4K stage of config 1 on the reference's own [256,384] CLS matrix (the ViT-256 stage of the full region is
    exercised on the GPU box, where the oracle's 11 s forward is the bench's cpu_baseline)."""
    _, sd4k = vits
    g = gold["config1_region"]
    grid = g["cls256"].reshape(16, 16, 384).transpose(0, 1).transpose(0, 2).unsqueeze(0)
    with torch.no_grad():
        out = O.vit4k_forward(sd4k, grid)
    assert (out - g["out"]).abs().max().item() < 1e-4
    assert torch.equal(grid.flatten(2, 3).transpose(1, 2)[0], g["cls256"])      # grid shuffle = identity on tokens


def test_clam_cases(gold):
    from tests.common import seeded_clam
    for name, g in gold["clam"].items():
        model = seeded_clam(g["size_arg"], g["model_seed"], g["dropout"], g["n_classes"])
        sd = model.state_dict()
        assert O.sd_digest(sd)["sha1_small"] == g["digest"]["sha1_small"], name
        bag = torch.randn(g["n"], 192, generator=torch.Generator().manual_seed(g["bag_seed"]))
        with torch.no_grad():
            logits, y_prob, y_hat, a_raw, res = O.clam_sb_forward(sd, bag, return_features=True)
            a_only = O.clam_sb_forward(sd, bag, attention_only=True)
        assert torch.allclose(logits, g["logits"], atol=1e-5), name
        assert torch.allclose(y_prob, g["y_prob"], atol=1e-6), name
        assert torch.equal(y_hat, g["y_hat"]), name
        assert torch.allclose(a_raw, g["a_raw"], atol=1e-5), name
        assert torch.allclose(res["features"], g["features"], atol=1e-5), name
        assert torch.allclose(a_only, g["attention_only"], atol=1e-5), name
        assert a_raw.shape == (1, g["n"]) and y_hat.dtype == torch.int64


def test_clam_demo_checkpoint():
    g = torch.load(os.path.join(GOLD_DIR, "clam_demo_ckpt.pt"), map_location="cpu")
    bag = torch.randn(300, 1024, generator=torch.Generator().manual_seed(g["bag_seed"]))
    with torch.no_grad():
        logits, y_prob, y_hat, a_raw, _ = O.clam_sb_forward(g["state_dict"], bag)
    assert torch.allclose(logits, g["logits"], atol=1e-4)
    assert torch.equal(y_hat, g["y_hat"])
    assert torch.allclose(a_raw, g["a_raw"], atol=1e-4)
    assert a_raw.abs().max().item() > 10          # trained weights: a real stress input for the softmax over N


def test_mil_fc(gold):
    from hipt_abmil_atec23_b200.model_mil import MIL_fc
    g = gold["mil_fc"]
    torch.manual_seed(g["model_seed"])
    model = MIL_fc(n_classes=2).eval()
    bag = torch.randn(40, 1024, generator=torch.Generator().manual_seed(g["bag_seed"]))
    with torch.no_grad():
        top, yp, yh, yps, _ = O.mil_fc_forward(model.state_dict(), bag)
        top2, yp2, yh2, yps2, _ = model(bag)         # the product module (torch composition, API parity only)
    for a, b in ((top, g["top_instance"]), (yp, g["y_prob"]), (yps, g["y_probs"]), (top2, g["top_instance"]),
                 (yps2, g["y_probs"])):
        assert torch.allclose(a, b, atol=1e-5)
    assert torch.equal(yh, g["y_hat"]) and torch.equal(yh2, g["y_hat"])


# ------------------------------------------------------------------------------------------ CLAM_SB training-mode goldens
def _train_gold():
    return torch.load(os.path.join(GOLD_DIR, "clam_train_reference.pt"), map_location="cpu")


def test_clam_instance_eval_oracle_matches_reference():
    from tests.common import seeded_clam
    for name, g in _train_gold()["inst_eval"].items():
        torch.manual_seed(g["model_seed"])
        from hipt_abmil_atec23_b200.model_clam import CLAM_SB
        sd = CLAM_SB(size_arg=g["size_arg"], dropout=0.0, n_classes=g["n_classes"], subtyping=g["subtyping"]).state_dict()
        bag = torch.randn(g["n"], 192, generator=torch.Generator().manual_seed(g["bag_seed"]))
        with torch.no_grad():
            logits, _, _, a_raw, res = O.clam_sb_forward_train(sd, bag, None, torch.tensor([g["label"]]), True, 8, g["subtyping"],
                                                               g["n_classes"])
        assert torch.allclose(logits, g["logits"], atol=1e-6) and torch.allclose(a_raw, g["a_raw"], atol=1e-6), name
        assert torch.allclose(torch.as_tensor(res["instance_loss"]), g["instance_loss"], atol=1e-6), name
        assert torch.equal(res["inst_preds"], g["inst_preds"]) and torch.equal(res["inst_labels"], g["inst_labels"]), name


def test_clam_training_oracle_matches_reference_gradients_with_known_dropout_masks():
    """Oracle (mask application points restated from model_clam.py:50-52, 84-85) vs the reference module run with the same
    masks injected through forward hooks: outputs, loss and every gradient."""
    import torch.nn.functional as F
    from hipt_abmil_atec23_b200 import clam_engine
    from hipt_abmil_atec23_b200.model_clam import CLAM_SB
    for name, g in _train_gold()["train"].items():
        torch.manual_seed(g["model_seed"])
        mod = CLAM_SB(size_arg=g["size_arg"], dropout=g["dropout"], n_classes=g["n_classes"], subtyping=g["subtyping"])
        sd = {k: v.detach().clone().requires_grad_(True) for k, v in mod.state_dict().items()}
        bag = torch.randn(g["n"], 192, generator=torch.Generator().manual_seed(g["bag_seed"]))
        L1, D = sd["attention_net.0.weight"].shape[0], sd["classifiers.weight"].shape[1] // 2
        masks = clam_engine.dropout_masks(g["n"], L1, D, g["dropout"], g["mask_seed"]) if g["dropout"] > 0 else None
        lab = torch.tensor([g["label"]])
        logits, _, _, a_raw, res = O.clam_sb_forward_train(sd, bag, masks, lab, g["instance_eval"], 8, g["subtyping"], g["n_classes"])
        loss = F.cross_entropy(logits, lab)
        total = 0.7 * loss + 0.3 * res["instance_loss"] if g["instance_eval"] else loss
        total.backward()
        assert torch.allclose(logits, g["logits"], atol=1e-5) and torch.allclose(a_raw, g["a_raw"], atol=1e-5), name
        assert torch.allclose(total.detach(), g["loss"], atol=1e-5), name
        for k, ref in g["grads"].items():
            got = sd[k].grad
            if ref is None:
                assert got is None or float(got.abs().max()) == 0.0, (name, k)
            else:
                assert torch.allclose(got, ref, atol=1e-5, rtol=1e-4), (name, k, float((got - ref).abs().max()))
