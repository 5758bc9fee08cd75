"""End-to-end parity of the CUDA path (through the reference-shaped Python API and the C ABI) against the golden
outputs of the reference (tests/golden/) and against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): region / patch features cosine >= 0.999 vs the reference fp32 path with max-abs
error reported; CLAM attention scores and slide logits within 1e-3 on identical input features; identical labels.
"""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import hipt_oracle as O
from tests.common import seeded_clam, seeded_modules
from hipt_abmil_atec23_b200 import _lib

pytestmark = pytest.mark.gpu
GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD_DIR, "hipt_reference_outputs.pt"), map_location="cpu")


@pytest.fixture(scope="module")
def hipt():
    from hipt_abmil_atec23_b200.hipt_4k import HIPT_4K
    m256, m4k = seeded_modules(0)
    return HIPT_4K.from_modules(m256, m4k, DEV, DEV)


def _cos(a, b):
    return F.cosine_similarity(a.double().flatten(), b.double().flatten(), dim=0).item()


def _min_row_cos(a, b):
    return F.cosine_similarity(a.double(), b.double(), dim=1).min().item()


def test_vit256_per_block_against_reference(gold, hipt):
    g = gold["vit256_small"]
    px = torch.randint(0, 256, (2, 3, 256, 256), dtype=torch.uint8,
                       generator=torch.Generator().manual_seed(g["pixels_seed"]))
    x = O.eval_transforms_u8(px).to(DEV)
    eng = hipt.model256._engine(DEV)
    try:
        for depth in (1, 2, 6, 12):
            eng.set_depth_limit(depth)
            eng.forward_patches(x)
            torch.cuda.synchronize()
            # the last block updates only the CLS rows (forward() returns x[:, 0]): compare 1 token at full depth
            nt = 1 if depth == 12 else 3
            if depth == 12:      # the CLS-only tail keeps its compact [n_seq, 384] stream in the head of the qkv buffer
                tok = eng.buffer(2, 2, 384, torch.bfloat16).view(2, 1, 384).float().cpu()
            else:
                tok = eng.buffer(1, 2 * 257, 384, torch.bfloat16).view(2, 257, 384)[:, :nt].float().cpu()
            ref = g["tokens_first3_per_block"][depth][:, :nt]
            err = (tok - ref).abs().max().item()
            print(f"depth {depth}: max abs {err:.4e}, cos {_cos(tok, ref):.6f}")
            assert _cos(tok, ref) > 0.9995, depth
    finally:
        eng.set_depth_limit(0)
    cls = hipt.model256(x).cpu()
    assert cls.shape == (2, 384) and cls.dtype == torch.float32
    assert _min_row_cos(cls, g["cls"]) >= 0.999
    print("vit256 cls max abs", (cls - g["cls"]).abs().max().item())


def test_tokens_before_blocks_match_oracle(hipt):
    """Patch embed (im2col + GEMM + bias + pos) and CLS rows against the oracle's prepare_tokens."""
    px = torch.randint(0, 256, (3, 3, 256, 256), dtype=torch.uint8, generator=torch.Generator().manual_seed(77))
    x = O.eval_transforms_u8(px)
    sd = {k: v.detach().cpu() for k, v in hipt.model256.state_dict().items()}
    ref = O.vit256_tokens(sd, x)
    eng = hipt.model256._engine(DEV)
    # depth limit cannot be 0 through the ABI (0 = all), so compare after zero blocks via the u8 path's token buffer:
    # run one block and check the token rows written by the embed kernels are consistent with block 1 of the oracle.
    eng.set_depth_limit(1)
    try:
        eng.forward_patches(x.to(DEV))
        torch.cuda.synchronize()
        tok = eng.buffer(1, 3 * 257, 384, torch.bfloat16).view(3, 257, 384).float().cpu()
    finally:
        eng.set_depth_limit(0)
    ref1 = O.block(sd, "blocks.0.", ref, 6)
    assert _cos(tok, ref1) > 0.9999
    assert (tok - ref1).abs().max().item() < 0.05


def test_fused_u8_patch_embed_tokens_match_oracle(hipt):
    """hb_vit256_forward_u8: unfold + ToTensor/Normalize + patch-embed conv + positional add from raw uint8 regions as one
    tensor-core kernel (hipt_4k.py:64-65, hipt_model_utils.py:113-118, vision_transformer.py:165-170, 240-244), checked one
    block downstream against the oracle on two regions of 2 x 3 patches each, every patch position and both tile halves."""
    from hipt_abmil_atec23_b200.hipt_model_utils import HIPT_MEAN, HIPT_STD
    reg = torch.randint(0, 256, (2, 3, 512, 768), dtype=torch.uint8, generator=torch.Generator().manual_seed(78))
    sd = {k: v.detach().cpu() for k, v in hipt.model256.state_dict().items()}
    patches = torch.cat([O.unfold_region(O.eval_transforms_u8(reg[i:i + 1])) for i in range(2)])      # [12, 3, 256, 256]
    ref1 = O.block(sd, "blocks.0.", O.vit256_tokens(sd, patches), 6)
    eng = hipt.model256._engine(DEV)
    eng.set_depth_limit(1)
    try:
        eng.forward_patches(reg.to(DEV), mean=HIPT_MEAN, std=HIPT_STD, want_f32=False)
        torch.cuda.synchronize()
        tok = eng.buffer(1, 12 * 257, 384, torch.bfloat16).view(12, 257, 384).float().cpu()
    finally:
        eng.set_depth_limit(0)
    assert _cos(tok, ref1) > 0.9999
    assert (tok - ref1).abs().max().item() < 0.05
    # and the unfused route (HB_EMBED_UNFUSED=1: im2col + GEMM with bf16 weights) lands on the same tokens
    os.environ["HB_EMBED_UNFUSED"] = "1"
    try:
        eng.set_depth_limit(1)
        eng.forward_patches(reg.to(DEV), mean=HIPT_MEAN, std=HIPT_STD, want_f32=False)
        torch.cuda.synchronize()
        tok2 = eng.buffer(1, 12 * 257, 384, torch.bfloat16).view(12, 257, 384).float().cpu()
    finally:
        os.environ.pop("HB_EMBED_UNFUSED", None)
        eng.set_depth_limit(0)
    assert _cos(tok, tok2) > 0.9999
    # a single [3, H, W] region and a patch sub-range of the batch go through the same kernel
    _, one = eng.forward_patches(reg[1].to(DEV), mean=HIPT_MEAN, std=HIPT_STD, want_f32=False)
    _, both = eng.forward_patches(reg.to(DEV), mean=HIPT_MEAN, std=HIPT_STD, want_f32=False)
    _, sub = eng.forward_patches(reg.to(DEV), patch_begin=4, n_patches=5, mean=HIPT_MEAN, std=HIPT_STD, want_f32=False)
    assert torch.equal(one, both[6:]) and torch.equal(sub, both[4:9])


def test_mini_region_forward_fp32_with_crop(gold, hipt):
    g = gold["mini_region"]
    reg = torch.randint(0, 256, g["shape"], dtype=torch.uint8, generator=torch.Generator().manual_seed(g["pixels_seed"]))
    x = O.eval_transforms_u8(reg).to(DEV)
    out = hipt(x)
    assert out.shape == (1, 192) and out.dtype == torch.float32 and out.device == DEV
    c = _cos(out.cpu(), g["out"])
    print("mini region cos", c, "max abs", (out.cpu() - g["out"]).abs().max().item())
    assert c >= 0.999
    # asset dict variant (hipt_4k.py:79-118)
    d = hipt.forward_asset_dict(x)
    assert d["features_cls256"].shape == (6, 384) and d["features_mean256_cls4k"].shape == (1, 576)
    assert _min_row_cos(torch.from_numpy(d["features_cls256"]), g["cls256"]) >= 0.999


def test_u8_path_equals_fp32_path(hipt):
    reg = torch.randint(0, 256, (2, 3, 512, 768), dtype=torch.uint8, generator=torch.Generator().manual_seed(5)).to(DEV)
    out_u8 = hipt.forward_regions_u8(reg)
    outs = [hipt(O.eval_transforms_u8(reg[i:i + 1].cpu()).to(DEV)) for i in range(2)]
    out_f = torch.cat(outs)
    assert out_u8.shape == (2, 192)
    assert _min_row_cos(out_u8, out_f) > 0.9995
    # the heatmap pipeline normalises with the ImageNet statistics (datasets/wsi_dataset.py:12-16): folded the same way
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    out_im = hipt.forward_regions_u8(reg, mean, std)
    m = torch.tensor(mean).view(1, 3, 1, 1)
    sdv = torch.tensor(std).view(1, 3, 1, 1)
    outs = [hipt(((reg[i:i + 1].cpu().float() / 255.0 - m) / sdv).to(DEV)) for i in range(2)]
    assert _min_row_cos(out_im, torch.cat(outs)) > 0.9995


def test_vit4k_many_regions_chunked_launches_match_oracle_and_single_calls(hipt):
    """ViT-4K over 150 region grids in one forward_grid call (the engine walks them in chunks of 64: two full chunks and a partial
    one) and forward_regions_u8 over 5 small regions: every region equals its own stand-alone call bit for bit, and a sample
    of regions on both sides of the chunk boundaries matches the fp32 oracle (cosine >= 0.999)."""
    eng = hipt.model4k._engine(DEV)
    R, T = 150, 256
    grid = (torch.randn(R * T, 384, generator=torch.Generator().manual_seed(21)) * 0.5).to(torch.bfloat16)
    out = eng.forward_grid(grid.to(DEV), R, 16, 16).clone()
    assert out.shape == (R, 192) and bool(torch.isfinite(out).all())
    sd4k = {k: v.detach().cpu() for k, v in hipt.model4k.state_dict().items()}
    for r in (0, 63, 64, 127, 128, 149):
        single = eng.forward_grid(grid[r * T:(r + 1) * T].contiguous().to(DEV), 1, 16, 16)
        assert torch.equal(single[0], out[r]), r
        g = grid[r * T:(r + 1) * T].float().view(1, 16, 16, 384).permute(0, 3, 1, 2).contiguous()
        with torch.no_grad():
            ref = O.vit4k_forward(sd4k, g)
        assert _cos(out[r].cpu(), ref[0]) >= 0.999, r
    reg = torch.randint(0, 256, (5, 3, 256, 512), dtype=torch.uint8, generator=torch.Generator().manual_seed(22)).to(DEV)
    all5 = hipt.forward_regions_u8(reg).clone()
    for i in range(5):
        assert torch.equal(hipt.forward_regions_u8(reg[i:i + 1])[0], all5[i]), i


def test_config1_full_region(gold, hipt):
    g = gold["config1_region"]
    reg = O.synthetic_region_u8(seed=g["pixels_seed"]).to(DEV)
    out, cls_bf16 = hipt.forward_regions_u8(reg, return_cls256=True)
    torch.cuda.synchronize()
    rc = _min_row_cos(cls_bf16.float().cpu(), g["cls256"])
    c = _cos(out.cpu(), g["out"])
    print(f"config1: min patch-CLS cosine {rc:.6f}, region cosine {c:.6f}, region max abs "
          f"{(out.cpu() - g['out']).abs().max().item():.4e}")
    assert rc >= 0.999 and c >= 0.999
    # the reference-API call on the normalised fp32 tensor gives the same region embedding
    out2 = hipt(O.eval_transforms_u8(reg.cpu()).to(DEV))
    assert _cos(out2, out) > 0.9995


def test_full_size_regions_are_grouping_invariant_and_deterministic(hipt):
    """Size-independent properties at BASELINE.json's full size (3 x 4096 x 4096 uint8): a region's embedding does not
    depend on which other regions share its launch (so a slide's bag is bit-identical for any rank count, SURVEY §8d config
    3), repeated runs are bit-identical, and the pinned-host pipeline returns the bits of the device-resident one."""
    from hipt_abmil_atec23_b200.pipeline import SlidePipeline
    gen = torch.Generator(device=DEV).manual_seed(1003)
    regs = torch.randint(0, 256, (3, 3, 4096, 4096), dtype=torch.uint8, device=DEV, generator=gen)
    together = hipt.forward_regions_u8(regs)                       # launches of two regions + one region
    again = hipt.forward_regions_u8(regs)
    alone = torch.cat([hipt.forward_regions_u8(regs[i:i + 1]) for i in range(3)])
    swapped = hipt.forward_regions_u8(regs.flip(0)).flip(0)
    assert torch.isfinite(together).all()
    assert torch.equal(together, again) and torch.equal(together, alone) and torch.equal(together, swapped)
    clam = [seeded_clam("hipt_smaller", 10 + f).to(DEV) for f in range(5)]
    pipe = SlidePipeline(hipt, clam)
    dev_out = pipe.run_device(regs)
    host_out = pipe.run_host(regs.cpu().pin_memory())
    for k in ("features", "logits", "a_raw", "y_hat"):
        assert torch.equal(dev_out[k].cpu(), host_out[k]), k
    # pooling does not depend on the order of the instances in the bag (up to fp32 summation order)
    perm = torch.tensor([2, 0, 1], device=DEV)
    r1 = pipe._pool(dev_out["features"])
    r2 = pipe._pool(dev_out["features"][perm])
    assert (r1["logits"] - r2["logits"]).abs().max().item() < 1e-5
    assert (r1["a_raw"][:, perm] - r2["a_raw"]).abs().max().item() < 1e-5


def test_last_selfattention_maps(hipt):
    """get_last_selfattention of both ViTs (used by the reference's hipt_heatmap_utils.get_region_attention_scores :328-335,
    which reads attention[:, :, 0, 1:]) against the oracle; bf16 residual stream -> probabilities within 2e-3."""
    px = torch.randint(0, 256, (3, 3, 256, 256), dtype=torch.uint8, generator=torch.Generator().manual_seed(31))
    x = O.eval_transforms_u8(px)
    sd = {k: v.detach().cpu() for k, v in hipt.model256.state_dict().items()}
    att = hipt.model256.get_last_selfattention(x.to(DEV)).cpu()
    ref = O.last_selfattention(sd, O.vit256_tokens(sd, x), 6)
    assert att.shape == (3, 6, 257, 257)
    assert (att.sum(-1) - 1).abs().max().item() < 1e-4
    assert (att - ref).abs().max().item() < 2e-3, (att - ref).abs().max().item()
    assert (att[:, :, 0, 1:] - ref[:, :, 0, 1:]).abs().max().item() < 2e-3
    # the regular forward is unaffected by the depth-limited pass
    out1 = hipt.model256(x.to(DEV))
    hipt.model256.get_last_selfattention(x.to(DEV))
    assert torch.equal(out1, hipt.model256(x.to(DEV)))
    sd4 = {k: v.detach().cpu() for k, v in hipt.model4k.state_dict().items()}
    grid = torch.randn(2, 384, 16, 16, generator=torch.Generator().manual_seed(32))
    att4 = hipt.model4k.get_last_selfattention(grid.to(DEV)).cpu()
    ref4 = O.last_selfattention(sd4, O.vit4k_tokens(sd4, grid), 6)
    assert att4.shape == (2, 6, 257, 257)
    assert (att4 - ref4).abs().max().item() < 2e-3, (att4 - ref4).abs().max().item()


def test_region_cls_attention_from_the_fused_kernel(hipt):
    """f4: HIPT_4K._get_region_attention_scores (hipt_4k.py:121-164) from the CLS-row probabilities the fused CLS-only
    attention launch writes during the ordinary forward pass — against the oracle's full attention maps (row 0, keys 1..),
    and against the same module's get_last_selfattention; the features of that pass are the ordinary forward's."""
    px = torch.randint(0, 256, (1, 3, 512, 768), dtype=torch.uint8, generator=torch.Generator().manual_seed(33))
    x = O.eval_transforms_u8(px)
    a256, a4k, w, h = hipt.region_cls_attention(x.to(DEV))
    assert (w, h) == (2, 3) and a256.shape == (6, 6, 256) and a4k.shape == (6, 6)
    sd = {k: v.detach().cpu() for k, v in hipt.model256.state_dict().items()}
    sd4 = {k: v.detach().cpu() for k, v in hipt.model4k.state_dict().items()}
    patches = O.unfold_region(x)
    ref256 = O.last_selfattention(sd, O.vit256_tokens(sd, patches), 6)[:, :, 0, 1:]
    assert (a256.cpu() - ref256).abs().max().item() < 2e-3, (a256.cpu() - ref256).abs().max().item()
    cls = O.vit256_forward(sd, patches)
    grid = cls.reshape(w, h, 384).transpose(0, 1).transpose(0, 2).unsqueeze(0)
    ref4k = O.last_selfattention(sd4, O.vit4k_tokens(sd4, grid), 6)[0, :, 0, 1:]
    assert (a4k.cpu() - ref4k).abs().max().item() < 3e-3, (a4k.cpu() - ref4k).abs().max().item()
    full = hipt.model256.get_last_selfattention(patches.to(DEV))[:, :, 0, 1:]
    assert (a256 - full).abs().max().item() < 2e-3
    # the reference-shaped entry point: numpy arrays with nearest-neighbour upsampling
    import numpy as np
    img = px[0].permute(1, 2, 0).numpy()
    b, att256, att4k = hipt._get_region_attention_scores(img, scale=4)
    assert b.shape == (6, 64, 64, 3) and b.dtype == np.uint8
    assert att256.shape == (6, 6, 64, 64) and att4k.shape == (6, 2 * 64, 3 * 64)
    assert np.allclose(att256[:, :, ::4, ::4], a256.reshape(6, 6, 16, 16).cpu().numpy(), atol=1e-6)
    assert np.allclose(att4k[:, ::64, ::64], a4k.reshape(6, 2, 3).cpu().numpy(), atol=1e-6)


def test_batch_gt_1_rejected_like_reference(hipt):
    with pytest.raises(RuntimeError):
        hipt(torch.zeros(2, 3, 256, 256, device=DEV))


def test_cpu_input_fails_loudly():
    m256, _ = seeded_modules(0)
    with pytest.raises(RuntimeError):
        m256(torch.zeros(1, 3, 256, 256))
    with pytest.raises(RuntimeError):
        seeded_clam()(torch.zeros(4, 192))


def test_loader_roundtrip(tmp_path, hipt):
    """get_vit256 / get_vit4k through a DINO-style checkpoint ('teacher' dict, module./backbone. prefixes, head.* extras)."""
    from hipt_abmil_atec23_b200.hipt_4k import HIPT_4K
    sd256 = {"module.backbone." + k: v.cpu() for k, v in hipt.model256.state_dict().items()}
    sd256["module.head.mlp.0.weight"] = torch.zeros(4, 4)
    sd4k = {"backbone." + k: v.cpu() for k, v in hipt.model4k.state_dict().items()}
    p256, p4k = str(tmp_path / "vit256.pth"), str(tmp_path / "vit4k.pth")
    torch.save({"teacher": sd256, "student": {}}, p256)
    torch.save({"teacher": sd4k}, p4k)
    model = HIPT_4K(p256, p4k, DEV, DEV).to(DEV).eval()
    assert not any(p.requires_grad for p in model.parameters())
    x = O.eval_transforms_u8(torch.randint(0, 256, (1, 3, 256, 512), dtype=torch.uint8,
                                           generator=torch.Generator().manual_seed(9))).to(DEV)
    assert torch.equal(model(x), hipt(x))
    with pytest.raises(AssertionError):
        HIPT_4K(str(tmp_path / "missing.pth"), p4k, DEV, DEV)


# ------------------------------------------------------------------------------------------------------- CLAM
def test_clam_cases_against_reference(gold):
    for name, g in gold["clam"].items():
        model = seeded_clam(g["size_arg"], g["model_seed"], g["dropout"], g["n_classes"]).to(DEV)
        bag = torch.randn(g["n"], 192, generator=torch.Generator().manual_seed(g["bag_seed"])).to(DEV)
        with torch.no_grad():
            logits, y_prob, y_hat, a_raw, res = model(bag, return_features=True)
            a_only = model(bag, attention_only=True)
        assert logits.shape == (1, g["n_classes"]) and a_raw.shape == (1, g["n"]) and y_hat.shape == (1, 1)
        assert y_hat.dtype == torch.int64
        for got, ref, what in ((logits, g["logits"], "logits"), (a_raw, g["a_raw"], "a_raw"), (a_only, g["a_raw"], "a_only"),
                               (y_prob, g["y_prob"], "y_prob"), (res["features"], g["features"], "M")):
            err = (got.cpu() - ref).abs().max().item()
            assert err < 1e-3, (name, what, err)
        assert torch.equal(y_hat.cpu(), g["y_hat"]), name


def test_clam_autograd_path_matches_fused(gold):
    g = gold["clam"]["hipt_smaller_64"]
    model = seeded_clam(g["size_arg"], g["model_seed"], 0.0, 2).to(DEV)
    bag = torch.randn(64, 192, generator=torch.Generator().manual_seed(3)).to(DEV)
    logits, _, _, a_raw, _ = model(bag)                    # grad enabled + trainable params -> torch composition
    assert logits.requires_grad
    F.cross_entropy(logits, torch.tensor([1], device=DEV)).backward()
    with torch.no_grad():
        l2, _, _, a2, _ = model(bag)
    assert (logits - l2).abs().max().item() < 1e-4 and (a_raw - a2).abs().max().item() < 1e-4


def _oracle_grads(sd_cpu, bag, label, extra=None):
    """Gradients of CE(logits, label) (+ extra(logits, y_prob, a_raw, M)) through the CPU oracle (plain torch autograd)."""
    sd = {k: v.clone().requires_grad_(True) for k, v in sd_cpu.items() if not k.startswith("instance_classifiers")}
    logits, y_prob, _, a_raw, res = O.clam_sb_forward(sd, bag, return_features=True)
    loss = F.cross_entropy(logits, torch.tensor([label]))
    if extra is not None:
        loss = loss + extra(logits, y_prob, a_raw, res["features"])
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in sd.items()}


@pytest.mark.parametrize("size_arg,n", [("hipt_smaller", 64), ("hipt_smaller", 333), ("hipt_smallest", 50), ("hipt_small", 200),
                                        ("hipt_medium", 130), ("hipt_big", 257)])
def test_clam_training_step_gradients_match_reference(size_arg, n):
    """a21: loss.backward() through CLAM_SB.forward runs the fused backward; gradients within 1e-3 (relative to the
    largest entry of each tensor) of autograd through the CPU oracle; tolerance from BASELINE.json's north_star."""
    model = seeded_clam(size_arg, 2).to(DEV).train()
    sd_cpu = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    bag = torch.randn(n, 192, generator=torch.Generator().manual_seed(3))
    logits, y_prob, y_hat, a_raw, res = model(bag.to(DEV), return_features=True)
    assert logits.requires_grad and logits.grad_fn is not None and "ClamSB" in type(logits.grad_fn).__name__
    extra = lambda lg, yp, ar, m: 0.3 * yp[0, 0] + 0.01 * (ar * ar).mean() + 0.1 * m.sum()
    loss = F.cross_entropy(logits, torch.tensor([1], device=DEV)) + extra(logits, y_prob, a_raw, res["features"])
    loss.backward()
    ref_loss, ref = _oracle_grads(sd_cpu, bag, 1, extra)
    assert abs(loss.item() - ref_loss.item()) < 1e-3
    for k, p in model.named_parameters():
        if k.startswith("instance_classifiers"):
            assert p.grad is None
            continue
        err = (p.grad.cpu() - ref[k]).abs().max().item()
        assert err < 1e-3 * max(1e-2, ref[k].abs().max().item()), (k, err, ref[k].abs().max().item())


def test_clam_five_classes_forward_and_gradients():
    """n_classes = 5 (the subtyping tasks of main.py:443-459) through the tensor-core score kernel, the fused backward and the
    lean training step, against the CPU oracle."""
    from hipt_abmil_atec23_b200.clam_engine import FusedAdam, TrainStep, _param_list
    model = seeded_clam("hipt_smaller", 7, 0.0, 5).to(DEV).train()
    sd_cpu = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    bag = torch.randn(300, 192, generator=torch.Generator().manual_seed(9))
    logits, y_prob, y_hat, a_raw, _ = model(bag.to(DEV))
    rl, rp, rh, ra, _ = O.clam_sb_forward(sd_cpu, bag)
    assert logits.shape == (1, 5) and (logits.detach().cpu() - rl).abs().max().item() < 1e-3
    assert (a_raw.detach().cpu() - ra).abs().max().item() < 1e-3 and int(y_hat) == int(rh)
    F.cross_entropy(logits, torch.tensor([3], device=DEV)).backward()
    ref_loss, ref = _oracle_grads(sd_cpu, bag, 3)
    for k, p in model.named_parameters():
        if k.startswith("instance_classifiers"):
            continue
        err = (p.grad.cpu() - ref[k]).abs().max().item()
        assert err < 1e-3 * max(1e-2, ref[k].abs().max().item()), (k, err)
    m2 = seeded_clam("hipt_smaller", 7, 0.0, 5).to(DEV).train()
    ts = TrainStep(m2, FusedAdam(_param_list(m2), lr=1e-3), 300)
    loss = ts.step(bag.to(DEV), torch.tensor([3], device=DEV))
    assert abs(loss.item() - ref_loss.item()) < 1e-4
    for (k, p) in m2.named_parameters():
        if not k.startswith("instance_classifiers"):
            assert (p.grad.cpu() - ref[k]).abs().max().item() < 1e-3 * max(1e-2, ref[k].abs().max().item()), k


def test_fused_adam_matches_torch_adam():
    from hipt_abmil_atec23_b200.clam_engine import FusedAdam
    torch.manual_seed(0)
    shapes = [(16, 192), (16,), (8, 16), (8,), (1, 8), (1,), (2, 16), (2,)]
    p_ref = [torch.randn(s) for s in shapes]
    p_our = [p.clone().to(DEV).requires_grad_(True) for p in p_ref]
    p_ref = [p.requires_grad_(True) for p in p_ref]
    ref = torch.optim.Adam(p_ref, lr=2e-3, weight_decay=1e-2)       # get_optim: Adam(lr=args.lr, weight_decay=args.reg)
    our = FusedAdam(p_our, lr=2e-3, weight_decay=1e-2)
    for step in range(4):
        for a, b in zip(p_ref, p_our):
            g = torch.randn(a.shape, generator=torch.Generator().manual_seed(100 * step + a.numel()))
            a.grad = g.clone()
            b.grad = g.to(DEV)
        ref.step()
        our.step()
    for a, b in zip(p_ref, p_our):
        assert (a.detach() - b.detach().cpu()).abs().max().item() < 1e-6


def test_clam_three_training_steps_track_the_reference_loop():
    """train_loop (utils/core_utils.py:409-423) for three bags: model(data) -> CE -> backward -> Adam step, fused kernels
    against the same loop on the CPU oracle's parameters with torch.optim.Adam."""
    from hipt_abmil_atec23_b200.clam_engine import FusedAdam
    model = seeded_clam("hipt_smaller", 2).to(DEV).train()
    ref_params = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()
                  if not k.startswith("instance_classifiers")}
    opt = FusedAdam(filter(lambda p: p.requires_grad, model.parameters()), lr=2e-4, weight_decay=1e-5)
    ref_opt = torch.optim.Adam(ref_params.values(), lr=2e-4, weight_decay=1e-5)
    n0 = _lib.launch_count()
    for step, (n, label) in enumerate([(70, 0), (129, 1), (64, 1)]):
        bag = torch.randn(n, 192, generator=torch.Generator().manual_seed(20 + step))
        logits, _, _, _, _ = model(bag.to(DEV))
        loss = F.cross_entropy(logits, torch.tensor([label], device=DEV))
        loss.backward()
        opt.step()
        opt.zero_grad()
        rl, _, _, _, _ = O.clam_sb_forward(ref_params, bag)
        rloss = F.cross_entropy(rl, torch.tensor([label]))
        rloss.backward()
        ref_opt.step()
        ref_opt.zero_grad()
        assert abs(loss.item() - rloss.item()) < 1e-4
    assert _lib.launch_count() - n0 >= 3 * 6                        # work table, scores, combine, prep, backward, adam per step
    for k, p in model.named_parameters():
        # attention_c.bias shifts every score equally: its gradient is identically zero (softmax is shift-invariant), what
        # is left is rounding noise of either sign, and Adam turns any non-zero value into a full +-lr step -- in the
        # reference as well.  Every other parameter must track.
        if k in ref_params and not k.endswith("attention_c.bias"):
            assert (p.detach().cpu() - ref_params[k].detach()).abs().max().item() < 2e-5, k


def test_clam_lean_train_step_matches_the_autograd_route():
    """clam_engine.TrainStep (forward, backward with the cross-entropy fused in, one-launch Adam; no autograd) against the
    CLAM_SB.forward + F.cross_entropy + loss.backward() + FusedAdam route on the same bags: losses and parameters agree."""
    from hipt_abmil_atec23_b200.clam_engine import FusedAdam, TrainStep
    ma = seeded_clam("hipt_smaller", 2).to(DEV).train()
    mb = seeded_clam("hipt_smaller", 2).to(DEV).train()
    oa = FusedAdam(filter(lambda p: p.requires_grad, ma.parameters()), lr=2e-4, weight_decay=1e-5)
    from hipt_abmil_atec23_b200.clam_engine import _param_list
    ob = FusedAdam(_param_list(mb), lr=2e-4, weight_decay=1e-5)
    ts = TrainStep(mb, ob, max_instances=400)
    for step, (n, label) in enumerate([(70, 0), (333, 1), (128, 1), (1, 0)]):
        bag = torch.randn(n, 192, generator=torch.Generator().manual_seed(40 + step)).to(DEV)
        y = torch.tensor([label], device=DEV)
        logits = ma(bag)[0]
        la = F.cross_entropy(logits, y)
        la.backward()
        oa.step()
        oa.zero_grad()
        lb = ts.step(bag, y)
        assert abs(la.item() - lb.item()) < 1e-5
    for (k, pa), pb in zip(ma.named_parameters(), mb.parameters()):
        if k.startswith("instance_classifiers") or k.endswith("attention_c.bias"):
            continue
        assert (pa - pb).abs().max().item() < 1e-6, k


@pytest.mark.parametrize("size_arg,folds", [("hipt_smaller", 5), ("hipt_smaller", 1), ("hipt_smaller", 3), ("hipt_small", 2),
                                            ("hipt_medium", 1)])
def test_clam_ragged_bags_and_fold_ensemble(size_arg, folds):
    """Every head / fold count the tensor-core score kernel is instantiated for (first Linear and both gate Linears as split-TF32
    tcgen05 GEMMs), ragged bags from 0 to 20,000 instances, against the oracle."""
    from hipt_abmil_atec23_b200 import clam_engine
    gen = torch.Generator().manual_seed(4)
    lens = [50, 75, 200, 1000, 5000, 20000, 1, 129, 128, 0, 333]
    offs = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    feats = torch.randn(sum(lens), 192, generator=gen)
    models = [seeded_clam(size_arg, 10 + i) for i in range(folds)]
    r = clam_engine.forward_bags([m.to(DEV) for m in models], feats.to(DEV), offs)
    torch.cuda.synchronize()
    for mi, m in enumerate(models):
        sd = {k: v.cpu() for k, v in m.state_dict().items()}
        for b, n in enumerate(lens):
            if n == 0:
                assert torch.allclose(r["logits"][mi, b].cpu(), sd["classifiers.bias"], atol=1e-6)
                continue
            bag = feats[offs[b]:offs[b + 1]]
            logits, y_prob, y_hat, a_raw, _ = O.clam_sb_forward(sd, bag)
            # bar (BASELINE.json north_star): 1e-3.  This call runs the split-TF32 tensor-core kernel (5 folds x L1 16): the
            # hi/lo operand split keeps it at fp32 level, which the tighter bound pins.
            assert (r["a_raw"][mi, offs[b]:offs[b + 1]].cpu() - a_raw[0]).abs().max().item() < 5e-5
            assert (r["logits"][mi, b].cpu() - logits[0]).abs().max().item() < 5e-5
            assert (r["y_prob"][mi, b].cpu() - y_prob[0]).abs().max().item() < 1e-3
            assert int(r["y_hat"][mi, b]) == int(y_hat)


@pytest.mark.parametrize("size_arg,folds", [("hipt_smaller", 5), ("hipt_small", 2), ("hipt_medium", 1)])
def test_clam_tensor_core_forward_is_bit_identical_run_to_run(size_arg, folds):
    """The tensor-core score kernel hands tiles between five roles through mbarriers and relies on the tensor pipe executing in
    issue order (the gate of tile t reads its A operand from the accumulator that GEMM 1 of tile t + 2 overwrites): a protocol
    error shows up as a run-to-run difference long before it breaks a tolerance.  60 launches over ragged bags, all outputs
    compared bit for bit with the first."""
    from hipt_abmil_atec23_b200 import clam_engine
    lens = [5000, 129, 20000, 777, 128, 3000]
    offs = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=DEV)
    feats = torch.randn(sum(lens), 192, generator=torch.Generator().manual_seed(11)).to(DEV)
    models = [seeded_clam(size_arg, 20 + i).to(DEV) for i in range(folds)]
    first = clam_engine.forward_bags(models, feats, offs, max_bag_len=max(lens))
    torch.cuda.synchronize()
    for it in range(60):
        r = clam_engine.forward_bags(models, feats, offs, max_bag_len=max(lens))
        for k in ("a_raw", "m", "logits", "y_prob", "y_hat"):
            assert torch.equal(r[k], first[k]), (size_arg, folds, it, k)


def test_clam_forward_replays_from_a_cuda_graph():
    """The tensor-core path launches its score and combine kernels as programmatic dependents (cudaLaunchKernelEx); captured
    into a CUDA graph on a side stream and replayed on new features, it must give what the eager call gives."""
    from hipt_abmil_atec23_b200 import clam_engine
    lens = [700, 129, 4000, 128]
    offs = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=DEV)
    gen = torch.Generator().manual_seed(31)
    feats = torch.randn(sum(lens), 192, generator=gen).to(DEV)
    models = [seeded_clam("hipt_smaller", 40 + i).to(DEV) for i in range(2)]
    clam_engine.forward_bags(models, feats, offs, max_bag_len=max(lens))          # warm-up outside the capture
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        captured = clam_engine.forward_bags(models, feats, offs, max_bag_len=max(lens))
    feats.copy_(torch.randn(sum(lens), 192, generator=gen).to(DEV))               # new inputs, same buffers
    graph.replay()
    torch.cuda.synchronize()
    eager = clam_engine.forward_bags(models, feats, offs, max_bag_len=max(lens))
    for k in ("a_raw", "m", "logits", "y_prob", "y_hat"):
        assert torch.equal(captured[k], eager[k]), k


def test_clam_forward_writes_stay_inside_their_buffers():
    """Guard words around every output and the workspace of hb_clam_sb_forward (tensor-core and CUDA-core score kernels, ragged
    bags incl. a partial last chunk and an empty bag): nothing outside the documented extents is written."""
    import ctypes as C
    from hipt_abmil_atec23_b200 import clam_engine
    lib = _lib.load()
    lens = [300, 0, 129, 1000, 57]
    total = sum(lens)
    offs = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=DEV)
    feats = torch.randn(total, 192, generator=torch.Generator().manual_seed(8)).to(DEV)
    GUARD, SENT = 64, 12345.0
    # tensor-core path: hipt_smaller x 1 / 3 / 5, hipt_small x 2, hipt_medium x 1; CUDA-core path: 8 folds, hipt_big
    for size_arg, n_models in (("hipt_smaller", 1), ("hipt_smaller", 3), ("hipt_smaller", 5), ("hipt_small", 2), ("hipt_medium", 1),
                               ("hipt_smaller", 8), ("hipt_big", 2)):
        models = [seeded_clam(size_arg, 10 + i).to(DEV) for i in range(n_models)]
        L1 = models[0].attention_net[0].out_features
        D = clam_engine._gate_module(models[0]).attention_c.in_features
        keep = [clam_engine._weights(m, DEV) for m in models]
        arr = (C.c_void_p * (10 * n_models))(*[t.data_ptr() for w in keep for t in w])
        def guarded(n, dtype=torch.float32):
            buf = torch.full((n + 2 * GUARD,), SENT, dtype=dtype, device=DEV)
            return buf, buf[GUARD:GUARD + n]
        a_buf, a_raw = guarded(n_models * total)
        m_buf, m_out = guarded(n_models * len(lens) * L1)
        l_buf, logits = guarded(n_models * len(lens) * 2)
        p_buf, y_prob = guarded(n_models * len(lens) * 2)
        ws_bytes = lib.hb_clam_workspace_bytes(max(lens), len(lens), n_models, L1)
        w_buf, ws = guarded(ws_bytes // 4 + 4)
        y_hat = torch.empty(n_models * len(lens), dtype=torch.int64, device=DEV)
        _lib.check(lib.hb_clam_sb_forward(_lib.ptr(feats), _lib.ptr(offs), len(lens), total, max(lens), arr, n_models, 192, L1, D, 2,
                                          _lib.ptr(a_raw), _lib.ptr(m_out), _lib.ptr(logits), _lib.ptr(y_prob), _lib.ptr(y_hat),
                                          _lib.ptr(ws), ws.numel() * 4, _lib.stream_ptr()))
        torch.cuda.synchronize()
        for name, buf in (("a_raw", a_buf), ("m", m_buf), ("logits", l_buf), ("y_prob", p_buf), ("workspace", w_buf)):
            assert bool((buf[:GUARD] == SENT).all()) and bool((buf[-GUARD:] == SENT).all()), (size_arg, n_models, name)
        assert bool(torch.isfinite(a_raw).all()) and bool((a_raw != SENT).all())


def test_clam_demo_checkpoint_trained_weights():
    from hipt_abmil_atec23_b200.model_clam import CLAM_SB
    g = torch.load(os.path.join(GOLD_DIR, "clam_demo_ckpt.pt"), map_location="cpu")
    model = CLAM_SB(size_arg="small", dropout=True, n_classes=2)
    model.load_state_dict(g["state_dict"], strict=True)          # gate at attention_net.3 with dropout=True
    model = model.to(DEV).eval()
    bag = torch.randn(300, 1024, generator=torch.Generator().manual_seed(g["bag_seed"])).to(DEV)
    with torch.no_grad():
        logits, y_prob, y_hat, a_raw, _ = model(bag)
    assert (a_raw.cpu() - g["a_raw"]).abs().max().item() < 1e-3 * max(1.0, g["a_raw"].abs().max().item())
    assert (logits.cpu() - g["logits"]).abs().max().item() < 1e-3 * max(1.0, g["logits"].abs().max().item())
    assert torch.equal(y_hat.cpu(), g["y_hat"])
