"""Generate tests/golden/*.pt by running the REFERENCE's own modules (imported unmodified from /root/reference).

    python oracle/make_golden.py            # in the build container only; /root/reference does not travel

TEST INFRASTRUCTURE.  The reference has no tests or golden vectors for this path, so these files pin the oracle
(oracle/hipt_oracle.py) — and through it the CUDA path — to outputs of the reference implementation at fixed seeds
(SURVEY.md §8c/§8d).  Importable from the reference: HIPT_4K.vision_transformer, HIPT_4K.vision_transformer4k,
models.model_clam, models.model_mil, utils.utils.  NOT importable: HIPT_4K/hipt_4k.py and hipt_model_utils.py (TabError
at hipt_model_utils.py:72; h5py / matplotlib / skimage / webdataset missing), so the HIPT_4K.forward glue in the goldens
is the oracle's restatement of hipt_4k.py:63-76 driving the reference's ViT modules.
"""
import os
import sys
import time

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")


def _import_reference():
    sys.dont_write_bytecode = True
    for k in [k for k in sys.modules if k.split(".")[0] in ("HIPT_4K", "models", "utils")]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    import HIPT_4K.vision_transformer as rv
    import HIPT_4K.vision_transformer4k as rv4
    import models.model_clam as rclam
    import models.model_mil as rmil
    sys.path.remove(REF)
    return rv, rv4, rclam, rmil


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    rv, rv4, rclam, rmil = _import_reference()
    sys.path.insert(0, ROOT)
    from oracle import hipt_oracle as O

    # ---------------------------------------------------------------- models at the §8d seeds
    torch.manual_seed(0)
    m256 = rv.vit_small(patch_size=16, num_classes=0).eval()
    m4k = rv4.vit4k_xs(num_classes=0).eval()
    sd256 = {k: v.detach() for k, v in m256.state_dict().items()}
    sd4k = {k: v.detach() for k, v in m4k.state_dict().items()}
    gold = {"digest256": O.sd_digest(sd256), "digest4k": O.sd_digest(sd4k)}

    with torch.no_grad():
        # ------------------------------------------------------------ small ViT-256 case: 2 patches, per-block slices
        g = torch.Generator().manual_seed(11)
        px = torch.randint(0, 256, (2, 3, 256, 256), dtype=torch.uint8, generator=g)
        x = O.eval_transforms_u8(px)
        t = m256.prepare_tokens(x)
        per_block = [t[:, :3].clone()]
        for blk in m256.blocks:
            t = blk(t)
            per_block.append(t[:, :3].clone())
        gold["vit256_small"] = {"pixels_seed": 11, "tokens_first3_per_block": torch.stack(per_block),
                                "cls": m256(x).clone()}
        assert torch.equal(m256.norm(t)[:, 0], gold["vit256_small"]["cls"])

        # ------------------------------------------------------------ mini region 512 x 768 (2 x 3 grid), non-square 4K
        g = torch.Generator().manual_seed(12)
        reg = torch.randint(0, 256, (1, 3, 512 + 37, 768 + 11), dtype=torch.uint8, generator=g)   # exercises the crop
        xr = O.eval_transforms_u8(reg)
        img, w_256, h_256 = O.prepare_img_tensor(xr)
        batch = O.unfold_region(img)
        cls = m256(batch)
        grid = cls.reshape(w_256, h_256, 384).transpose(0, 1).transpose(0, 2).unsqueeze(0)
        out4k = m4k(grid)
        gold["mini_region"] = {"pixels_seed": 12, "shape": tuple(reg.shape), "cls256": cls.clone(),
                               "out": out4k.clone(), "w_256": w_256, "h_256": h_256}

        # ------------------------------------------------------------ config 1: full 4096 x 4096 region, seed 1
        t0 = time.time()
        reg = O.synthetic_region_u8(seed=1)
        xr = O.eval_transforms_u8(reg)
        img, w_256, h_256 = O.prepare_img_tensor(xr)
        batch = O.unfold_region(img)
        cls = torch.vstack([m256(batch[i:i + 64]) for i in range(0, 256, 64)])
        grid = cls.reshape(w_256, h_256, 384).transpose(0, 1).transpose(0, 2).unsqueeze(0)
        out4k = m4k(grid)
        print(f"reference full region: {time.time() - t0:.1f} s")
        gold["config1_region"] = {"pixels_seed": 1, "cls256": cls.clone(), "out": out4k.clone()}
        # the grid shuffle is the identity on token order
        assert torch.equal(grid.flatten(2, 3).transpose(1, 2)[0], cls)

        # ------------------------------------------------------------ CLAM_SB cases
        clam = {}
        for name, size_arg, seed, n, dropout, ncls in (("hipt_smaller_64", "hipt_smaller", 2, 64, 0.0, 2),
                                                       ("hipt_big_333", "hipt_big", 21, 333, 0.0, 2),
                                                       ("hipt_small_do_1000", "hipt_small", 22, 1000, 0.25, 2),
                                                       ("hipt_medium_5c_50", "hipt_medium", 23, 50, 0.0, 5),
                                                       ("hipt_smallest_1", "hipt_smallest", 24, 1, 0.0, 2)):
            torch.manual_seed(seed)
            mod = rclam.CLAM_SB(size_arg=size_arg, dropout=dropout, n_classes=ncls).eval()
            bag = torch.randn(n, 192, generator=torch.Generator().manual_seed(3 if n == 64 else seed + 100))
            logits, y_prob, y_hat, a_raw, res = mod(bag, return_features=True)
            clam[name] = {"size_arg": size_arg, "model_seed": seed, "dropout": dropout, "n_classes": ncls,
                          "bag_seed": 3 if n == 64 else seed + 100, "n": n, "digest": O.sd_digest(mod.state_dict()),
                          "logits": logits.clone(), "y_prob": y_prob.clone(), "y_hat": y_hat.clone(),
                          "a_raw": a_raw.clone(), "features": res["features"].clone(),
                          "attention_only": mod(bag, attention_only=True).clone()}
        gold["clam"] = clam

        # demo checkpoint (real trained CLAM_SB 'small', 1024-d): key cleaning as eval_utils.py:52-57, stress scores
        ck = torch.load(os.path.join(REF, "heatmaps/demo/ckpts/s_0_checkpoint.pt"), map_location="cpu")
        clean = {k.replace(".module", ""): v for k, v in ck.items() if "instance_loss_fn" not in k}
        mod = rclam.CLAM_SB(size_arg="small", dropout=True, n_classes=2)
        mod.load_state_dict(clean, strict=True)
        mod.eval()
        bag = torch.randn(300, 1024, generator=torch.Generator().manual_seed(5))
        logits, y_prob, y_hat, a_raw, _ = mod(bag)
        gold["clam_demo_ckpt"] = {"state_dict": {k: v.clone() for k, v in clean.items()}, "bag_seed": 5,
                                  "logits": logits.clone(), "y_prob": y_prob.clone(), "y_hat": y_hat.clone(),
                                  "a_raw": a_raw.clone()}

        # MIL_fc
        torch.manual_seed(31)
        mil = rmil.MIL_fc(n_classes=2).eval()
        bag = torch.randn(40, 1024, generator=torch.Generator().manual_seed(32))
        top, yp, yh, yps, _ = mil(bag)
        gold["mil_fc"] = {"model_seed": 31, "bag_seed": 32, "top_instance": top.clone(), "y_prob": yp.clone(),
                          "y_hat": yh.clone(), "y_probs": yps.clone()}

    # the demo checkpoint weights (2.6 MB) go to their own file
    torch.save(gold.pop("clam_demo_ckpt"), os.path.join(OUT, "clam_demo_ckpt.pt"))
    torch.save(gold, os.path.join(OUT, "hipt_reference_outputs.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))

    # ---------------------------------------------------------------- pin the oracle right away
    with torch.no_grad():
        o = O.hipt4k_forward(sd256, sd4k, O.eval_transforms_u8(
            torch.randint(0, 256, (1, 3, 549, 779), dtype=torch.uint8, generator=torch.Generator().manual_seed(12))))
        print("oracle vs reference, mini region max abs:", (o - gold["mini_region"]["out"]).abs().max().item())


if __name__ == "__main__":
    main()
