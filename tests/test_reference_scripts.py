"""The reference's own scripts, run UNCHANGED against this repository's import shims ("extract_features_fp.py,
main.py, eval.py and create_heatmaps.py run unchanged").

The scripts come from the reference checkout when it is present (/root/reference, build container) and otherwise from
oracle/_ref, where oracle/build_ref.py byte-compiled them (GPU box).  The repository root is placed BEFORE the reference on
sys.path, exactly as INTEGRATION.md tells a maintainer to do; modules the image lacks (h5py, openslide, timm, torchstain,
ray, matplotlib, cv2) are test-only stand-ins.  What runs is the reference's code: compute_w_loader
(extract_features_fp.py:26-173) with its Whole_Slide_Bag_FP dataset, DataLoader worker, collate_features and save_hdf5;
eval_utils.initiate_model + summary (utils/eval_utils.py:25-60, 115-179); create_heatmaps.infer_single_slide (:34-57);
Generic_MIL_Dataset's .pt read (datasets/dataset_generic.py:505-528) on files written by the feature sink.
"""
import argparse
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HB_TEST_REF") or ("/root/reference" if os.path.isdir("/root/reference") else os.path.join(ROOT, "oracle", "_ref"))
HAVE_REF = os.path.exists(os.path.join(REF, "extract_features_fp.py")) or os.path.exists(os.path.join(REF, "extract_features_fp.pyc"))
STUBS = os.path.join(ROOT, "tests", "stubs")
SHIMS = ("HIPT_4K", "models", "utils", "datasets")
REF_TOP = SHIMS + ("extract_features_fp", "create_heatmaps", "wsi_core", "vis_utils")
FAKE = ("openslide", "timm", "torchstain", "ray", "matplotlib", "cv2", "seaborn")

pytestmark = pytest.mark.skipif(not HAVE_REF, reason="neither /root/reference nor oracle/_ref (python oracle/build_ref.py) is present")


class _Anything(types.ModuleType):
    """A module whose every attribute exists: enough for `import matplotlib.pyplot as plt` / `from ray import tune` at import
    time of reference modules whose plotting / tuning code these tests never reach."""
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return self


class _FakeFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in FAKE:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Anything(spec.name)

    def exec_module(self, module):
        pass


class _SourcelessScripts(importlib.abc.MetaPathFinder):
    """Top-level scripts byte-compiled into oracle/_ref (extract_features_fp.pyc, create_heatmaps.pyc) import like modules."""

    def find_spec(self, fullname, path=None, target=None):
        f = os.path.join(REF, fullname + ".pyc")
        if "." not in fullname and os.path.isfile(f) and not os.path.isfile(os.path.join(REF, fullname + ".py")):
            return importlib.util.spec_from_file_location(fullname, f, loader=importlib.machinery.SourcelessFileLoader(fullname, f))
        return None


@pytest.fixture(scope="module")
def ref_env():
    """sys.path = [stubs, repo root, reference]; the shim packages (possibly imported earlier without the reference in
    sight) re-extend their __path__; everything is undone afterwards."""
    import importlib.util  # noqa: F401
    from hipt_abmil_atec23_b200 import shim
    saved_path, saved_argv = list(sys.path), list(sys.argv)
    finders = [_FakeFinder(), _SourcelessScripts()]
    sys.meta_path[:0] = finders
    sys.path[:] = [STUBS, ROOT] + [p for p in sys.path if p not in (STUBS, ROOT, REF)] + [REF]
    for k in [k for k in sys.modules if k.split(".")[0] in ("extract_features_fp", "create_heatmaps", "wsi_core", "vis_utils", "h5py")]:
        del sys.modules[k]
    for name in SHIMS:
        pkg = importlib.import_module(name)
        shim.extend_package_path(name, pkg.__path__)
    if "utils.utils" in sys.modules:
        importlib.reload(sys.modules["utils.utils"])
    yield
    for f in finders:
        sys.meta_path.remove(f)
    sys.path[:] = saved_path
    sys.argv[:] = saved_argv
    for k in list(sys.modules):
        top = k.split(".")[0]
        mod_file = getattr(sys.modules[k], "__file__", None) or ""
        if top in FAKE or top == "h5py" or (top in REF_TOP and mod_file.startswith(REF)):
            del sys.modules[k]
    for name in SHIMS:
        pkg = sys.modules.get(name)
        if pkg is not None:
            pkg.__path__[:] = [p for p in pkg.__path__ if not os.path.realpath(p).startswith(os.path.realpath(REF))]
    if "utils.utils" in sys.modules:
        importlib.reload(sys.modules["utils.utils"])


def test_reference_modules_resolve_next_to_the_shims(ref_env):
    """Import resolution only (no GPU): the accelerated classes come from this repository, everything else from the
    reference — including `datasets`, which must not be the HuggingFace distribution."""
    import datasets.dataset_h5 as dh5
    import models.model_clam as mc
    import models.resnet_custom as rc
    import utils.file_utils as fu
    import utils.utils as uu
    import HIPT_4K.hipt_4k as h4k
    assert mc.CLAM_SB.__module__ == "hipt_abmil_atec23_b200.model_clam"
    assert h4k.HIPT_4K.__module__ == "hipt_abmil_atec23_b200.hipt_4k"
    for m in (dh5, rc, fu):
        assert os.path.realpath(m.__file__).startswith(os.path.realpath(REF)), m.__file__
    assert uu._reference_file is not None and callable(uu.collate_features) and callable(uu.get_simple_loader)
    assert hasattr(dh5, "Whole_Slide_Bag_FP") and hasattr(fu, "save_hdf5")
    sys.argv[:] = ["extract_features_fp.py", "--model_type", "HIPT_4K", "--use_transforms", "HIPT", "--batch_size", "1"]
    E = importlib.import_module("extract_features_fp")
    assert E.HIPT_4K.__module__ == "hipt_abmil_atec23_b200.hipt_4k" and callable(E.compute_w_loader)
    import utils.eval_utils as EU
    assert EU.CLAM_SB.__module__ == "hipt_abmil_atec23_b200.model_clam"


def test_feature_sink_round_trip_through_the_reference_dataset(ref_env, tmp_path):
    """f2: a slide written once by feature_sink is what Generic_MIL_Dataset.__getitem__ reads (dataset_generic.py:505-528),
    and the .h5 carries `features` / `coords` under the reference's dataset names (file_utils.py:16-35)."""
    import pandas as pd
    import h5py
    from datasets.dataset_generic import Generic_MIL_Dataset
    from hipt_abmil_atec23_b200 import feature_sink
    g = torch.Generator().manual_seed(0)
    bags = {"slide_a": torch.randn(37, 192, generator=g), "slide_b": torch.randn(5, 192, generator=g)}
    coords = {k: torch.stack([torch.arange(len(v)) * 4096, torch.zeros(len(v), dtype=torch.int64)], 1) for k, v in bags.items()}
    feat_dir = str(tmp_path / "features")
    for k in bags:
        out = feature_sink.save_slide_features(feat_dir, k, bags[k], coords[k], attrs={"patch_level": 0, "patch_size": 4096})
        assert out["h5"] is not None
        assert torch.equal(feature_sink.load_slide_features(feat_dir, k), bags[k])
        with h5py.File(out["h5"], "r") as f:
            assert np.array_equal(f["features"][:], bags[k].numpy()) and np.array_equal(f["coords"][:], coords[k].numpy())
            assert f["coords"].attrs["patch_size"] == 4096
    csv = tmp_path / "set.csv"
    # integer labels: the reference's df_prep writes the mapped int back into the label column, which the image's pandas 3
    # refuses for a string column (the reference pins an older pandas)
    pd.DataFrame({"case_id": ["p0", "p1"], "slide_id": ["slide_a", "slide_b"], "label": [0, 1]}).to_csv(csv, index=False)
    ds = Generic_MIL_Dataset(data_dir=feat_dir, coords_path=None, csv_path=str(csv), shuffle=False, seed=1, print_info=False,
                             label_dict={0: 0, 1: 1}, patient_strat=False, ignore=[])
    ds.load_from_h5(False)
    for i, k in enumerate(("slide_a", "slide_b")):
        feats, label = ds[i]
        assert torch.equal(feats, bags[k]) and int(label) == i


class _FakeWSI:
    """openslide-like object: read_region((x, y), level, (w, h)) -> RGBA PIL image of deterministic noise."""

    def __init__(self, seed=0):
        self.seed = seed

    def pixels(self, coord, size):
        rs = np.random.RandomState(self.seed + int(coord[0]) * 7 + int(coord[1]) * 13)
        return rs.randint(0, 256, (size[1], size[0], 3), dtype=np.uint8)

    def read_region(self, coord, level, size):
        from PIL import Image
        rgb = self.pixels(coord, size)
        return Image.fromarray(np.concatenate([rgb, np.full(rgb.shape[:2] + (1,), 255, np.uint8)], axis=2), "RGBA")


@pytest.mark.gpu
def test_compute_w_loader_runs_unchanged_on_the_cuda_path(ref_env, tmp_path):
    """extract_features_fp.compute_w_loader — the reference's loop, dataset, DataLoader worker, collate and save_hdf5 — with
    `model` = this repository's HIPT_4K: features land in the .h5 and match the CPU oracle (cosine >= 0.999)."""
    import h5py
    from oracle import hipt_oracle as O
    from tests.common import seeded_modules
    sys.argv[:] = ["extract_features_fp.py", "--model_type", "HIPT_4K", "--use_transforms", "HIPT", "--batch_size", "1"]
    E = importlib.import_module("extract_features_fp")
    assert E.device.type == "cuda"
    coords = np.array([[0, 0], [512, 0], [0, 768]], dtype=np.int64)
    bag_h5 = str(tmp_path / "slide0.h5")
    with h5py.File(bag_h5, "w") as f:
        d = f.create_dataset("coords", data=coords)
        d.attrs["patch_level"] = 0
        d.attrs["patch_size"] = 512                                  # 2 x 2 patches of 256 per region: small and quick
    m256, m4k = seeded_modules(0)
    sd256 = {k: v.detach().clone() for k, v in m256.state_dict().items()}
    sd4k = {k: v.detach().clone() for k, v in m4k.state_dict().items()}
    model = E.HIPT_4K.from_modules(m256, m4k, torch.device("cuda:0"), torch.device("cuda:0"))
    model = model.to(E.device)
    model.eval()
    wsi = _FakeWSI(3)
    out_h5 = str(tmp_path / "features_slide0.h5")
    from hipt_abmil_atec23_b200 import _lib
    n0 = _lib.launch_count()
    path = E.compute_w_loader(bag_h5, out_h5, wsi, model=model, batch_size=1, verbose=0, print_every=20,
                              custom_downsample=1, target_patch_size=-1)
    assert _lib.launch_count() - n0 > 3 * 50                          # three regions went through the CUDA kernels
    with h5py.File(path, "r") as f:
        feats, got_coords = torch.from_numpy(f["features"][:]), f["coords"][:]
    assert feats.shape == (3, 192) and np.array_equal(got_coords, coords)
    with torch.no_grad():
        ref = torch.cat([O.hipt4k_forward(sd256, sd4k, O.eval_transforms_u8(
            torch.from_numpy(wsi.pixels(c, (512, 512))).permute(2, 0, 1)[None])) for c in coords])
    cos = F.cosine_similarity(feats.double(), ref.double(), dim=1).min().item()
    assert cos >= 0.999, cos


@pytest.mark.gpu
def test_eval_utils_initiate_model_and_summary_run_unchanged(ref_env, tmp_path):
    """utils/eval_utils.py: initiate_model cleans a checkpoint with `.module` infixes and `instance_loss_fn` entries and loads
    it strict=True into THIS repository's CLAM_SB; summary() walks a loader; probabilities match the CPU oracle."""
    import pandas as pd
    import utils.eval_utils as EU
    import utils.utils as UU
    from oracle import hipt_oracle as O
    from tests.common import seeded_clam
    src = seeded_clam("hipt_smaller", 2, 0.25)
    sd = src.state_dict()
    ckpt = {k.replace("attention_net.3", "attention_net.module.3"): v for k, v in sd.items()}
    ckpt["instance_loss_fn.weight"] = torch.zeros(1)
    ckpt_path = str(tmp_path / "s_0_checkpoint.pt")
    torch.save(ckpt, ckpt_path)
    args = argparse.Namespace(drop_out=0.25, n_classes=2, model_size="hipt_smaller", model_type="clam_sb", micro_average=False)
    model = EU.initiate_model(args, ckpt_path)
    assert type(model).__module__ == "hipt_abmil_atec23_b200.model_clam" and next(model.parameters()).is_cuda

    class Bags(torch.utils.data.Dataset):
        def __init__(self):
            g = torch.Generator().manual_seed(9)
            self.bags = [torch.randn(n, 192, generator=g) for n in (75, 200, 10, 64)]
            self.labels = [0, 1, 1, 0]
            self.slide_data = pd.DataFrame({"slide_id": [f"s{i}" for i in range(4)]})

        def __len__(self):
            return len(self.bags)

        def __getitem__(self, i):
            return self.bags[i], self.labels[i]

    ds = Bags()
    loader = torch.utils.data.DataLoader(ds, batch_size=1, collate_fn=UU.collate_MIL)
    test_error, auc, df, acc_logger, loss = EU.summary(model, loader, args)
    sd_cpu = {k: v.cpu() for k, v in sd.items()}
    for i, bag in enumerate(ds.bags):
        _, yp, yh, _, _ = O.clam_sb_forward(sd_cpu, bag)
        assert abs(df["p_1"].iloc[i] - float(yp[0, 1])) < 1e-3 and int(df["Y_hat"].iloc[i]) == int(yh)
    assert 0.0 <= test_error <= 1.0 and np.isfinite(loss)


@pytest.mark.gpu
def test_create_heatmaps_infer_single_slide_runs_unchanged(ref_env):
    """create_heatmaps.infer_single_slide (:34-57): model(features) twice, A.view(-1, 1).cpu().numpy(), topk on Y_prob."""
    from oracle import hipt_oracle as O
    from tests.common import seeded_clam
    sys.argv[:] = ["create_heatmaps.py"]
    H = importlib.import_module("create_heatmaps")
    model = seeded_clam("hipt_smaller", 2).to("cuda").eval()
    assert isinstance(model, H.CLAM_SB)
    feats = torch.randn(123, 192, generator=torch.Generator().manual_seed(4))
    ids, preds_str, probs, A = H.infer_single_slide(model, feats, "pos", {0: "neg", 1: "pos"}, k=2)
    rl, rp, rh, ra, _ = O.clam_sb_forward({k: v.cpu() for k, v in model.state_dict().items()}, feats)
    assert A.shape == (123, 1) and np.abs(A[:, 0] - ra[0].numpy()).max() < 1e-3
    assert ids[0] == int(rh) and abs(probs[0] - float(rp.max())) < 1e-3
