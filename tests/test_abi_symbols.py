"""The C-ABI library loads without a GPU and exports every symbol include/hipt_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "hipt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from hipt_abmil_atec23_b200 import _lib, build
    build.build()                                     # no-op when the in-tree .so is current
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert set(_lib.SIGNATURES) == set(names)         # the ctypes table binds exactly the declared surface


def test_abi_version_and_error_string():
    from hipt_abmil_atec23_b200 import _lib
    lib = _lib.load()
    assert lib.hb_abi_version() == 4
    assert isinstance(lib.hb_last_error(), bytes)
    # [prefix: n_bags + 1 ints][work: 2 ints per (bag, chunk) item][partials: n_models x items x (L1 + 2) floats], items bounded
    # with the smallest chunk (32 instances); each part 16-byte aligned
    items = 256 * ((20000 + 31) // 32)
    al = lambda v: (v + 15) & ~15
    assert lib.hb_clam_workspace_bytes(20000, 256, 5, 16) == al(al(257 * 4) + 2 * items * 4) + 5 * items * 18 * 4


def test_config_struct_layout_matches_header():
    from hipt_abmil_atec23_b200._lib import HbVitConfig
    assert [f[0] for f in HbVitConfig._fields_] == ["dim", "heads", "depth", "mlp_dim", "max_rows", "ln_eps"]
    assert ctypes.sizeof(HbVitConfig) == 24
    lib = __import__("hipt_abmil_atec23_b200._lib", fromlist=["load"]).load()
    cfg = HbVitConfig(384, 6, 12, 1536, 256 * 257, 1e-6)
    rows = 65792                                       # = 257 * 256: already a multiple of the 256-row pair tile
    # bf16 residual stream + qkv + attention out + MLP hidden + two sets of 6 statistics planes
    expect = rows * 384 * 2 + rows * 384 * 6 + rows * 384 * 2 + rows * 1536 * 2 + 2 * 6 * rows * 8
    assert lib.hb_vit_workspace_bytes(ctypes.byref(cfg)) == expect
