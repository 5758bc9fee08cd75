"""Recipe for oracle/_ref: the reference's own importable modules, byte-compiled where they lie (TEST INFRASTRUCTURE).

    python oracle/build_ref.py            # in the build container, where /root/reference exists

The reference is pure Python (no native sources), so "compiling it" means byte-compiling.  Only OUTPUTS go to oracle/_ref/
(sourceless .pyc files; the directory is git-ignored and travels to the GPU box with the working tree); no reference source
is copied into the repository.  Modules: HIPT_4K/vision_transformer.py (ViT-256), HIPT_4K/vision_transformer4k.py (ViT-4K),
models/model_clam.py + models/model_mil.py (CLAM_SB / MIL_fc), utils/utils.py (initialize_weights).  HIPT_4K/hipt_4k.py and
hipt_model_utils.py cannot be byte-compiled into something importable here: hipt_model_utils.py:72 is a TabError and both
need h5py / matplotlib / skimage / webdataset; oracle/ref_runner.py restates those ~15 lines of glue.
`bench.py --impl reference` and its cpu_baseline leg time these modules (kind "reference"); without oracle/_ref they time the
oracle port (kind "port").
"""
import os
import py_compile
import shutil
import sys

REF = os.environ.get("HB_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
MODULES = ["HIPT_4K/vision_transformer.py", "HIPT_4K/vision_transformer4k.py", "models/model_clam.py", "models/model_mil.py",
           "utils/utils.py"]
# The reference's CALLERS of the hot path, for tests/test_reference_scripts.py ("the scripts run unchanged" against this
# repository's import shims, on the GPU box where /root/reference does not exist): the scripts and the modules they import.
SCRIPTS = ["extract_features_fp.py", "create_heatmaps.py", "utils/file_utils.py", "utils/core_utils.py", "utils/eval_utils.py",
           "utils/sampling_utils.py", "datasets/dataset_h5.py", "datasets/dataset_generic.py", "datasets/wsi_dataset.py",
           "models/resnet_custom.py", "wsi_core/batch_process_utils.py", "wsi_core/wsi_utils.py", "wsi_core/WholeSlideImage.py",
           "wsi_core/util_classes.py", "vis_utils/heatmap_utils.py"]


def build(quiet=False):
    if not os.path.isdir(REF):
        if not quiet:
            print(f"{REF} is not present: oracle/_ref left as it is")
        return False
    for rel in MODULES + [r for r in SCRIPTS if os.path.exists(os.path.join(REF, r))]:
        src = os.path.join(REF, rel)
        dst = os.path.join(OUT, rel[:-3] + ".pyc")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile=rel, doraise=True)
        if rel in MODULES:                                             # gpurun's snapshot drops *.pyc: the five modules the CPU
            shutil.copyfile(dst, dst[:-4] + ".refbin")                 # baseline needs travel under another extension
        init = os.path.join(os.path.dirname(dst), "__init__.pyc")    # a REGULAR package, so it wins over same-named shims
        if os.path.dirname(rel) and not os.path.exists(init):
            import tempfile
            with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as t:
                t.write("")
            py_compile.compile(t.name, cfile=init, dfile=os.path.join(os.path.dirname(rel), "__init__.py"), doraise=True)
            os.unlink(t.name)
    with open(os.path.join(OUT, "BUILT_FROM.txt"), "w") as f:
        f.write(f"byte-compiled from {REF} by oracle/build_ref.py with Python {sys.version.split()[0]}\n" + "\n".join(MODULES + SCRIPTS) + "\n")
    if not quiet:
        print("oracle/_ref:", ", ".join(MODULES))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
