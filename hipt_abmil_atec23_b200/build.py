"""Build libhipt_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m hipt_abmil_atec23_b200.build [--force]

The shared library is git-ignored but travels with the working tree to the GPU box.
"""
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libhipt_b200.so")
SOURCES = ["hb_api.cu", "hb_gemm.cu", "hb_mlp.cu", "hb_rowops.cu", "hb_attention.cu", "hb_attention_tc.cu", "hb_clam.cu", "hb_embed.cu", "hb_ingest.cu"]
HEADERS = ["hb_ptx.cuh", "hb_internal.h", os.path.join("..", "..", "include", "hipt_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-diag-suppress", "128",
]
OBJ_DIR = os.path.join(PKG_DIR, "lib", "obj")


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile every CUDA source into one shared library. Returns the library path."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    # one nvcc per translation unit, in parallel, then one link
    procs = []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, log = [], ""
    for src, obj, pr in procs:
        out = pr.communicate()[0]
        log += out
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + out)
        objs.append(obj)
    res = subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + objs + ["-ldl"],
                         cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(log)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
