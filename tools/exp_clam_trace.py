"""In-kernel clock64 trace of the CLAM tensor-core score kernel (CTA 0, first 40 tiles), experimental build only:
    tools/build_exp.sh trace   (with the TR() probes patched in: `__device__ unsigned long long g_tc_trace[64 * 16]`, a macro
    TR(tile, slot) storing clock64() for blockIdx.x == 0, and an exported hb_exp_clam_trace(out) that copies the symbol; slots:
    0 / 1 producer after the x_empty wait of slice 0 / 5, 2 / 3 MMA warp after x_full of slice 0 / 5, 4 after lo_full of slice 5,
    5 / 6 around the h_full wait in issue_gate, 8 / 9 around the epilogue's acc_full wait, 10 after the h_full arrive, 11 after
    g_done, 12 at the end of the tile)   ->   HB_LIB_PATH=.../exp_trace.so python tools/exp_clam_trace.py 5
Result of the round: profiles/r02ad_clam_trace.txt.
Columns (cycles relative to the tile's first probe): TMA slice 0 / 5 issued, MMA sees x_full of slice 0 / 5, lo_full of slice 5,
gate: before / after the h_full wait, epilogue: before / after acc_full, h_full arrive, g_done passed, tile done."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.bench_clam import bag_lengths
from hipt_abmil_atec23_b200 import clam_engine, _lib
from hipt_abmil_atec23_b200.model_clam import CLAM_SB
folds = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda:0")
lens = bag_lengths()
offs = torch.zeros(lens.numel() + 1, dtype=torch.int32); offs[1:] = torch.cumsum(lens, 0)
feats = torch.randn((int(offs[-1]), 192), generator=torch.Generator().manual_seed(5)).to(dev)
models = [CLAM_SB(size_arg="hipt_smaller", dropout=0.0, n_classes=2).eval().to(dev) for _ in range(folds)]
for _ in range(3):
    clam_engine.forward_bags(models, feats, offs.to(dev), max_bag_len=int(lens.max()))
torch.cuda.synchronize()
buf = (C.c_ulonglong * (64 * 16))()
lib = _lib.load()
lib.hb_exp_clam_trace.argtypes = [C.POINTER(C.c_ulonglong)]
assert lib.hb_exp_clam_trace(buf) == 0
t0 = buf[2 * 16 + 0]
names = ["tma0", "tma5", "x0", "x5", "lo5", "gate<", "gate>", "-", "epi<acc", "acc>", "h_arr", "g_done", "end"]
print("tile " + " ".join(f"{n:>8s}" for n in names))
for t in range(2, 40):
    row = [buf[t * 16 + k] for k in range(13)]
    print(f"{t:4d} " + " ".join(f"{(v - t0) if v else 0:8d}" for v in row))
