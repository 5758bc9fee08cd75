#!/usr/bin/env python
"""bench.py — HIPT_4K + CLAM_SB slide inference throughput on B200 (BASELINE.json metric: 4K regions/s, slides/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A STEP is one synthetic slide: `--regions-per-step` uint8 4096x4096 regions -> ViT-256 over their 256 patches ->
ViT-4K -> [R,192] bag -> 5-fold CLAM_SB(hipt_smaller) ensemble (config 3/5 of BASELINE.json at one slide per step).
`value` times K steps with the slide resident in HBM; `e2e` times the same K steps from PINNED HOST memory through
hipt_abmil_atec23_b200.pipeline.SlidePipeline.run_host (H2D of every region and D2H of the results inside the timed
region).  Multi-GPU: one process per GPU, each rank owns whole slides (no data-path collective), weak scaling, time =
max over ranks.  `--impl reference` times the CPU oracle port of the reference path on the host cores.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# algorithmic work (SURVEY.md §8d / BASELINE.md §3), multiply-add = 2
FLOPS_PER_REGION = 3_146_029_797_888          # dense count the reference executes (256 x ViT-256 + ViT-4K)
M_TOK = 256 * 257
FULL = {            # one ViT-256 launch over one region (256 patches), all tokens
    "qkv_gemm": 2 * M_TOK * 1152 * 384,
    "proj_gemm": 2 * M_TOK * 384 * 384,
    "fc1_gemm": 2 * M_TOK * 1536 * 384,
    "fc2_gemm": 2 * M_TOK * 384 * 1536,
    "attention": 2 * 2 * 257 * 257 * 64 * 6 * 256,
    "embed_gemm": 2 * 65536 * 768 * 384,
}
CLS = {             # the last block after its qkv GEMM: CLS rows only (forward() returns x[:, 0])
    "proj_gemm": 2 * 256 * 384 * 384, "fc1_gemm": 2 * 256 * 1536 * 384, "fc2_gemm": 2 * 256 * 384 * 1536,
    "attention": 2 * 2 * 1 * 257 * 64 * 6 * 256,
}
# executed FLOPs per region and kernel: 12 blocks, of which the last is CLS-only for everything but qkv; blocks 1-11 run
# fc1 + GELU + fc2 + residual as ONE kernel (mlp_fused), the CLS-only tail of block 12 as two small GEMMs
KERNEL_FLOPS_PER_REGION = {
    "qkv_gemm": 12 * FULL["qkv_gemm"],
    "embed_gemm": FULL["embed_gemm"],
    "attention": 11 * FULL["attention"] + CLS["attention"],
    "proj_gemm": 11 * FULL["proj_gemm"] + CLS["proj_gemm"],
    "mlp_fused": 11 * (FULL["fc1_gemm"] + FULL["fc2_gemm"]),
    "fc1_gemm": CLS["fc1_gemm"],
    "fc2_gemm": CLS["fc2_gemm"],
}
VIT4K_FLOPS_PER_REGION = 1_706_365_440
EXECUTED_FLOPS_PER_REGION = sum(KERNEL_FLOPS_PER_REGION.values()) + VIT4K_FLOPS_PER_REGION
KERNEL_BYTES_PER_LAUNCH = {            # HBM-bound row kernels: bytes that must move per region
    "im2col": 3 * 4096 * 4096 * (1 + 2),
}
METRIC = "4K regions/sec (HIPT_4K extraction + CLAM_SB 5-fold pooling)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"tflops_sustained": d.get("bf16_tflops_sustained", 1400.0), "tflops_burst": d.get("bf16_tflops", 1590.0),
                "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        self.result = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            return
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            busy = sorted(sm)[len(sm) // 2:]            # upper half = samples under load
            self.result = {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                           "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def build_models(device, seed=0):
    from hipt_abmil_atec23_b200 import vision_transformer as vits
    from hipt_abmil_atec23_b200 import vision_transformer4k as vits4k
    from hipt_abmil_atec23_b200.hipt_4k import HIPT_4K
    from hipt_abmil_atec23_b200.model_clam import CLAM_SB
    torch.manual_seed(seed)
    m256 = vits.vit_small(patch_size=16, num_classes=0)
    m4k = vits4k.vit4k_xs(num_classes=0)
    hipt = HIPT_4K.from_modules(m256, m4k, device, device)
    folds = []
    for f in range(5):
        torch.manual_seed(10 + f)
        folds.append(CLAM_SB(size_arg="hipt_smaller", dropout=0.0, n_classes=2).eval().to(device))
    return hipt, folds


# ------------------------------------------------------------------------------------------------------ CPU legs
def cpu_reference_sample(patches=32, regions_per_slide=50, threads=None):
    """Oracle port of the reference path on the host cores, on a bounded sample: ViT-256 over `patches` of one synthetic
    region (scaled to 256), ViT-4K on one 16x16 grid, CLAM_SB 5-fold on one bag.  Returns (regions/s, seconds spent, info)."""
    from oracle import hipt_oracle as O
    from tests.common import seeded_clam, seeded_vits
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    sd256, sd4k = seeded_vits(0)
    folds = [seeded_clam("hipt_smaller", 10 + f).state_dict() for f in range(5)]
    g = torch.Generator().manual_seed(1)
    px = torch.randint(0, 256, (patches, 3, 256, 256), dtype=torch.uint8, generator=g)
    t_start = time.perf_counter()
    with torch.no_grad():
        x = O.eval_transforms_u8(px)
        O.vit256_forward(sd256, x[:4])                                   # warm-up (thread pool, allocator)
        t0 = time.perf_counter()
        cls = O.vit256_forward(sd256, x)
        t256 = time.perf_counter() - t0
        grid = cls.repeat(256 // patches + 1, 1)[:256].reshape(16, 16, 384).transpose(0, 1).transpose(0, 2).unsqueeze(0)
        O.vit4k_forward(sd4k, grid)
        t0 = time.perf_counter()
        feat = O.vit4k_forward(sd4k, grid)
        t4k = time.perf_counter() - t0
        bag = feat.repeat(regions_per_slide, 1) + 0.01 * torch.randn(regions_per_slide, 192, generator=g)
        t0 = time.perf_counter()
        for sd in folds:
            O.clam_sb_forward(sd, bag)
        tclam = time.perf_counter() - t0
    per_region = t256 * (256.0 / patches) + t4k + tclam / regions_per_slide
    info = {"vit256_s_per_region": t256 * 256.0 / patches, "vit4k_s_per_region": t4k, "clam_s_per_slide": tclam}
    return 1.0 / per_region, time.perf_counter() - t_start, info


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count()
    vals, spent = [], 0.0
    for i in range(args.warmup + args.steps):
        v, s, info = cpu_reference_sample(patches=args.ref_patches, regions_per_slide=args.regions_per_step, threads=cores)
        spent += s
        if i >= args.warmup:
            vals.append(v)
    value = len(vals) / sum(1.0 / v for v in vals)                        # steps / total time
    sample = (f"oracle port (CPU fp32 torch, {cores} threads) per step: ViT-256 on {args.ref_patches} of 256 patches scaled x{256 // args.ref_patches}, "
              f"ViT-4K on one grid, CLAM_SB 5-fold on one {args.regions_per_step}-region bag")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "regions/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * args.regions_per_step / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": "regions/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "regions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "HIPT_4K(ViT-S/16 ViT-256 x256 patches + ViT-4K) on synthetic uint8 4096x4096 regions, then "
                        "CLAM_SB(hipt_smaller) 5-fold gated-attention pooling per slide; random-init weights",
            "regions_per_step": args.regions_per_step, "slides_per_step": 1,
            "region_shape": [3, 4096, 4096], "clam_folds": 5,
            "l2": "inputs larger than L2 (50.3 MB/region x regions_per_step, activations 0.5 GB/region)"}


# ------------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    rank, world, local = dist_env()
    import torch.distributed as dist
    from hipt_abmil_atec23_b200 import _lib
    from hipt_abmil_atec23_b200.pipeline import SlidePipeline

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA (B200) device: there is no CPU path for --impl ours")
    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION / WARN: send its log to stderr so that stdout
        # carries the one JSON line only
        # (NCCL honours NCCL_DEBUG_FILE only above the VERSION level)
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    R = args.regions_per_step
    hipt, folds = build_models(dev)
    pipe = SlidePipeline(hipt, folds)

    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    regions = torch.randint(0, 256, (R, 3, 4096, 4096), dtype=torch.uint8, device=dev, generator=gen)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """Exactly `steps` calls between CUDA events on the launching stream; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return ms.item()

    # ---------------------------------------------------------------- value: inputs resident in HBM
    for _ in range(args.warmup):
        out = pipe.run_device(regions)
    torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    _lib.prof_enable(True)
    with ClockSampler(local) as clk:
        ms_total = timed(lambda: pipe.run_device(regions), args.steps)
    prof = _lib.prof_read()
    _lib.prof_enable(False)
    launches = _lib.launch_count() - launches0
    ms_step = ms_total / args.steps
    value = world * R * args.steps / (ms_total / 1000.0)

    # ---------------------------------------------------------------- e2e: pinned host -> device -> host
    host = torch.empty((R, 3, 4096, 4096), dtype=torch.uint8).pin_memory()
    host.copy_(regions)
    torch.cuda.synchronize()
    for _ in range(max(1, args.warmup // 2)):
        res = pipe.run_host(host)
    e2e_ms = timed(lambda: pipe.run_host(host), args.steps)
    e2e_value = world * R * args.steps / (e2e_ms / 1000.0)
    d2h = sum(v.numel() * v.element_size() for v in res.values())
    ok = bool(torch.isfinite(res["features"]).all() and torch.isfinite(res["logits"]).all())

    # ---------------------------------------------------------------- roofline of the dominant kernel (rank 0's events)
    step_kernel_ms = sum(ms for ms, _ in prof.values())
    kernels = {}
    for name, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        ent = {"ms_per_step": ms / args.steps, "launches_per_step": cnt / args.steps, "share": ms / step_kernel_ms,
               "avg_launch_us": 1000.0 * ms / cnt}
        if name in KERNEL_FLOPS_PER_REGION:        # executed FLOPs of this kernel per step / its measured time per step
            ent["tflops"] = KERNEL_FLOPS_PER_REGION[name] * R / (ms / args.steps * 1e-3) / 1e12
        if name in KERNEL_BYTES_PER_LAUNCH:       # bytes per region; one timed scope covers the regions of one ViT-256 launch
            ent["gbs"] = KERNEL_BYTES_PER_LAUNCH[name] * R * args.steps / (ms * 1e-3) / 1e9
        kernels[name] = ent
    top = next(iter(kernels))
    traffic = None                      # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(top, {}).get("dram_bytes_per_launch")
    except Exception:
        traffic = None
    if top in KERNEL_FLOPS_PER_REGION:
        ach = kernels[top]["tflops"]
        roofline = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach / peaks["tflops_sustained"], "traffic": traffic,
                    "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)"}
    else:
        ach = kernels[top].get("gbs", 0.0)
        roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peaks["source"]}
    model_tflops = value / world * EXECUTED_FLOPS_PER_REGION / 1e12      # executed, not the reference's dense count

    line = {"metric": METRIC, "value": value, "unit": "regions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args),
            "slides_per_sec": value / R,
            "model_tflops_per_gpu": model_tflops, "model_frac_of_bf16_sustained": model_tflops / peaks["tflops_sustained"],
            "flops_per_region": {"executed": EXECUTED_FLOPS_PER_REGION, "reference_dense": FLOPS_PER_REGION,
                                 "note": "last ViT-256 block computes only the CLS rows after its qkv GEMM"},
            "clocks": clk.result,
            "e2e": {"value": e2e_value, "unit": "regions/s", "h2d_bytes_per_step": R * 3 * 4096 * 4096,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps, "finite": ok},
            "gpu_launches": launches, "roofline": roofline, "kernels": kernels}

    if rank == 0 and world == 1:
        line["vit256_config2"] = vit256_config2(hipt, dev, peaks)
        line["clam_config4"] = clam_config4(dev, peaks)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        vals, spent, info = [], 0.0, {}
        for _ in range(args.ref_regions):                   # bounded sample: ~10-30 s of CPU work on the box's host cores
            v, s1, info = cpu_reference_sample(patches=args.ref_patches, regions_per_slide=R)
            vals.append(v)
            spent += s1
        v = len(vals) / sum(1.0 / x for x in vals)
        line["cpu_baseline"] = {"value": v, "unit": "regions/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"oracle port (CPU fp32 torch, {torch.get_num_threads()} threads): {args.ref_regions} x (ViT-256 on "
                                          f"{args.ref_patches}/256 patches of one region, scaled to 256; ViT-4K on one grid; CLAM 5-fold on "
                                          f"one {R}-region bag); {spent:.1f} s of CPU work", **info}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def vit256_config2(hipt, dev, peaks):
    """BASELINE.json config 2: the ViT-256 (ViT-S/16) stage alone on a batch of 256 synthetic 256 x 256 patches (uint8,
    normalisation folded into the patch embed), bf16; patches/s and executed TFLOP/s against the sustained bf16 peak."""
    from hipt_abmil_atec23_b200.hipt_model_utils import HIPT_MEAN, HIPT_STD
    eng = hipt.model256._engine(dev)
    px = torch.randint(0, 256, (256, 3, 256, 256), dtype=torch.uint8, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    fn = lambda: eng.forward_patches(px, mean=HIPT_MEAN, std=HIPT_STD, want_f32=False)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    flops = EXECUTED_FLOPS_PER_REGION - VIT4K_FLOPS_PER_REGION
    return {"batch": 256, "ms": ms, "patches_per_s": 256 / ms * 1e3, "tflops_executed": flops / ms / 1e9,
            "frac_of_bf16_sustained": flops / ms / 1e9 / peaks["tflops_sustained"],
            "note": "one launch sequence over 256 patches (half the two-region launch the slide path uses)"}


def clam_config4(dev, peaks):
    """BASELINE.json config 4 beside the headline: CLAM_SB(hipt_smaller) over 256 ragged bags of 50-20,000 x 192-d instances
    in one launch (772 algorithmic bytes per instance: features read once + one score written), and the one-bag training
    step (forward + fused backward + one-launch Adam).  HBM-bound path: reported as GB/s against the measured copy peak."""
    import math
    import torch.nn.functional as F
    from hipt_abmil_atec23_b200 import clam_engine
    from hipt_abmil_atec23_b200.model_clam import CLAM_SB
    g = torch.Generator().manual_seed(4)
    lens = torch.exp(math.log(50) + torch.rand(256, generator=g) * (math.log(20000) - math.log(50))).long().clamp(50, 20000)
    offs = torch.zeros(257, dtype=torch.int32)
    offs[1:] = torch.cumsum(lens, 0)
    total = int(offs[-1])
    feats = torch.randn((total, 192), generator=torch.Generator().manual_seed(5)).to(dev)
    offs_d, mx = offs.to(dev), int(lens.max())
    out = {"bags": 256, "instances": total, "bytes_per_instance": 772, "feature_MB": total * 768 / 1e6}
    for folds in (1, 5):
        models = []
        for f in range(folds):
            torch.manual_seed(10 + f)
            models.append(CLAM_SB(size_arg="hipt_smaller", dropout=0.0, n_classes=2).eval().to(dev))
        fn = lambda: clam_engine.forward_bags(models, feats, offs_d, max_bag_len=mx, want=("logits", "y_prob", "y_hat"))
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gbs = total * (768 + 4 * folds) / ms / 1e6
        out[f"folds{folds}"] = {"ms": ms, "bags_per_s": 256 / ms * 1e3, "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"]}
    torch.manual_seed(2)
    model = CLAM_SB(size_arg="hipt_smaller", dropout=0.0, n_classes=2).to(dev).train()
    opt = clam_engine.FusedAdam(filter(lambda p: p.requires_grad, model.parameters()), lr=2e-4, weight_decay=1e-5)
    bag, label = feats[:1000].contiguous(), torch.tensor([1], device=dev)

    def step():
        logits = model(bag)[0]
        F.cross_entropy(logits, label).backward()
        opt.step()
        opt.zero_grad()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        step()
    e1.record()
    torch.cuda.synchronize()
    out["train_step_1000_instances"] = {"ms": e0.elapsed_time(e1) / 30, "steps_per_s": 30e3 / e0.elapsed_time(e1),
                                        "note": "one bag per step as in train_loop (model(bag) -> CE -> backward -> Adam through "
                                                "autograd): launch / host-latency bound"}
    torch.manual_seed(2)
    model2 = CLAM_SB(size_arg="hipt_smaller", dropout=0.0, n_classes=2).to(dev).train()
    ts = clam_engine.TrainStep(model2, clam_engine.FusedAdam(clam_engine._param_list(model2), lr=2e-4, weight_decay=1e-5), 1000)
    for _ in range(5):
        ts.step(bag, label)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(100):
        ts.step(bag, label)
    e1.record()
    torch.cuda.synchronize()
    out["lean_train_step_1000_instances"] = {"ms": e0.elapsed_time(e1) / 100, "steps_per_s": 100e3 / e0.elapsed_time(e1),
                                             "note": "clam_engine.TrainStep: the same step as 6 launches, cross-entropy fused "
                                                     "into the backward, no autograd"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--regions-per-step", type=int, default=16, help="regions per synthetic slide (one slide per step)")
    ap.add_argument("--ref-patches", type=int, default=256, help="patches per CPU-oracle sample (of 256 per region)")
    ap.add_argument("--ref-regions", type=int, default=4, help="CPU-oracle samples in the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-warmup", type=int, default=3, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "ours":
        args.warmup = max(args.warmup, args.min_warmup)      # timing hygiene: at least 3 untimed steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
