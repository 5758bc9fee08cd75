#!/usr/bin/env python
"""bench.py — HIPT_4K + CLAM_SB slide-set inference throughput on B200 (BASELINE.json metric: 4K regions/s, slides/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A STEP is one pass over BASELINE.json's config 3 / 5 work list: a FIXED synthetic slide set of 2,000 uint8 4096x4096 regions
(ragged slides of 10...300 regions; region g is drawn from seed 1000 + g whatever rank owns it), every region through
ViT-256 (256 patches) and ViT-4K, every slide's [n,192] bag through a 5-fold CLAM_SB(hipt_smaller) ensemble with per-region
attention scores.  The regions are sharded over the N ranks by hipt_abmil_atec23_b200.sharding.SlideSetLayout (exactly
balanced contiguous cuts); the rows of the slides cut by a rank boundary move in ONE NCCL all-gather per step, inside the
timed region (STRONG scaling: the same 2,000 regions at every N).  `value` times K steps with each rank's regions resident in
HBM; `e2e` times the same pass from PINNED HOST memory through ShardedSlideSet.run_host (H2D of every region and D2H of the
results inside the timed region).  After the timed region every slide's bag is recomputed standalone (one slide at a time on
one rank, no sharding) and must match the sharded pass bit for bit (`bags_identical`; `bags_digest` is equal at every N).
`--impl reference` times the reference's own modules (oracle/_ref) on the host cores, one region per step.
One JSON line on stdout (rank 0).
"""
import argparse
import hashlib
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
REGION_BYTES = 3 * 4096 * 4096


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


# ------------------------------------------------------------------------------------------------------ work list
def ragged_slide_set(total, lo=10, hi=300, seed=6):
    """Slides of lo..hi regions (uniform draws, fixed seed) until `total` regions are reached; the last slide is clipped."""
    g = torch.Generator().manual_seed(seed)
    out = []
    while sum(out) < total:
        out.append(min(int(torch.randint(lo, hi + 1, (1,), generator=g)), total - sum(out)))
    return out


def uniform_slide_set(total, per_slide=50):
    out = [per_slide] * (total // per_slide)
    if total % per_slide:
        out.append(total % per_slide)
    return out


def fill_region(dst, g, dev):
    """Synthetic region g: uint8 noise from seed 1000 + g — independent of the rank that draws it."""
    gen = torch.Generator(device=dev).manual_seed(1000 + g)
    dst.random_(0, 256, generator=gen)


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


def workload_config(args, slides, world=1, layout=None):
    cfg = {"workload": f"BASELINE config 3/5: HIPT_4K (ViT-S/16 ViT-256 x 256 patches + ViT-4K) over a fixed synthetic slide set of "
                       f"{sum(slides)} uint8 4096x4096 regions in {len(slides)} ragged slides (10..300 regions, seed 6), then 5-fold "
                       f"CLAM_SB(hipt_smaller) pooling + per-region attention scores per slide; random-init weights",
           "regions_per_step": sum(slides), "slides_per_step": len(slides), "regions_per_slide": slides,
           "region_shape": [3, 4096, 4096], "clam_folds": 5, "sharding": args.policy,
           "l2": "inputs larger than L2 (50.3 MB per region, every region distinct; activations 0.5 GB per region)"}
    if layout is not None:
        load = layout.load()
        cfg.update({"regions_per_rank": load, "load_max_over_mean": max(load) / (sum(load) / len(load)),
                    "slides_cut_by_a_rank_boundary": len(layout.spanning), "allgather_rows_per_rank": layout.pad_rows})
    return cfg


# ------------------------------------------------------------------------------------------------------ CPU legs
def cpu_reference_sample(state, threads=None, regions_per_slide=50):
    """One region through the reference path on the host cores: eval_transforms + HIPT_4K.forward (unfold / rearrange copy,
    ViT-256 over 256 patches, ViT-4K) and the region's share of a 5-fold CLAM_SB pass over a `regions_per_slide` bag.
    Uses the reference's own modules from oracle/_ref when present (kind "reference"), else the oracle port (kind "port").
    Returns (seconds per region, breakdown)."""
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(1)
    region = torch.randint(0, 256, (1, 3, 4096, 4096), dtype=torch.uint8, generator=g)
    with torch.no_grad():
        if state["kind"] == "reference":
            from oracle import ref_runner as R
            m256, m4k, folds = state["models"]
            t0 = time.perf_counter()
            feat = R.hipt4k_forward(m256, m4k, R.eval_transforms_u8(region))
            t_region = time.perf_counter() - t0
            bag = feat.repeat(regions_per_slide, 1) + 0.01 * torch.randn(regions_per_slide, 192, generator=g)
            t0 = time.perf_counter()
            for f in folds:
                f(bag)
            t_clam = time.perf_counter() - t0
        else:
            from oracle import hipt_oracle as O
            sd256, sd4k, folds = state["models"]
            t0 = time.perf_counter()
            feat = O.hipt4k_forward(sd256, sd4k, O.eval_transforms_u8(region))
            t_region = time.perf_counter() - t0
            bag = feat.repeat(regions_per_slide, 1) + 0.01 * torch.randn(regions_per_slide, 192, generator=g)
            t0 = time.perf_counter()
            for sd in folds:
                O.clam_sb_forward(sd, bag)
            t_clam = time.perf_counter() - t0
    return t_region + t_clam / regions_per_slide, {"hipt4k_s_per_region": t_region, "clam5_s_per_slide": t_clam}


def cpu_reference_state():
    from oracle import ref_runner as R
    if R.available():
        return {"kind": "reference", "models": R.build_models(0)}
    from tests.common import seeded_clam, seeded_vits
    sd256, sd4k = seeded_vits(0)
    return {"kind": "port", "models": (sd256, sd4k, [seeded_clam("hipt_smaller", 10 + f).state_dict() for f in range(5)])}


def cpu_sample_text(state, cores):
    what = ("the reference's own modules byte-compiled into oracle/_ref (vision_transformer, vision_transformer4k, model_clam) + "
            "the restated hipt_4k.py:63-76 glue" if state["kind"] == "reference" else "oracle port (oracle/hipt_oracle.py)")
    return (f"{what}, CPU fp32 torch, {cores} threads; one sample = ONE synthetic 4096x4096 region through eval_transforms + "
            f"HIPT_4K.forward (256 patches, unfold copy included) + its 1/50 share of a 5-fold CLAM_SB pass over a 50-region bag")


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count()
    slides = ragged_slide_set(args.regions)
    state = cpu_reference_state()
    secs = []
    for i in range(args.warmup + args.steps):
        s, info = cpu_reference_sample(state, threads=cores)
        if i >= args.warmup:
            secs.append(s)
    value = len(secs) / sum(secs)                                    # regions per second over the timed samples
    cfg = workload_config(args, slides)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "regions/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * sum(secs) / len(secs),
            "step_is": "one bounded SAMPLE of the workload = one region (a full 2,000-region step would take "
                       f"{sum(slides) / value / 60.0:.0f} min on these cores); value = regions / measured seconds, not extrapolated",
            "regions_per_timed_step": 1,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": "regions/s", "cores": cores, "kind": state["kind"],
                             "sample": cpu_sample_text(state, cores), **info},
            "e2e": {"value": value, "unit": "regions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    rank, world, local = dist_env()
    import torch.distributed as dist
    from hipt_abmil_atec23_b200 import _lib
    from hipt_abmil_atec23_b200.pipeline import SlidePipeline
    from hipt_abmil_atec23_b200.slideset import ShardedSlideSet, all_digests

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA (B200) device: there is no CPU path for --impl ours")
    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION / WARN: send its log to stderr so that stdout
        # carries the one JSON line only (NCCL honours NCCL_DEBUG_FILE only above the VERSION level)
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    hipt, folds = build_models(dev)

    slides = ragged_slide_set(args.regions)
    total = sum(slides)
    runner = ShardedSlideSet(hipt, folds, slides, rank, world, policy=args.policy)
    lay = runner.lay
    n_local = runner.n_local
    regions = torch.empty((max(n_local, 1), 3, 4096, 4096), dtype=torch.uint8, device=dev)[:n_local]
    for i, g in enumerate(lay.regions):
        fill_region(regions[i], g, dev)
    torch.cuda.synchronize()

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

    # ---------------------------------------------------------------- value: each rank's regions resident in HBM
    for _ in range(args.warmup):
        out = runner.run_device(regions)
    torch.cuda.synchronize()
    runner.gather_events = []
    launches0 = _lib.launch_count()
    with ClockSampler(local) as clk:
        ms_total = timed(lambda: runner.run_device(regions, record_gather=True), args.steps)
    launches = _lib.launch_count() - launches0
    gather = runner.gather_ms()
    ms_step = ms_total / args.steps
    value = total * args.steps / (ms_total / 1000.0)
    out = runner.run_device(regions)
    torch.cuda.synchronize()
    feats_dev = out["features"].clone()

    # ---------------------------------------------------------------- bit-exactness: sharded pass vs every slide standalone
    sharded = all_digests(runner.bag_digests(out), world)
    alone_local = {}
    chunk = torch.empty((16, 3, 4096, 4096), dtype=torch.uint8, device=dev)
    t_check = time.perf_counter()
    for s, n in enumerate(slides):
        if n == 0 or s % world != rank:
            continue
        first = runner.layout.first_region[s]
        bag = torch.empty((n, 192), dtype=torch.float32, device=dev)
        for c0 in range(0, n, 16):
            m = min(16, n - c0)
            for j in range(m):
                fill_region(chunk[j], first + c0 + j, dev)
            bag[c0:c0 + m] = hipt.forward_regions_u8(chunk[:m])
        alone_local[s] = hashlib.sha256(bag.cpu().numpy().tobytes()).hexdigest()
    alone = all_digests(alone_local, world)
    t_check = time.perf_counter() - t_check
    bags_identical = (sorted(sharded) == sorted(alone) == [s for s, n in enumerate(slides) if n]
                      and all(sharded[s] == alone[s] for s in alone))
    bags_digest = hashlib.sha256("".join(sharded[s] for s in sorted(sharded)).encode()).hexdigest()[:16]
    del chunk

    # ---------------------------------------------------------------- e2e: pinned host -> device -> host
    P = max(1, min(n_local, args.host_pool))
    host = torch.empty((P, 3, 4096, 4096), dtype=torch.uint8).pin_memory()
    host.copy_(regions[:P])
    torch.cuda.synchronize()
    pool_index = [i % P for i in range(n_local)]
    res = runner.run_host(host, pool_index)                            # warm-up (staging buffers, copy stream)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_ms = timed(lambda: runner.run_host(host, pool_index), e2e_steps)
    e2e_value = total * e2e_steps / (e2e_ms / 1000.0)
    d2h = sum(v.numel() * v.element_size() for v in res.values() if isinstance(v, torch.Tensor))
    d2h_t = torch.tensor([d2h], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(d2h_t)
    finite = bool(torch.isfinite(res["features"]).all()) and (res["logits"] is None or bool(torch.isfinite(res["logits"]).all()))
    e2e_same = bool(torch.equal(res["features"][:P], feats_dev[:P].cpu()))     # host path == device path, bit for bit
    del host

    # ---------------------------------------------------------------- secondary lines (same regions, same models)
    uni = {}
    if total % 50 == 0 and args.secondary_steps > 0 and args.policy == "contiguous":
        u_slides = uniform_slide_set(total)                            # contiguous cuts: identical region ranges per rank
        r2 = ShardedSlideSet(hipt, folds, u_slides, rank, world, policy=args.policy)
        assert r2.lay.regions == lay.regions
        if True:
            r2.run_device(regions)
            ms = timed(lambda: r2.run_device(regions), args.secondary_steps)
            uni = {"slides": len(u_slides), "regions_per_slide": 50, "steps": args.secondary_steps,
                   "value": total * args.secondary_steps / (ms / 1000.0), "unit": "regions/s",
                   "slides_per_sec": len(u_slides) * args.secondary_steps / (ms / 1000.0),
                   "slides_cut_by_a_rank_boundary": len(r2.layout.spanning)}
        del r2
    weak = {}
    if total // world >= 16:                                           # round-1 line: 16 private regions per rank, one slide
        pipe = SlidePipeline(hipt, folds)
        for _ in range(2):
            pipe.run_device(regions[:16])
        ms = timed(lambda: pipe.run_device(regions[:16]), 6)
        weak = {"value": world * 16 * 6 / (ms / 1000.0), "unit": "regions/s", "ms_per_step": ms / 6, "regions_per_rank_per_step": 16,
                "scaling": "weak", "note": "round-1 headline: one 16-region slide per rank per step, no collective"}

    # ---------------------------------------------------------------- per-kernel times + roofline of the dominant kernel
    # hb_prof brackets every launch with two cudaEventRecords, so it is OFF for the headline above and ON for a separate
    # pass over 32 of the same regions right after it (same kernels, same launch shapes: two regions per ViT-256 launch)
    n_prof = min(n_local, 32)
    pipe = SlidePipeline(hipt, folds)
    pipe.run_device(regions[:n_prof])
    torch.cuda.synchronize()
    _lib.prof_enable(True)
    prof_steps = 2
    for _ in range(prof_steps):
        pipe.run_device(regions[:n_prof])
    prof = _lib.prof_read()
    _lib.prof_enable(False)
    step_kernel_ms = sum(ms for ms, _ in prof.values())
    kernels = {}
    for name, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        ent = {"ms_per_region": ms / (prof_steps * n_prof), "launches_per_region": cnt / (prof_steps * n_prof),
               "share": ms / step_kernel_ms, "avg_launch_us": 1000.0 * ms / cnt}
        if name in KERNEL_FLOPS_PER_REGION:        # executed FLOPs of this kernel / its measured time
            ent["tflops"] = KERNEL_FLOPS_PER_REGION[name] * n_prof * prof_steps / (ms * 1e-3) / 1e12
        if name in KERNEL_BYTES_PER_LAUNCH:
            ent["gbs"] = KERNEL_BYTES_PER_LAUNCH[name] * n_prof * prof_steps / (ms * 1e-3) / 1e9
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
                    "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                    "measured_on": f"profiled pass over {n_prof} regions x {prof_steps} right after the timed region (hb_prof on)"}
    else:
        ach = kernels[top].get("gbs", 0.0)
        roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peaks["source"]}
    model_tflops = value / world * EXECUTED_FLOPS_PER_REGION / 1e12      # executed, not the reference's dense count

    line = {"metric": METRIC, "value": value, "unit": "regions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, slides, world, runner.layout),
            "slides_per_sec": len(slides) * args.steps / (ms_total / 1000.0),
            "model_tflops_per_gpu": model_tflops, "model_frac_of_bf16_sustained": model_tflops / peaks["tflops_sustained"],
            "model_frac_dense_count": value / world * FLOPS_PER_REGION / 1e12 / peaks["tflops_sustained"],
            "flops_per_region": {"executed": EXECUTED_FLOPS_PER_REGION, "reference_dense": FLOPS_PER_REGION,
                                 "note": "last ViT-256 block computes only the CLS rows after its qkv GEMM"},
            "clocks": clk.result,
            "collective": {"kind": "NCCL all_gather_into_tensor of the cut slides' rows (fp32 [pad_rows,192] per rank), one per step"
                                   if runner.layout.needs_collective else "none needed (no slide is cut by a rank boundary)",
                           "ms_per_step_rank0": (sum(gather) / len(gather)) if gather else 0.0,
                           "ms_max_rank0": max(gather) if gather else 0.0,
                           "bytes_per_rank": runner.layout.pad_rows * 192 * 4 if runner.layout.needs_collective else 0,
                           "note": "CUDA events around the collective on the launching stream; includes waiting for the slowest rank"},
            "bags_identical": bags_identical, "bags_digest": bags_digest,
            "bags_check": f"every slide recomputed standalone (one slide at a time on rank slide % N, regions re-drawn from their "
                          f"seeds) after the timed region, {t_check:.1f} s; sha256 of each [n,192] fp32 bag",
            "e2e": {"value": e2e_value, "unit": "regions/s", "h2d_bytes_per_step": total * REGION_BYTES,
                    "d2h_bytes_per_step": int(d2h_t.item()), "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps, "finite": finite,
                    "features_equal_device_path": e2e_same,
                    "host_pool": f"{P} distinct pinned regions per rank, local region i is copied from pool[i mod {P}] every step"},
            "gpu_launches": launches, "roofline": roofline, "kernels": kernels}
    if uni:
        line["config3_uniform_40x50"] = uni
    if weak:
        line["weak_16_regions_per_rank"] = weak

    if rank == 0 and world == 1 and not args.no_sections:
        line["vit256_config2"] = vit256_config2(hipt, dev, peaks)
        line["clam_config4"] = clam_config4(dev, peaks)
        line["ingest_jpeg"] = ingest_jpeg(hipt, dev)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        state = cpu_reference_state()
        secs, info = [], {}
        for i in range(1 + args.ref_regions):                   # bounded sample: ~15-30 s of CPU work on the box's host cores
            s1, info = cpu_reference_sample(state)
            if i:
                secs.append(s1)
        v = len(secs) / sum(secs)
        line["cpu_baseline"] = {"value": v, "unit": "regions/s", "cores": os.cpu_count(), "kind": state["kind"],
                                "sample": cpu_sample_text(state, torch.get_num_threads()) + f"; 1 warm-up + {args.ref_regions} timed samples, "
                                          f"{sum(secs):.1f} s of CPU work", **info}
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


def ingest_jpeg(hipt, dev, regions=8, tile=256, quality=90):
    """the slide pipeline fed from COMPRESSED regions (grids of 256 x 256 JPEG tiles, the storage
    layout of the pyramidal TIFF / SVS files the reference reads through OpenSlide) — nvJPEG batched GPU decode into the
    staging ring (hb_jpeg_decode_tiles), next group overlapped with the current ViT-256 pass — beside the decode alone and
    the CPU decoder of the reference's loader (PIL / libjpeg, one core).  Tiles are encoded here with PIL from smooth
    synthetic regions (noise does not compress); skipped when PIL is not importable."""
    import io
    import time
    try:
        from PIL import Image
        import numpy as np
    except Exception as ex:                                  # pragma: no cover
        return {"skipped": f"PIL is not importable ({ex})"}
    from hipt_abmil_atec23_b200.ingest import JpegRegionDecoder
    from hipt_abmil_atec23_b200.pipeline import SlidePipeline
    S = 4096
    g = torch.Generator().manual_seed(77)
    grids = []
    for _ in range(2):                                       # two distinct regions, reused
        img = torch.zeros(3, S, S)
        for sc in (16, 64, 256):
            img += torch.nn.functional.interpolate(torch.rand((1, 3, S // sc, S // sc), generator=g), size=(S, S), mode="bilinear")[0]
        px = ((img / 3.0 + 0.02 * torch.randn((3, S, S), generator=g)).clamp(0, 1) * 255).round().to(torch.uint8).permute(1, 2, 0).numpy()
        grid = []
        for y in range(0, S, tile):
            for x in range(0, S, tile):
                buf = io.BytesIO()
                Image.fromarray(px[y:y + tile, x:x + tile]).save(buf, format="JPEG", quality=quality, subsampling=0)
                grid.append(buf.getvalue())
        grids.append(grid)
    per = len(grids[0])
    tiles = [b for i in range(regions) for b in grids[i % 2]]
    t0 = time.time()
    for b in grids[0][:32]:
        np.asarray(Image.open(io.BytesIO(b)).convert("RGB"))
    cpu_s = (time.time() - t0) * per / 32
    dec = JpegRegionDecoder(dev, max_batch=2 * per)
    out = torch.empty((2, 3, S, S), dtype=torch.uint8, device=dev)
    for _ in range(2):
        dec.decode(tiles[:2 * per], S, S, out=out, tile=(tile, tile))
    torch.cuda.synchronize()
    t0 = time.time()
    for r0 in range(0, regions, 2):
        dec.decode(tiles[r0 * per:(r0 + 2) * per], S, S, out=out, tile=(tile, tile))
    torch.cuda.synchronize()
    dec_rps = regions / (time.time() - t0)
    from hipt_abmil_atec23_b200.model_clam import CLAM_SB
    folds = []
    for f in range(5):
        torch.manual_seed(10 + f)
        folds.append(CLAM_SB(size_arg="hipt_smaller", dropout=0.0, n_classes=2).eval().to(dev))
    pipe = SlidePipeline(hipt, folds)
    for _ in range(2):
        pipe.run_jpeg(tiles, S, S, tile=(tile, tile))
    t0 = time.time()
    r = pipe.run_jpeg(tiles, S, S, tile=(tile, tile))
    wall = time.time() - t0
    return {"regions": regions, "tile": tile, "tiles_per_region": per, "quality": quality, "subsampling": "4:4:4", "backend": dec.backend,
            "jpeg_MB_per_region": sum(map(len, grids[0])) / 1e6, "decode_only_regions_per_s": dec_rps,
            "decode_only_decoded_GBps": dec_rps * 3 * S * S / 1e9, "pipeline_from_jpeg_regions_per_s": regions / wall,
            "h2d_MB_per_region": sum(map(len, tiles)) / regions / 1e6, "finite": bool(torch.isfinite(r["features"]).all()),
            "cpu_pil_decode_s_per_region_one_core": cpu_s,
            "note": "decode and ViT share the SMs, so the JPEG-fed pipeline runs at about 1 / (1 / decode + 1 / model); the "
                    "uint8-fed e2e line above is the headline"}


def clam_config4(dev, peaks):
    """BASELINE.json config 4 beside the headline: CLAM_SB over 256 ragged bags of 50-20,000 x 192-d instances in one launch
    (772 algorithmic bytes per instance: features read once + one score written), and the one-bag training step (forward +
    fused backward + one-launch Adam).  HBM-bound path: reported as GB/s against the measured copy peak."""
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
    from hipt_abmil_atec23_b200 import _lib
    for size, fold_list in (("hipt_smaller", (1, 5)), ("hipt_small", (1, 2)), ("hipt_medium", (1,)), ("hipt_big", (1, 5))):
        for folds in fold_list:
            models = []
            for f in range(folds):
                torch.manual_seed(10 + f)
                models.append(CLAM_SB(size_arg=size, dropout=0.0, n_classes=2).eval().to(dev))
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
            key = f"folds{folds}" if size == "hipt_smaller" else f"{size}_folds{folds}"
            out[key] = {"ms": ms, "bags_per_s": 256 / ms * 1e3, "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"]}
            # the score kernel alone (CUDA events around each launch in a second pass; `ms` above is the whole forward:
            # work table + score kernel + combine)
            _lib.prof_enable(True)
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            pr = _lib.prof_read()
            _lib.prof_enable(False)
            if "clam_scores" in pr:
                us = pr["clam_scores"][0] / pr["clam_scores"][1] * 1e3
                out[key]["score_kernel_us"] = us
                out[key]["score_kernel_GBps"] = total * (768 + 4 * folds) / us / 1e3
                out[key]["score_kernel_frac_of_hbm_peak"] = out[key]["score_kernel_GBps"] / peaks["hbm_gbs"]
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
    # T independent trials (own weights, Adam state, bag, dropout seed, lr) per fused step, at the
    # reference's training dropout 0.85 — six launches for all of them (hb_clam_sb_train_step_trials)
    for T in (5, 8):
        models = []
        for t in range(T):
            torch.manual_seed(300 + t)
            models.append(CLAM_SB(size_arg="hipt_smaller", dropout=0.85, n_classes=2).to(dev).train())
        tb = clam_engine.TrialBatchStep(models, lr=[2e-4 * (1 + t) for t in range(T)], weight_decay=1e-5, max_instances=1000)
        bags = [torch.randn((1000, 192), generator=torch.Generator().manual_seed(50 + t)).to(dev) for t in range(T)]
        labels = torch.tensor([t & 1 for t in range(T)], device=dev)
        for _ in range(5):
            tb.step(bags, labels)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(100):
            tb.step(bags, labels)
        e1.record()
        torch.cuda.synchronize()
        out[f"trial_batch_step_{T}_trials_1000_instances"] = {
            "ms": e0.elapsed_time(e1) / 100, "trial_steps_per_s": T * 100e3 / e0.elapsed_time(e1), "dropout": 0.85,
            "note": f"clam_engine.TrialBatchStep: {T} independent trials advance one train_loop step each in six launches"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--regions", type=int, default=2000, help="regions in the fixed slide set (BASELINE config 3: 2,000)")
    ap.add_argument("--policy", default="contiguous", choices=["contiguous", "lpt"], help="region sharding policy")
    ap.add_argument("--e2e-steps", type=int, default=3, help="timed passes of the pinned-host leg (at most --steps)")
    ap.add_argument("--host-pool", type=int, default=64, help="distinct pinned host regions per rank in the e2e leg")
    ap.add_argument("--secondary-steps", type=int, default=2, help="timed passes of the uniform 40 x 50 slide set")
    ap.add_argument("--ref-regions", type=int, default=4, help="timed CPU samples in the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sections", action="store_true", help="skip the config-2 / config-4 sections")
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
