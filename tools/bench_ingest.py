"""Region ingest from JPEG tiles: decode-only throughput of every nvJPEG backend that comes up on this GPU,
and the slide pipeline fed from compressed bytes (SlidePipeline.run_jpeg) next to the pinned-uint8 path (run_host).
python tools/bench_ingest.py [--regions 16] [--quality 90] [--subsampling 0]   -> one JSON line"""
import argparse, io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from PIL import Image


def smooth_rgb(rows, cols, seed):
    g = torch.Generator().manual_seed(seed)
    img = torch.zeros(3, rows, cols)
    for s in (16, 64, 256):
        low = torch.rand((1, 3, rows // s, cols // s), generator=g)
        img += torch.nn.functional.interpolate(low, size=(rows, cols), mode="bilinear", align_corners=False)[0]
    img = img / 3.0 + 0.02 * torch.randn((3, rows, cols), generator=g)
    return (img.clamp(0, 1) * 255).round().to(torch.uint8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--regions", type=int, default=16)
    ap.add_argument("--distinct", type=int, default=4)
    ap.add_argument("--quality", type=int, default=90)
    ap.add_argument("--subsampling", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    from hipt_abmil_atec23_b200.ingest import JpegRegionDecoder
    from hipt_abmil_atec23_b200.hipt_4k import HIPT_4K
    from hipt_abmil_atec23_b200.pipeline import SlidePipeline
    from tests.common import seeded_clam, seeded_modules
    dev = torch.device("cuda:0")
    S = 4096
    t0 = time.time()
    pixels = [smooth_rgb(S, S, 100 + i) for i in range(a.distinct)]

    def encode(px, t):
        out = []
        for y in range(0, S, t):
            for x in range(0, S, t):
                buf = io.BytesIO()
                Image.fromarray(px[:, y:y + t, x:x + t].permute(1, 2, 0).numpy()).save(buf, format="JPEG", quality=a.quality, subsampling=a.subsampling)
                out.append(buf.getvalue())
        return out
    res = {"region": [S, S], "quality": a.quality, "subsampling": a.subsampling, "regions": a.regions, "decode": {}}
    best = None
    for t in (256, 512, 4096):
        grids = [encode(px, t) for px in pixels]
        per = len(grids[0])
        tiles = [b for i in range(a.regions) for b in grids[i % a.distinct]]
        w0 = time.time()
        for b in grids[0][:max(1, 16 * 256 * 256 // (t * t))]:
            np.asarray(Image.open(io.BytesIO(b)).convert("RGB"))
        cpu_s = (time.time() - w0) * per / max(1, 16 * 256 * 256 // (t * t))
        entry = {"tiles_per_region": per, "jpeg_MB_per_region": sum(map(len, grids[0])) / 1e6, "cpu_pil_decode_s_per_region": cpu_s, "backends": {}}
        out = torch.empty((2, 3, S, S), dtype=torch.uint8, device=dev)
        for name in ("gpu_hybrid", "default", "hardware"):
            try:
                dec = JpegRegionDecoder(dev, max_batch=2 * per, backend=name)
                for _ in range(2):
                    dec.decode(tiles[:2 * per], S, S, out=out, tile=(t, t))
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                w0 = time.time()
                e0.record()
                for r0 in range(0, a.regions, 2):
                    dec.decode(tiles[r0 * per:(r0 + 2) * per], S, S, out=out, tile=(t, t))
                e1.record(); torch.cuda.synchronize()
                wall = time.time() - w0
                ms = e0.elapsed_time(e1)
                rps = a.regions / max(ms / 1e3, wall)
                entry["backends"][name] = {"backend": dec.backend, "regions_per_s": rps, "device_ms_per_region": ms / a.regions,
                                           "wall_ms_per_region": wall * 1e3 / a.regions, "decoded_GBps": rps * 3 * S * S / 1e9}
                if best is None or rps > best[0]:
                    best = (rps, t, name, tiles, per)
                dec.close()
            except Exception as ex:
                entry["backends"][name] = {"error": str(ex)[:200]}
        res["decode"][f"tile_{t}"] = entry
    res["encode_and_decode_sweep_s"] = time.time() - t0
    if best is not None:
        _, t, name, tiles, per = best
        res["best"] = {"tile": t, "backend": name}
        m256, m4k = seeded_modules(0)
        hipt = HIPT_4K.from_modules(m256, m4k, dev, dev)
        pipe = SlidePipeline(hipt, [seeded_clam("hipt_smaller", 2 + f).to(dev) for f in range(5)])
        host = torch.stack([pixels[i % a.distinct] for i in range(a.regions)]).pin_memory()
        for fn, key in ((lambda: pipe.run_jpeg(tiles, S, S, tile=(t, t), backend=name), "pipeline_from_jpeg"), (lambda: pipe.run_host(host), "pipeline_from_pinned_uint8")):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            w0 = time.time()
            for _ in range(a.reps):
                r = fn()
            wall = (time.time() - w0) / a.reps
            res[key] = {"regions_per_s": a.regions / wall, "ms_per_slide": wall * 1e3, "h2d_MB": (sum(map(len, tiles)) if "jpeg" in key else host.numel()) / 1e6,
                        "finite": bool(torch.isfinite(r["features"]).all())}
    res["model_appetite_GBps_at_280_regions_per_s"] = 280 * 3 * S * S / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    main()
