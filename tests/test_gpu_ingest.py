"""Region ingest from JPEG tiles (include/hipt_b200.h hb_jpeg_*): the GPU decode against the CPU decoder
the reference's pipeline ends in (OpenSlide -> PIL / libjpeg, datasets/dataset_h5.py:194-207), and the slide pipeline fed
from compressed bytes against the same pipeline fed from the decoded pixels."""
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
PIL = pytest.importorskip("PIL.Image")
DEV = torch.device("cuda:0")


def _smooth_rgb(rows, cols, seed):
    """A tile with structure at several scales (pure noise is not what a stained slide looks like and does not compress)."""
    g = torch.Generator().manual_seed(seed)
    img = torch.zeros(3, rows, cols)
    for s in (8, 32, 128):
        low = torch.rand((1, 3, max(2, rows // s), max(2, cols // s)), generator=g)
        img += torch.nn.functional.interpolate(low, size=(rows, cols), mode="bilinear", align_corners=False)[0]
    img = img / 3.0 + 0.02 * torch.randn((3, rows, cols), generator=g)
    return (img.clamp(0, 1) * 255).round().to(torch.uint8)


def _jpeg(chw_u8, quality=90, subsampling=0):
    buf = io.BytesIO()
    PIL.fromarray(chw_u8.permute(1, 2, 0).numpy()).save(buf, format="JPEG", quality=quality, subsampling=subsampling)
    return buf.getvalue()


def _pil_decode(blob):
    return torch.from_numpy(np.asarray(PIL.open(io.BytesIO(blob)).convert("RGB")).copy()).permute(2, 0, 1).contiguous()


def _tile_grid(chw_u8, t, **kw):
    """Row-major grid of t x t JPEG tiles of one region."""
    _, rows, cols = chw_u8.shape
    return [_jpeg(chw_u8[:, y:y + t, x:x + t], **kw) for y in range(0, rows, t) for x in range(0, cols, t)]


@pytest.mark.parametrize("subsampling,tol_max,tol_mean", [(0, 6, 0.8), (2, 64, 2.5)])
def test_jpeg_decode_matches_the_cpu_decoder(subsampling, tol_max, tol_mean):
    """4:4:4 tiles: the two IDCT / colour-conversion implementations agree to a few levels (measured: max 4, mean 0.52).
    4:2:0 tiles: libjpeg's 'fancy' chroma upsampling and nvJPEG's differ at chroma edges — bounded on average, reported
    at the maximum."""
    from hipt_abmil_atec23_b200.ingest import JpegRegionDecoder
    dec = JpegRegionDecoder(DEV, max_batch=3)
    rows, cols = 512, 768
    tiles = [_smooth_rgb(rows, cols, 40 + i) for i in range(3)]
    blobs = [_jpeg(t, subsampling=subsampling) for t in tiles]
    assert dec.probe(blobs[0])[:3] == (rows, cols, 3)
    out = dec.decode(blobs, rows, cols)
    torch.cuda.synchronize()
    ref = torch.stack([_pil_decode(b) for b in blobs])
    diff = (out.cpu().int() - ref.int()).abs()
    print(f"backend {dec.backend}; subsampling {subsampling}: max |d| {diff.max().item()}, mean |d| {diff.float().mean().item():.4f}")
    assert diff.max().item() <= tol_max and diff.float().mean().item() <= tol_mean
    # a smaller batch re-initialises the batched decoder; a wrong size is refused
    one = dec.decode(blobs[1:2], rows, cols)
    torch.cuda.synchronize()
    assert torch.equal(one[0], out[1])
    with pytest.raises(RuntimeError):
        dec.decode(blobs[:1], rows, cols + 256)
    with pytest.raises(RuntimeError):
        dec.decode([b"not a jpeg"], rows, cols)


def test_tile_grid_lands_in_the_region_planes():
    """Regions stored as grids of 256 x 256 tiles (the pyramidal-TIFF layout; 2 x 6 x 9 = 108 bitstreams in one batched
    call): every tile must come out exactly where a single-image decode of that tile puts it."""
    from hipt_abmil_atec23_b200.ingest import JpegRegionDecoder
    rows, cols, t = 1536, 2304, 256
    regions = [_smooth_rgb(rows, cols, 50 + i) for i in range(2)]
    blobs = [b for r in regions for b in _tile_grid(r, t)]
    dec = JpegRegionDecoder(DEV, max_batch=len(blobs))
    out = dec.decode(blobs, rows, cols, tile=(t, t))
    torch.cuda.synchronize()
    assert tuple(out.shape) == (2, 3, rows, cols)
    one = JpegRegionDecoder(DEV, max_batch=1)
    per = (rows // t) * (cols // t)
    for i in (0, 7, per - 1, per, per + 20, 2 * per - 1):
        r, k = divmod(i, per)
        y, x = (k // (cols // t)) * t, (k % (cols // t)) * t
        ref = one.decode(blobs[i:i + 1], t, t)
        torch.cuda.synchronize()
        d = (out[r, :, y:y + t, x:x + t].int() - ref[0].int()).abs().max().item()
        assert d <= 1, (i, d)                         # the batched GPU path and the single-image path may round differently
    cpu = torch.stack([_pil_decode(b) for b in blobs[:per]])
    got = out[0].cpu().view(3, rows // t, t, cols // t, t).permute(1, 3, 0, 2, 4).reshape(per, 3, t, t)
    assert (got.int() - cpu.int()).abs().float().mean().item() <= 0.8
    with pytest.raises(RuntimeError):
        dec.decode(blobs[:per - 1], rows, cols, tile=(t, t))


def test_slide_pipeline_from_jpeg_tiles_equals_the_pipeline_from_decoded_pixels():
    """run_jpeg (decode on the GPU into the staging ring, next group overlapped with the current ViT-256 pass) returns the
    bits of run_device on the pixels nvJPEG decodes; against the pixels of the CPU decoder (the two decoders differ by
    half a grey level on average, four at most) the random-init features keep a cosine of 0.9988 (measured; bound 0.995 —
    this compares two JPEG decoders through the network, not the network's arithmetic)."""
    from hipt_abmil_atec23_b200.hipt_4k import HIPT_4K
    from hipt_abmil_atec23_b200.ingest import JpegRegionDecoder, JpegTileBag, collate_jpeg
    from hipt_abmil_atec23_b200.pipeline import SlidePipeline
    from tests.common import seeded_clam, seeded_modules
    m256, m4k = seeded_modules(0)
    hipt = HIPT_4K.from_modules(m256, m4k, DEV, DEV)
    pipe = SlidePipeline(hipt, [seeded_clam("hipt_smaller", 2).to(DEV)])
    rows, cols, R = 512, 768, 5
    tiles = [_smooth_rgb(rows, cols, 70 + i) for i in range(R)]
    bag = JpegTileBag([_tile_grid(t, 256) for t in tiles], [(i * cols, 0) for i in range(R)])
    blobs, coords = collate_jpeg([bag[i] for i in range(R)])
    assert coords.shape == (R, 2) and len(blobs) == R * 6
    got = pipe.run_jpeg(blobs, rows, cols, tile=(256, 256))
    dec = JpegRegionDecoder(DEV, max_batch=len(blobs))
    px = dec.decode(blobs, rows, cols, tile=(256, 256)).clone()
    torch.cuda.synchronize()
    want = pipe.run_device(px)
    assert torch.equal(got["features"], want["features"].cpu())
    assert torch.equal(got["a_raw"], want["a_raw"].cpu())
    cpu_px = torch.stack([_pil_decode(b) for b in blobs]).view(R, 2, 3, 3, 256, 256).permute(0, 3, 1, 4, 2, 5).reshape(R, 3, rows, cols).to(DEV)
    ref = pipe.run_device(cpu_px)["features"]
    cos = torch.nn.functional.cosine_similarity(want["features"].double(), ref.double(), dim=1).min().item()
    print(f"features from nvJPEG pixels vs libjpeg pixels: min cosine {cos:.6f}")
    assert cos >= 0.995
