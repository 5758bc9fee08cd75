"""Slide-level inference pipeline: uint8 regions -> HIPT_4K region embeddings -> CLAM_SB fold ensemble.

This is the call a user of the accelerated path makes for one slide (the reference spreads it over three scripts and
the file system: extract_features_fp.py:159-171 -> .h5/.pt -> eval.py / create_heatmaps.py:34-57).  `run_device` takes
regions already resident in HBM; `run_host` takes pinned host memory and overlaps the host->device copy of the next group of regions
(one ViT-256 launch = two 4096 x 4096 regions) with the ViT-256 pass of the current one on a second stream (two staging
buffers); `run_jpeg` takes the JPEG bytes of the region tiles and decodes the next group on the GPU (nvJPEG, ingest.py)
into the same staging buffers while the current group runs.
"""
import torch

from . import clam_engine
from .hipt_model_utils import HIPT_MEAN, HIPT_STD


class SlidePipeline:
    def __init__(self, hipt, clam_models, mean=HIPT_MEAN, std=HIPT_STD):
        self.hipt = hipt
        self.clam_models = list(clam_models)
        self.mean, self.std = mean, std
        self.device = torch.device(hipt.device256)
        self._stage = None
        self._copy_stream = None
        self._offs = {}
        self._decoder = None

    # ------------------------------------------------------------------------------------------------ device input
    @torch.no_grad()
    def run_device(self, regions_u8):
        """regions_u8 [R,3,4096,4096] uint8 on the GPU -> dict(features [R,192], logits [F,1,C], y_prob, y_hat, a_raw [F,R])."""
        feats = self.hipt.forward_regions_u8(regions_u8, self.mean, self.std)
        return self._pool(feats)

    def _pool(self, feats):
        R = feats.shape[0]
        offs = self._offs.get(R)
        if offs is None:                                           # built once per bag size, on the device
            offs = self._offs[R] = torch.tensor([0, R], dtype=torch.int32, device=feats.device)
        r = clam_engine.forward_bags(self.clam_models, feats, offs, max_bag_len=R, want=("logits", "y_prob", "y_hat"))
        r["features"] = feats
        return r

    # -------------------------------------------------------------------------------------------------- host input
    @torch.no_grad()
    def run_host(self, regions_u8_pinned):
        """Same as run_device for a pinned HOST tensor; returns host copies of features / logits / y_prob / y_hat / a_raw.
        Bytes moved per call: R * 3 * H * W in, R * 192 * 4 + small out."""
        assert not regions_u8_pinned.is_cuda and regions_u8_pinned.dtype == torch.uint8
        R, _, W, H = regions_u8_pinned.shape

        def fill(stage, r0, n):
            stage[:n].copy_(regions_u8_pinned[r0:r0 + n], non_blocking=True)
        return self._run_staged(R, W, H, fill)

    @torch.no_grad()
    def run_jpeg(self, tiles, rows, cols, tile=None, backend="auto"):
        """Same as run_host for regions of `rows` x `cols` pixels (multiples of 256) stored as JPEG: `tiles` is the flat
        list of byte strings, region-major then row-major over each region's grid of tile = (tile_h, tile_w) tiles (None:
        one bitstream per region).  The compressed bytes are what crosses PCIe; nvJPEG decodes group g + 1 into the
        second staging buffer while ViT-256 runs group g.  Bytes moved per call: sum(len(t)) in."""
        from .ingest import JpegRegionDecoder
        th, tw = (rows, cols) if tile is None else tile
        per = (rows // th) * (cols // tw)
        assert len(tiles) % per == 0
        R = len(tiles) // per
        eng = self.hipt.model256._engine(self.device)
        k = max(1, eng.max_seqs // ((rows // 256) * (cols // 256)))
        if self._decoder is None or self._decoder.max_batch < k * per:
            self._decoder = JpegRegionDecoder(self.device, max_batch=k * per, backend=backend)
        tiles = [bytes(t) if not isinstance(t, bytes) else t for t in tiles]      # alive until the final synchronize

        def fill(stage, r0, n):
            self._decoder.decode(tiles[r0 * per:(r0 + n) * per], rows, cols, out=stage, tile=(th, tw))
        out = self._run_staged(R, rows, cols, fill)
        self._decoder.release()
        return out

    def _run_staged(self, R, W, H, fill):
        """Groups of k regions through two staging buffers: fill(stage_buffer, first_region, n) runs on the copy stream."""
        dev = self.device
        T = (W // 256) * (H // 256)
        with torch.cuda.device(dev):
            eng = self.hipt.model256._engine(dev)
            k = max(1, eng.max_seqs // T)                       # regions per ViT-256 launch (two 4096x4096 regions)
            shape = (2, k, 3, W, H)
            if self._stage is None or tuple(self._stage.shape) != shape:
                self._stage = torch.empty(shape, dtype=torch.uint8, device=dev)
                self._copy_stream = torch.cuda.Stream(device=dev)
            main = torch.cuda.current_stream(dev)
            cls_bf16 = torch.empty((R * T, eng.dim), dtype=torch.bfloat16, device=dev)
            copied = [torch.cuda.Event() for _ in range(2)]
            consumed = [torch.cuda.Event() for _ in range(2)]
            for g, r0 in enumerate(range(0, R, k)):
                b = g & 1
                n = min(k, R - r0)
                with torch.cuda.stream(self._copy_stream):
                    if g >= 2:
                        self._copy_stream.wait_event(consumed[b])
                    elif g == 0:                                   # g == 1: staging buffer 1 has no consumer yet in this call
                        self._copy_stream.wait_stream(main)        # orders against the previous call's use of the buffers
                    fill(self._stage[b], r0, n)
                    copied[b].record(self._copy_stream)
                main.wait_event(copied[b])
                eng.forward_patches(self._stage[b, :n], mean=self.mean, std=self.std, want_f32=False,
                                    out_bf16=cls_bf16[r0 * T:(r0 + n) * T])
                consumed[b].record(main)
            feats = self.hipt.model4k._engine(dev).forward_grid(cls_bf16, R, W // 256, H // 256)
            out = self._pool(feats)
            host = {k: v.to("cpu", non_blocking=True) for k, v in out.items() if v is not None}
            main.synchronize()
        return host
