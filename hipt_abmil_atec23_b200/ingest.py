"""Region ingest from compressed tiles.

The reference reads every region with OpenSlide, decodes it on the CPU, converts it to a normalised fp32 tensor and ships
201 MB per region to the GPU (`Whole_Slide_Bag_FP.__getitem__`, datasets/dataset_h5.py:194-207; `collate_features`,
utils/utils.py:58-61; `batch.to(device)`, extract_features_fp.py:163).  Here a region arrives as the JPEG bytes of its tile
(the storage format of the pyramidal TIFF / SVS files OpenSlide reads), crosses PCIe compressed, and is decoded by nvJPEG
straight into the planar uint8 [R, 3, H, W] buffer that `HIPT_4K.forward_regions_u8` consumes: `hb_jpeg_decode_tiles`
(include/hipt_b200.h).  A region is a GRID of independently compressed tiles, as in the pyramidal TIFF / SVS files
themselves (256 x 256 or 512 x 512 tiles): hundreds of small bitstreams per launch are what nvJPEG's GPU Huffman stage
needs — one 4096 x 4096 bitstream per region is Huffman-decoded on one CPU thread (68 ms per region, measured) and is
kept only as the degenerate tile = region case.  There is no CPU decode path: without the CUDA library this raises.

`JpegTileBag` + `collate_jpeg` mirror the dataset / collate pair of the reference for such tiles: items are
(jpeg bytes, coord) and a batch is (list of bytes, coords [B, 2]).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

BACKENDS = {"auto": -1, "default": 0, "hybrid": 1, "gpu_hybrid": 2, "hardware": 3}


class JpegRegionDecoder:
    """nvJPEG batched decode of `max_batch` region tiles per call into a caller-visible uint8 tensor on `device`."""

    def __init__(self, device, max_batch=2, backend="auto"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("JpegRegionDecoder needs a CUDA device: this path has no CPU implementation")
        self.max_batch = int(max_batch)
        self.lib = _lib.load()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.device_check()
            _lib.check(self.lib.hb_jpeg_decoder_create(C.byref(h), self.max_batch, BACKENDS[backend]))
        self._h = h
        self.backend = self.lib.hb_jpeg_decoder_backend(h).decode()
        self._keep = None                       # bitstreams of the call in flight (the library reads them asynchronously)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.hb_jpeg_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def probe(self, blob):
        """(height, width, components, nvjpeg chroma subsampling code) of one JPEG."""
        w, h, c, s = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(self.lib.hb_jpeg_probe(self._h, blob, len(blob), C.byref(w), C.byref(h), C.byref(c), C.byref(s)))
        return h.value, w.value, c.value, s.value

    def decode(self, blobs, height, width, out=None, tile=None):
        """blobs: JPEG byte strings, region-major then row-major over each region's tile grid (tile = (tile_h, tile_w);
        None: one bitstream per region) -> uint8 [n_regions, 3, height, width] on the device (written into
        `out[:n_regions]` when given).  Work is queued on torch's current stream; the byte strings are kept alive by this
        object until the next call or `release()`."""
        th, tw = (height, width) if tile is None else tile
        per_region = (height // th) * (width // tw)
        n = len(blobs)
        if n == 0 or n > self.max_batch or n % per_region:
            raise RuntimeError(f"decode takes 1..{self.max_batch} tiles in whole regions of {per_region}, got {n}")
        R = n // per_region
        if out is None:
            out = torch.empty((R, 3, height, width), dtype=torch.uint8, device=self.device)
        assert out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and tuple(out.shape[1:]) == (3, height, width)
        assert out.shape[0] >= R
        blobs = [bytes(b) if not isinstance(b, bytes) else b for b in blobs]
        ptrs = (C.c_void_p * n)(*[C.cast(C.c_char_p(b), C.c_void_p) for b in blobs])
        lens = (C.c_size_t * n)(*[len(b) for b in blobs])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.hb_jpeg_decode_tiles(self._h, ptrs, lens, n, _lib.ptr(out), height, width, th, tw,
                                                     _lib.stream_ptr()))
        self._keep = (blobs, ptrs, lens)
        return out[:R]

    def release(self):
        """Drop the references to the last call's bitstreams (call after the stream has been synchronised)."""
        self._keep = None


class JpegTileBag(torch.utils.data.Dataset):
    """Dataset of (region tiles, coord) items — the compressed counterpart of Whole_Slide_Bag_FP
    (datasets/dataset_h5.py:151-207), whose items are (normalised fp32 image [1,3,H,W], coord).  A region's tiles are a
    list of JPEG byte strings, row-major over its tile grid (a single byte string = one tile covering the region)."""

    def __init__(self, tiles, coords):
        assert len(tiles) == len(coords)
        self.tiles = [[t] if isinstance(t, (bytes, bytearray, memoryview)) else list(t) for t in tiles]
        self.coords = np.asarray(coords).reshape(len(self.tiles), -1)

    def __len__(self):
        return len(self.tiles)

    def __getitem__(self, idx):
        return self.tiles[idx], self.coords[idx]


def collate_jpeg(batch):
    """collate_features (utils/utils.py:58-61) for compressed regions: (flat list of tile byte strings, region-major;
    coords [B, 2])."""
    return [t for item in batch for t in item[0]], np.vstack([item[1] for item in batch])
