"""Whole-slide-set inference across the GPUs of one box (BASELINE.json configs 3 and 5).

The reference walks a slide set through three scripts and the file system: extract_features_fp.py:223-255 loops over the
slides, runs HIPT_4K per region and appends to an .h5 per slide; eval.py / create_heatmaps.py:34-57 re-read the bags and
pool them with every fold's CLAM_SB.  Here one process per GPU holds a shard of the REGIONS (sharding.SlideSetLayout):

    regions (uint8, HBM or pinned host) --ViT-256--> CLS grid --ViT-4K--> [n_local, 192] feature buffer
        --one NCCL all-gather of the rows of slides cut by a rank boundary-->
        CLAM_SB fold ensemble over every bag this rank owns, in ONE launch --> logits / probabilities / per-region scores

Feature, gather and CLAM-input buffers, bag offsets and gather indices are built once; a pass only draws its small
CLAM output tensors from torch's caching allocator.
"""
import hashlib

import torch
import torch.distributed as dist

from . import clam_engine
from .hipt_model_utils import HIPT_MEAN, HIPT_STD
from .sharding import SlideSetLayout, assemble_owned_bags, gather_span_rows


class ShardedSlideSet:
    def __init__(self, hipt, clam_models, regions_per_slide, rank=0, world_size=1, group=None, policy="contiguous",
                 region_shape=(3, 4096, 4096), mean=HIPT_MEAN, std=HIPT_STD):
        self.hipt = hipt
        self.clam_models = list(clam_models)
        self.rank, self.world, self.group = rank, world_size, group
        self.mean, self.std = mean, std
        self.layout = SlideSetLayout(regions_per_slide, world_size, policy)
        self.lay = self.layout.ranks[rank]
        self.device = torch.device(hipt.device256)
        self.region_shape = tuple(region_shape)
        C, W, H = self.region_shape
        if W % 256 or H % 256:
            raise RuntimeError("regions must be cropped to multiples of 256")
        self.w256, self.h256 = W // 256, H // 256
        self.T = self.w256 * self.h256
        dev = self.device
        F = hipt.model4k.embed_dim
        self.F = F
        n_local, pad = len(self.lay.regions), self.layout.pad_rows
        self.n_local = n_local
        self.cap = max(n_local, 1)                            # rows of the local feature buffer (head of `pool`)
        with torch.cuda.device(dev):
            self.eng256 = hipt.model256._engine(dev)
            self.eng4k = hipt.model4k._engine(dev)
            self.cls_buf = torch.empty((self.cap * self.T, self.eng256.dim), dtype=torch.bfloat16, device=dev)
            gather_rows = world_size * pad if self.layout.needs_collective else 0
            self.pool = torch.zeros((self.cap + gather_rows, F), dtype=torch.float32, device=dev)
            self.feats_buf = self.pool[:self.cap]
            self.send = torch.zeros((max(pad, 1), F), dtype=torch.float32, device=dev)[:pad]
            self.send_index = torch.tensor(self.lay.send_index, dtype=torch.int64, device=dev)
            self.pool_rows = torch.tensor(self.lay.pool_rows(self.cap), dtype=torch.int64, device=dev)
            self.bag_offsets = torch.tensor(self.lay.bag_offsets, dtype=torch.int32, device=dev)
            self.n_owned_rows = self.lay.bag_offsets[-1]
            self.clam_in = torch.empty((max(self.n_owned_rows, 1), F), dtype=torch.float32, device=dev)
            lens = [b - a for a, b in zip(self.lay.bag_offsets[:-1], self.lay.bag_offsets[1:])]
            self.max_bag_len = max(lens) if lens else 0
            self._stage = None
            self._copy_stream = None
        self.gather_events = []               # (start, stop) CUDA-event pairs of the passes run with record_gather=True

    # ----------------------------------------------------------------------------------------------- bag assembly
    def _assemble_and_pool(self, record_gather=False):
        lay = self.lay
        if self.layout.needs_collective:
            if record_gather:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if lay.send_index:
                torch.index_select(self.feats_buf, 0, self.send_index, out=self.send[:len(lay.send_index)])
            gather_span_rows(self.send, self.pool[self.cap:], self.group)
            if record_gather:
                e1.record()
                self.gather_events.append((e0, e1))
        if self.n_owned_rows == 0:
            return {"features": self.feats_buf[:self.n_local], "owned": [], "logits": None, "y_prob": None, "y_hat": None,
                    "a_raw": None, "prob_median": None}
        assemble_owned_bags(self.layout, lay, self.pool, self.cap, self.send, self.send_index, self.pool_rows, self.clam_in,
                            self.group, collective=False)
        r = clam_engine.forward_bags(self.clam_models, self.clam_in[:self.n_owned_rows], self.bag_offsets,
                                     max_bag_len=self.max_bag_len, want=("logits", "y_prob", "y_hat"))
        r["prob_median"] = r["y_prob"].median(dim=0).values        # eval.py ensembles the folds' probabilities per slide
        r["features"] = self.feats_buf[:self.n_local]
        r["bags"] = self.clam_in[:self.n_owned_rows]
        r["owned"] = lay.owned
        return r

    # ------------------------------------------------------------------------------------------------ device input
    @torch.no_grad()
    def run_device(self, regions_u8, record_gather=False):
        """regions_u8 [n_local, 3, W, H] uint8 on this rank's GPU, in self.lay.regions order."""
        if regions_u8.shape[0] != self.n_local:
            raise RuntimeError(f"rank {self.rank} expects {self.n_local} regions, got {regions_u8.shape[0]}")
        with torch.cuda.device(self.device):
            if self.n_local:
                self.eng256.forward_patches(regions_u8, mean=self.mean, std=self.std, want_f32=False, out_bf16=self.cls_buf)
                self.eng4k.forward_grid(self.cls_buf, self.n_local, self.w256, self.h256, out=self.feats_buf)
            return self._assemble_and_pool(record_gather)

    # -------------------------------------------------------------------------------------------------- host input
    @torch.no_grad()
    def run_host(self, host_pool, pool_index=None, record_gather=False):
        """The same pass with the regions in PINNED HOST memory: local region i is host_pool[pool_index[i]] (identity if
        None).  The host->device copy of launch group g+1 overlaps the ViT-256 pass of group g (two staging buffers, a copy
        stream); the results come back to the host inside the call.  Bytes in: n_local * 3 * W * H."""
        assert not host_pool.is_cuda and host_pool.dtype == torch.uint8
        dev = self.device
        n = self.n_local
        idx = list(range(n)) if pool_index is None else list(pool_index)
        with torch.cuda.device(dev):
            k = max(1, self.eng256.max_seqs // self.T)            # regions per ViT-256 launch
            shape = (2, k) + self.region_shape
            if self._stage is None or tuple(self._stage.shape) != shape:
                self._stage = torch.empty(shape, dtype=torch.uint8, device=dev)
                self._copy_stream = torch.cuda.Stream(device=dev)
                self._copied = [torch.cuda.Event() for _ in range(2)]
                self._consumed = [torch.cuda.Event() for _ in range(2)]
            main = torch.cuda.current_stream(dev)
            cs = self._copy_stream
            for g, r0 in enumerate(range(0, n, k)):
                b = g & 1
                m = min(k, n - r0)
                with torch.cuda.stream(cs):
                    if g >= 2:
                        cs.wait_event(self._consumed[b])
                    elif g == 0:
                        cs.wait_stream(main)                      # orders against the previous call's use of the buffers
                    for j in range(m):
                        self._stage[b, j].copy_(host_pool[idx[r0 + j]], non_blocking=True)
                    self._copied[b].record(cs)
                main.wait_event(self._copied[b])
                self.eng256.forward_patches(self._stage[b, :m], mean=self.mean, std=self.std, want_f32=False,
                                            out_bf16=self.cls_buf[r0 * self.T:(r0 + m) * self.T])
                self._consumed[b].record(main)
            if n:
                self.eng4k.forward_grid(self.cls_buf, n, self.w256, self.h256, out=self.feats_buf)
            out = self._assemble_and_pool(record_gather)
            host = {k_: v.to("cpu", non_blocking=True) for k_, v in out.items()
                    if isinstance(v, torch.Tensor) and k_ != "bags"}
            main.synchronize()
            host["owned"] = out["owned"]
        return host

    # --------------------------------------------------------------------------------------------------- checking
    def gather_ms(self):
        """Milliseconds spent inside the collective per recorded pass (includes waiting for the slowest rank)."""
        torch.cuda.synchronize(self.device)
        ms = [a.elapsed_time(b) for a, b in self.gather_events]
        self.gather_events = []
        return ms

    def bag_digests(self, result):
        """{slide: sha256 of the bag's fp32 bytes} for the bags this rank owns (bit-exactness across rank counts)."""
        out = {}
        if not result["owned"]:
            return out
        bags = result["bags"].cpu() if "bags" in result else None
        offs = self.lay.bag_offsets
        for i, s in enumerate(result["owned"]):
            out[s] = hashlib.sha256(bags[offs[i]:offs[i + 1]].contiguous().numpy().tobytes()).hexdigest()
        return out


def all_digests(local: dict, world_size: int, group=None) -> dict:
    """Union of the per-rank {slide: digest} maps on every rank."""
    if world_size == 1:
        return dict(local)
    parts = [None] * world_size
    dist.all_gather_object(parts, local, group=group)
    merged = {}
    for p in parts:
        merged.update(p)
    return merged
