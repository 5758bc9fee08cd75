"""Region sharding across ranks and per-slide bag assembly.

Every 4096 x 4096 region is an independent work item through ViT-256 -> ViT-4K (hipt_4k.py:48-76 has no cross-region
state); the only cross-region step is collecting one slide's region embeddings into its bag before CLAM_SB.  The
reference does this through the file system (features appended to an .h5 per region, extract_features_fp.py:159-171,
re-read per slide by the MIL dataset); here regions are sharded over one process per GPU:

  * slide-aligned assignment (greedy longest-processing-time over slides by region count) keeps every bag on one rank,
    so the data path has NO collective;
  * when there are fewer slides than ranks, or one slide is too large for balance, that slide's regions are split
    contiguously across ranks and its bag is assembled with ONE all-gather of [n_r, 192] fp32 rows
    (torch.distributed: NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch
import torch.distributed as dist


@dataclass
class Shard:
    """Work of one rank: (slide, first region, n regions) triples in processing order."""
    rank: int
    items: List[Tuple[int, int, int]] = field(default_factory=list)

    @property
    def n_regions(self):
        return sum(n for _, _, n in self.items)


def plan_shards(regions_per_slide: List[int], world_size: int, imbalance_tol: float = 0.10) -> Tuple[List[Shard], Dict[int, List[int]]]:
    """Assign slides (with their region counts) to ranks.

    Returns (shards, spanning): spanning[slide] = ranks holding a piece of that slide (only for split slides).
    Slides are placed whole by LPT; a slide is split contiguously over all ranks only when keeping it whole would leave
    the most loaded rank more than `imbalance_tol` above the mean load."""
    shards = [Shard(r) for r in range(world_size)]
    spanning: Dict[int, List[int]] = {}
    total = sum(regions_per_slide)
    if total == 0 or world_size == 1:
        for s, n in enumerate(regions_per_slide):
            if n:
                shards[0].items.append((s, 0, n))
        return shards, spanning
    mean = total / world_size
    order = sorted(range(len(regions_per_slide)), key=lambda s: -regions_per_slide[s])
    load = [0] * world_size
    split = [s for s in order if regions_per_slide[s] > mean * (1.0 + imbalance_tol)]
    for s in split:                                   # too big for one rank: contiguous pieces on every rank
        n = regions_per_slide[s]
        base, extra = divmod(n, world_size)
        start = 0
        ranks = []
        for r in range(world_size):
            cnt = base + (1 if r < extra else 0)
            if cnt:
                shards[r].items.append((s, start, cnt))
                load[r] += cnt
                ranks.append(r)
                start += cnt
        spanning[s] = ranks
    for s in order:
        n = regions_per_slide[s]
        if s in spanning or n == 0:
            continue
        r = min(range(world_size), key=lambda i: (load[i], i))
        shards[r].items.append((s, 0, n))
        load[r] += n
    for sh in shards:
        sh.items.sort()
    return shards, spanning


def gather_bag(local_rows: torch.Tensor, counts: List[int], group=None) -> torch.Tensor:
    """All-gather the pieces of one spanning slide's bag: rank r contributes local_rows [counts[r], F]; every rank gets
    the [sum(counts), F] bag in region order.  Uneven pieces are padded to max(counts) for a single all_gather."""
    world = dist.get_world_size(group)
    assert len(counts) == world
    F = local_rows.shape[1]
    mx = max(counts)
    pad = torch.zeros((mx, F), dtype=local_rows.dtype, device=local_rows.device)
    pad[:local_rows.shape[0]] = local_rows
    out = torch.empty((world * mx, F), dtype=local_rows.dtype, device=local_rows.device)
    if local_rows.is_cuda:
        dist.all_gather_into_tensor(out, pad, group=group)          # NCCL, one contiguous buffer
    else:
        dist.all_gather(list(out.view(world, mx, F).unbind(0)), pad, group=group)
    pieces = [out.view(world, mx, F)[r, :counts[r]] for r in range(world)]
    return torch.cat(pieces, dim=0)


def assemble_bags(shard: Shard, local_feats: torch.Tensor, regions_per_slide: List[int], spanning: Dict[int, List[int]],
                  world_size: int, group=None) -> Dict[int, torch.Tensor]:
    """local_feats: this rank's region embeddings [shard.n_regions, F] in shard.items order.  Returns {slide: bag} for the
    slides this rank owns whole plus every spanning slide (all ranks take part in those collectives, in slide order)."""
    bags: Dict[int, torch.Tensor] = {}
    off = 0
    mine: Dict[int, torch.Tensor] = {}
    for slide, start, n in shard.items:
        mine[slide] = local_feats[off:off + n]
        off += n
    for slide in sorted(spanning):
        n = regions_per_slide[slide]
        base, extra = divmod(n, world_size)
        counts = [base + (1 if r < extra else 0) for r in range(world_size)]
        rows = mine.pop(slide, local_feats[:0])
        bags[slide] = gather_bag(rows, counts, group)
    bags.update(mine)
    return bags


# ----------------------------------------------------------------------------------------------------------------------
# Whole-slide-set layout (BASELINE.json configs 3 / 5): who extracts which region, where each bag is pooled, and the ONE
# collective per pass that moves the rows of slides cut by a rank boundary.
# ----------------------------------------------------------------------------------------------------------------------
def plan_contiguous(regions_per_slide: List[int], world_size: int) -> Tuple[List[Shard], Dict[int, List[int]]]:
    """Cut the global region list (slide order) into world_size contiguous pieces whose sizes differ by at most one.
    Loads are exactly balanced (strong scaling of a fixed slide set); at most world_size - 1 slides are cut by a rank
    boundary, and only their rows ever cross NVLink.  Same return convention as plan_shards."""
    shards = [Shard(r) for r in range(world_size)]
    spanning: Dict[int, List[int]] = {}
    total = sum(regions_per_slide)
    base, extra = divmod(total, world_size)
    bounds = [0]
    for r in range(world_size):
        bounds.append(bounds[-1] + base + (1 if r < extra else 0))
    g0 = 0
    for s, n in enumerate(regions_per_slide):
        if n == 0:
            continue
        ranks = []
        for r in range(world_size):
            lo, hi = max(g0, bounds[r]), min(g0 + n, bounds[r + 1])
            if hi > lo:
                shards[r].items.append((s, lo - g0, hi - lo))
                ranks.append(r)
        if len(ranks) > 1:
            spanning[s] = ranks
        g0 += n
    return shards, spanning


@dataclass
class RankLayout:
    """Static layout of one rank's pass over a slide set (host-side lists; SlideSetLayout builds one per rank).

    Local feature row i is global region `regions[i]` (ascending).  `send_index` lists the local rows that belong to slides
    cut by a rank boundary (they are what this rank contributes to the collective, padded to layout.pad_rows);
    `bag_index` says where every row of every owned bag comes from: (0, local row) or (1, row of the gathered buffer)."""
    rank: int
    regions: List[int] = field(default_factory=list)
    send_index: List[int] = field(default_factory=list)
    owned: List[int] = field(default_factory=list)              # slides pooled on this rank: whole slides, then homed cut slides
    bag_offsets: List[int] = field(default_factory=lambda: [0])  # row offsets of the owned bags in the CLAM input
    bag_index: List[Tuple[int, int]] = field(default_factory=list)

    def pool_rows(self, local_capacity: int) -> List[int]:
        """bag_index resolved against ONE allocation [local_capacity + world * pad_rows, F] whose head is the local feature
        buffer and whose tail is the gathered buffer."""
        return [row if src == 0 else local_capacity + row for src, row in self.bag_index]


class SlideSetLayout:
    """Everything that is known before the first kernel runs: per-rank region lists, the padded all-gather geometry
    (pad_rows rows per rank), the home rank of every cut slide (the rank holding most of its rows) and the gather indices
    that turn local + gathered rows into the bags each rank pools."""

    def __init__(self, regions_per_slide: List[int], world_size: int, policy: str = "contiguous", imbalance_tol: float = 0.10):
        self.regions_per_slide = list(regions_per_slide)
        self.world_size = world_size
        self.policy = policy
        if policy == "contiguous":
            self.shards, self.spanning = plan_contiguous(self.regions_per_slide, world_size)
        elif policy == "lpt":
            self.shards, self.spanning = plan_shards(self.regions_per_slide, world_size, imbalance_tol)
        else:
            raise ValueError(f"unknown sharding policy '{policy}' (contiguous | lpt)")
        self.first_region = [0]
        for n in self.regions_per_slide:
            self.first_region.append(self.first_region[-1] + n)
        self.ranks = [RankLayout(r) for r in range(world_size)]
        piece: Dict[Tuple[int, int], Tuple[int, int]] = {}       # (slide, rank) -> (offset in the rank's send rows, n)
        for sh, lay in zip(self.shards, self.ranks):
            for s, st, n in sorted(sh.items):                      # ascending global region index
                row0 = len(lay.regions)
                lay.regions += [self.first_region[s] + st + i for i in range(n)]
                if s in self.spanning:
                    piece[(s, sh.rank)] = (len(lay.send_index), n)
                    lay.send_index += list(range(row0, row0 + n))
                else:
                    lay.owned.append(s)
                    lay.bag_offsets.append(lay.bag_offsets[-1] + n)
                    lay.bag_index += [(0, row0 + i) for i in range(n)]
        self.pad_rows = max([len(lay.send_index) for lay in self.ranks] + [0])
        self.home: Dict[int, int] = {}
        for s in sorted(self.spanning):
            ranks = self.spanning[s]
            home = max(ranks, key=lambda r: (piece[(s, r)][1], -r))
            self.home[s] = home
            lay = self.ranks[home]
            for r in ranks:                                        # ranks ascend with the region index inside a cut slide
                off, n = piece[(s, r)]
                lay.bag_index += [(1, r * self.pad_rows + off + i) for i in range(n)]
            lay.owned.append(s)
            lay.bag_offsets.append(lay.bag_offsets[-1] + self.regions_per_slide[s])

    @property
    def needs_collective(self):
        return self.world_size > 1 and self.pad_rows > 0

    def load(self):
        return [len(lay.regions) for lay in self.ranks]


def gather_span_rows(send: torch.Tensor, gathered: torch.Tensor, group=None) -> torch.Tensor:
    """The one collective of a pass: every rank contributes `send` [pad_rows, F] (its rows of the cut slides, padded) and
    receives [world * pad_rows, F] in a pre-allocated destination.  NCCL all-gather on the GPU box; gloo in the CPU tests."""
    if send.is_cuda:
        dist.all_gather_into_tensor(gathered, send, group=group)
    else:
        world = dist.get_world_size(group)
        dist.all_gather(list(gathered.view(world, send.shape[0], -1).unbind(0)), send, group=group)
    return gathered


def assemble_owned_bags(layout: SlideSetLayout, lay: RankLayout, pool: torch.Tensor, local_capacity: int, send: torch.Tensor,
                        send_index: torch.Tensor, pool_rows: torch.Tensor, clam_in: torch.Tensor, group=None,
                        collective=True) -> torch.Tensor:
    """Turn this rank's feature rows into the contiguous CLAM input of the bags it owns (lay.owned, lay.bag_offsets).
    `pool` [local_capacity + world * pad_rows, F] is ONE allocation: head = local feature buffer (ViT-4K writes it), tail =
    destination of the collective.  Three steps, all on caller-owned buffers: index_select of the cut slides' local rows
    into `send`, the all-gather into pool's tail, index_select of every owned row (pool_rows) into `clam_in`.
    `collective=False` skips the all-gather (tail already filled: single-process simulation of all ranks)."""
    if layout.needs_collective and collective:
        if lay.send_index:
            torch.index_select(pool[:local_capacity], 0, send_index, out=send[:len(lay.send_index)])
        gather_span_rows(send, pool[local_capacity:], group)
    n_owned = lay.bag_offsets[-1]
    if n_owned:
        torch.index_select(pool, 0, pool_rows, out=clam_in[:n_owned])
    return clam_in[:n_owned]
