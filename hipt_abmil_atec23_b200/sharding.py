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
