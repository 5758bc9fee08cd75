"""Config 3 on real NCCL ranks: a small slide set (one slide large enough to be split across ranks) is sharded over the ranks,
every rank extracts its regions, spanning bags are assembled with the NCCL all-gather, and every bag must equal the bag a
single rank computes on its own, bit for bit (regions are generated from seed = 1000 + global region index).
    torchrun --nproc-per-node 2 tools/check_sharded_nccl.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from tests.common import seeded_modules, seeded_clam
from hipt_abmil_atec23_b200.hipt_4k import HIPT_4K
from hipt_abmil_atec23_b200 import sharding, clam_engine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ["NCCL_DEBUG"] = "WARN"
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
m256, m4k = seeded_modules(0)
hipt = HIPT_4K.from_modules(m256, m4k, dev, dev)
regions_per_slide = [7, 2, 1, 2]                       # 12 regions; slide 0 exceeds the balance tolerance and is split
first = [0, 7, 9, 10]
shards, spanning = sharding.plan_shards(regions_per_slide, world)
assert 0 in spanning, spanning

def region(gidx):
    g = torch.Generator(device=dev).manual_seed(1000 + gidx)
    return torch.randint(0, 256, (1, 3, 4096, 4096), dtype=torch.uint8, device=dev, generator=g)

def extract(items):
    idx = [first[s] + start + i for s, start, n in items for i in range(n)]
    if not idx:
        return torch.empty((0, 192), device=dev)
    return hipt.forward_regions_u8(torch.cat([region(i) for i in idx]))

mine = shards[rank]
feats = extract(mine.items)
bags = sharding.assemble_bags(mine, feats, regions_per_slide, spanning, world)
# single-rank truth for the bags this rank holds
ok = True
for slide, bag in sorted(bags.items()):
    ref = extract([(slide, 0, regions_per_slide[slide])])
    same = torch.equal(bag, ref)
    ok &= same
    print(f"rank {rank} slide {slide}: {tuple(bag.shape)} bit-identical to the 1-rank bag: {same}", flush=True)
clam = [seeded_clam("hipt_smaller", 10 + f).to(dev) for f in range(5)]
b0 = bags[0]
r = clam_engine.forward_bags(clam, b0, torch.tensor([0, b0.shape[0]], dtype=torch.int32), max_bag_len=b0.shape[0])
allr = [torch.empty_like(r["logits"]) for _ in range(world)]
dist.all_gather(allr, r["logits"])
ok &= all(torch.equal(allr[0], a) for a in allr)
t = torch.tensor([int(ok)], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("spanning:", spanning, "ALL OK" if int(t) == 1 else "MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t) == 1 else 1)
