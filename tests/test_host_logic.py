"""Host-side logic that needs no GPU: module trees / state_dict contract, loaders, crop, weight folding, sharding
(including world_size-2 gloo runs of the bag assembly)."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_import_paths_of_the_reference_resolve():
    from HIPT_4K.hipt_4k import HIPT_4K  # noqa: F401
    from HIPT_4K.hipt_model_utils import eval_transforms, get_vit256, get_vit4k, roll_batch2img, tensorbatch2im  # noqa: F401
    import HIPT_4K.vision_transformer as vits
    import HIPT_4K.vision_transformer4k as vits4k
    from models.model_clam import CLAM_MB, CLAM_SB  # noqa: F401
    from models.model_mil import MIL_fc, MIL_fc_mc  # noqa: F401
    from utils.utils import initialize_weights  # noqa: F401
    assert callable(vits.vit_small) and callable(vits4k.vit4k_xs)


def test_state_dict_contract():
    from tests.common import seeded_clam, seeded_modules
    m256, m4k = seeded_modules(0)
    sd = m256.state_dict()
    assert len(sd) == 150 and sum(p.numel() for p in m256.parameters()) == 21_665_664
    assert sd["pos_embed"].shape == (1, 197, 384) and sd["patch_embed.proj.weight"].shape == (384, 3, 16, 16)
    assert sd["blocks.11.attn.qkv.weight"].shape == (1152, 384) and sd["blocks.0.mlp.fc2.weight"].shape == (384, 1536)
    sd4 = m4k.state_dict()
    assert len(sd4) == 78 and sum(p.numel() for p in m4k.parameters()) == 2_781_504
    assert sd4["phi.0.weight"].shape == (192, 384) and sd4["blocks.5.attn.qkv.weight"].shape == (576, 192)
    c = seeded_clam("hipt_smaller", 2, 0.0)
    keys = list(c.state_dict())
    assert len(keys) == 14 and "attention_net.2.attention_a.0.weight" in keys
    assert sum(p.numel() for p in c.parameters()) == 3471
    c2 = seeded_clam("hipt_smaller", 2, 0.25)
    assert "attention_net.3.attention_c.bias" in c2.state_dict()          # gate index shifts with dropout
    with pytest.raises(KeyError):
        seeded_clam("no_such_size")


def test_eval_transforms_and_crop():
    from hipt_abmil_atec23_b200.hipt_4k import HIPT_4K, center_crop_offsets
    from hipt_abmil_atec23_b200.hipt_model_utils import eval_transforms
    from oracle import hipt_oracle as O
    import numpy as np
    arr = np.random.RandomState(0).randint(0, 256, (40, 50, 3), dtype=np.uint8)
    t = eval_transforms()(arr)
    ref = O.eval_transforms_u8(torch.from_numpy(arr).permute(2, 0, 1))
    assert torch.allclose(t, ref, atol=1e-7) and t.shape == (3, 40, 50)
    x = torch.arange(2 * 3 * 549 * 779, dtype=torch.float32).reshape(2, 3, 549, 779)[:1]
    img, w, h = HIPT_4K.prepare_img_tensor(None, x)
    ref_img, rw, rh = O.prepare_img_tensor(x)
    assert (w, h) == (rw, rh) == (2, 3) and torch.equal(img, ref_img)
    try:
        from torchvision import transforms
        assert torch.equal(img, transforms.CenterCrop((512, 768))(x))
    except ImportError:
        pass
    assert center_crop_offsets(549, 512) == 18 and center_crop_offsets(4096, 4096) == 0


def test_pos_table_matches_oracle_interpolation():
    from hipt_abmil_atec23_b200.vision_transformer import interpolate_pos_table
    from oracle import hipt_oracle as O
    pe = torch.randn(1, 197, 384, generator=torch.Generator().manual_seed(0))
    for (w, h) in ((16, 16), (2, 3), (14, 14), (1, 1), (5, 16)):
        got = interpolate_pos_table(pe, w * h, w, h)
        ref = O.interpolate_pos_encoding(pe, w * h, w, h)[0]
        assert torch.equal(got, ref), (w, h)


def test_folded_patch_embed_weights_reproduce_normalised_input():
    """W'/b' folding of ToTensor+Normalize (engine.embed_weights) in fp64: conv(norm(p)) == conv'(p)."""
    g = torch.Generator().manual_seed(0)
    W = torch.randn(8, 3, 16, 16, generator=g, dtype=torch.float64)
    b = torch.randn(8, generator=g, dtype=torch.float64)
    p = torch.randint(0, 256, (1, 3, 16, 16), generator=g).double()
    for mean, std in (((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)), ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))):
        m = torch.tensor(mean, dtype=torch.float64).view(1, 3, 1, 1)
        s = torch.tensor(std, dtype=torch.float64).view(1, 3, 1, 1)
        ref = ((p / 255 - m) / s * W).sum(dim=(1, 2, 3)) + b
        Wf = W / (255.0 * s)
        bf = b - (W * (m / s)).sum(dim=(1, 2, 3))
        assert torch.allclose((p * Wf).sum(dim=(1, 2, 3)) + bf, ref, atol=1e-9)


def test_shard_planning():
    from hipt_abmil_atec23_b200.sharding import plan_shards
    counts = [50] * 40
    for world in (1, 2, 4, 8):
        shards, spanning = plan_shards(counts, world)
        assert not spanning
        assert sum(s.n_regions for s in shards) == 2000
        assert max(s.n_regions for s in shards) - min(s.n_regions for s in shards) <= 50
        seen = sorted(sl for s in shards for sl, _, _ in s.items)
        assert seen == list(range(40))
    # ragged: one huge slide must be split, the rest stay whole
    counts = [300, 10, 20, 30, 15, 25]
    shards, spanning = plan_shards(counts, 4)
    assert list(spanning) == [0] and spanning[0] == [0, 1, 2, 3]
    pieces = sorted((st, n) for s in shards for sl, st, n in s.items if sl == 0)
    assert pieces == [(0, 75), (75, 75), (150, 75), (225, 75)]
    assert sum(s.n_regions for s in shards) == sum(counts)
    # fewer slides than ranks
    shards, spanning = plan_shards([64], 8)
    assert spanning == {0: list(range(8))} and all(s.n_regions == 8 for s in shards)
    shards, spanning = plan_shards([], 2)
    assert all(s.n_regions == 0 for s in shards)


def _gloo_worker(rank, world, port, counts, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from hipt_abmil_atec23_b200.sharding import assemble_bags, plan_shards
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shards, spanning = plan_shards(counts, world)
    sh = shards[rank]
    # feature of region r of slide s is the constant 1000*s + r in every column
    rows = [torch.full((1, 192), 1000.0 * s + (st + i)) for s, st, n in sh.items for i in range(n)]
    local = torch.cat(rows) if rows else torch.zeros(0, 192)
    bags = assemble_bags(sh, local, counts, spanning, world)
    ok = True
    for s, bag in bags.items():
        want = torch.arange(counts[s], dtype=torch.float32) + 1000.0 * s
        ok &= bag.shape == (counts[s], 192) and torch.equal(bag[:, 0], want) and torch.equal(bag[:, 191], want)
    q.put((rank, ok, sorted(bags)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("counts", [[301, 10, 20, 30], [7], [12, 12, 12, 12]])
def test_bag_assembly_world_size_2_gloo(counts):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + sum(counts)) % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    owned = [set(b) for _, _, b in sorted(res)]
    assert owned[0] | owned[1] == {s for s, n in enumerate(counts) if n}    # every slide's bag exists somewhere


# --------------------------------------------------------------------------------------------- slide-set layout (config 3/5)
def _ragged_counts(total=2000, lo=10, hi=300, seed=6):
    g = torch.Generator().manual_seed(seed)
    out = []
    while sum(out) < total:
        out.append(min(int(torch.randint(lo, hi + 1, (1,), generator=g)), total - sum(out)))
    return out


@pytest.mark.parametrize("policy", ["contiguous", "lpt"])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_slide_set_layout_simulated_ranks(policy, world):
    """Every rank's layout, run in one process: feature of global region g is the constant g; after the (simulated)
    collective every slide's bag must be arange(first, first + n) on exactly one rank."""
    from hipt_abmil_atec23_b200.sharding import SlideSetLayout, assemble_owned_bags
    for counts in (_ragged_counts(), [50] * 40, [7], [301, 10, 20, 30], [0, 5, 0, 9]):
        L = SlideSetLayout(counts, world, policy)
        total = sum(counts)
        assert sorted(g for lay in L.ranks for g in lay.regions) == list(range(total))
        if policy == "contiguous":
            assert max(L.load()) - min(L.load()) <= 1
            assert len(L.spanning) <= max(world - 1, 0)
        F = 4
        caps = [max(len(lay.regions), 1) for lay in L.ranks]
        sends = []
        for lay in L.ranks:                                       # what each rank would contribute to the all-gather
            send = torch.zeros(L.pad_rows, F)
            if lay.send_index:
                send[:len(lay.send_index)] = torch.tensor([lay.regions[i] for i in lay.send_index], dtype=torch.float32)[:, None]
            sends.append(send)
        gathered = torch.cat(sends) if L.needs_collective else torch.zeros(0, F)
        owners = {}
        for lay, cap in zip(L.ranks, caps):
            pool = torch.full((cap + gathered.shape[0], F), -1.0)
            if lay.regions:
                pool[:len(lay.regions)] = torch.tensor(lay.regions, dtype=torch.float32)[:, None]
            pool[cap:] = gathered
            clam_in = torch.full((max(lay.bag_offsets[-1], 1), F), -1.0)
            bags = assemble_owned_bags(L, lay, pool, cap, None, None, torch.tensor(lay.pool_rows(cap), dtype=torch.int64),
                                       clam_in, collective=False)
            assert bags.shape[0] == lay.bag_offsets[-1]
            for i, s in enumerate(lay.owned):
                bag = bags[lay.bag_offsets[i]:lay.bag_offsets[i + 1]]
                want = torch.arange(L.first_region[s], L.first_region[s + 1], dtype=torch.float32)
                assert torch.equal(bag[:, 0], want) and torch.equal(bag[:, 3], want), (policy, world, s)
                assert s not in owners
                owners[s] = lay.rank
        assert sorted(owners) == [s for s, n in enumerate(counts) if n]
        for s, home in L.home.items():
            assert owners[s] == home and home in L.spanning[s]


def _gloo_slideset_worker(rank, world, port, counts, policy, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from hipt_abmil_atec23_b200.sharding import SlideSetLayout, assemble_owned_bags
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = SlideSetLayout(counts, world, policy)
    lay = L.ranks[rank]
    cap = max(len(lay.regions), 1)
    pool = torch.zeros(cap + (world * L.pad_rows if L.needs_collective else 0), 192)
    if lay.regions:
        pool[:len(lay.regions)] = torch.tensor(lay.regions, dtype=torch.float32)[:, None]
    send = torch.zeros(L.pad_rows, 192)
    clam_in = torch.empty(max(lay.bag_offsets[-1], 1), 192)
    bags = assemble_owned_bags(L, lay, pool, cap, send, torch.tensor(lay.send_index, dtype=torch.int64),
                               torch.tensor(lay.pool_rows(cap), dtype=torch.int64), clam_in)
    ok = True
    for i, s in enumerate(lay.owned):
        want = torch.arange(L.first_region[s], L.first_region[s + 1], dtype=torch.float32)
        bag = bags[lay.bag_offsets[i]:lay.bag_offsets[i + 1]]
        ok &= bag.shape[0] == counts[s] and torch.equal(bag[:, 0], want) and torch.equal(bag[:, 191], want)
    q.put((rank, ok, list(lay.owned), bool(L.needs_collective)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("counts,policy", [(None, "contiguous"), ([301, 10, 20, 30], "lpt"), ([7], "contiguous"),
                                           ([12, 12, 12, 12], "contiguous")])
def test_slide_set_assembly_world_size_2_gloo(counts, policy):
    counts = _ragged_counts(200, 5, 40) if counts is None else counts
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + sum(counts) + len(policy)) % 2000
    procs = [ctx.Process(target=_gloo_slideset_worker, args=(r, 2, port, counts, policy, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res)
    owned = [set(o) for _, _, o, _ in sorted(res)]
    assert not (owned[0] & owned[1])                                        # a bag is pooled on exactly one rank
    assert owned[0] | owned[1] == {s for s, n in enumerate(counts) if n}
