"""Host-side multi-GPU logic on CPU: gloo, world_size 2 (scene sharding, single flat gradient all-reduce)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from seggroup_b200 import engine
    torch.manual_seed(rank)
    shapes = [(64, 18), (64,), (192, 192), (40, 128), (40,)]
    grads = [torch.full(s, float(rank + 1)) + torch.arange(int(torch.tensor(s).prod())).view(s) * 1e-3 for s in shapes]
    ref = [(g - float(rank + 1)) + 1.5 for g in grads]           # mean of (1, 2) = 1.5 plus the shared ramp
    extra = torch.tensor([1.0 + rank, 10.0 * (rank + 1)])
    got_extra = engine.allreduce_flat(grads, average=True, extra=extra)
    ok = all(torch.allclose(g, r) for g, r in zip(grads, ref)) and torch.allclose(got_extra, torch.tensor([3.0, 30.0]))
    mine = engine.shard_scenes(1201, rank, world)
    counts = torch.tensor([len(mine)])
    dist.all_reduce(counts)
    ok = ok and int(counts) == 1201 and mine[0] == rank and all(b - a == world for a, b in zip(mine, mine[1:]))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_allreduce_and_sharding_gloo_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0] and ret[1]


def test_shard_scenes_partition():
    from seggroup_b200 import engine
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in engine.shard_scenes(1201, r, world))
        assert seen == list(range(1201))


def test_shard_scenes_padded_mode_gives_equal_counts():
    """pad=True repeats scenes from the head of the list like the reference's DistributedSampler (train.py:102), so a loop with a
    per-step collective takes the same number of steps on every rank."""
    from seggroup_b200 import engine
    for world in (2, 4, 8):
        shards = [engine.shard_scenes(1201, r, world, pad=True) for r in range(world)]
        per = -(-1201 // world)
        assert all(len(s) == per for s in shards)
        seen = [i for s in shards for i in s]
        assert set(seen) == set(range(1201)) and len(seen) == per * world
    extra = torch.tensor([1, 2], dtype=torch.int64)
    out = engine.allreduce_flat([torch.zeros(3)], average=True, extra=extra)      # world size 1: a fresh fp32 tensor
    assert out.dtype == torch.float32 and out.data_ptr() != extra.data_ptr()
