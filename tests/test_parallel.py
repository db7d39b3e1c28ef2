"""world_size-2 gloo tests (CPU) of the data-parallel host logic: prompt sharding and the flat-gradient mean."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_range_partitions_exactly():
    from pokemon_sprite_generator_b200.parallel import shard_range
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pokemon_sprite_generator_b200.parallel import allreduce_mean_, shard_prompts, world_info
    assert world_info() == (rank, world)
    g = torch.Generator().manual_seed(5)
    full = torch.randn(2, 1003, generator=g)            # both ranks' "gradients", known everywhere
    mine = full[rank].clone()
    allreduce_mean_(mine, buckets=3)
    ok_mean = torch.allclose(mine, full.mean(0), atol=1e-7)
    pre = full[rank].clone() / world
    allreduce_mean_(pre, prescaled=True)
    ok_pre = torch.allclose(pre, full.mean(0), atol=1e-7)
    prompts = torch.arange(5 * 3, dtype=torch.float32).view(5, 3)
    part = shard_prompts(prompts)
    gathered = [None] * world
    dist.all_gather_object(gathered, part)
    ok_shard = torch.equal(torch.cat(gathered), prompts)
    q.put((rank, ok_mean, ok_pre, ok_shard))
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(all(r[1:]) for r in results), results
