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


class _FakeStore:
    """The part of engine.ParamStore that GradSync uses: a flat gradient buffer, a touch log, a generation counter."""

    def __init__(self, sizes):
        self.sizes, self.offsets, off = sizes, [], 0
        for n in sizes:
            self.offsets.append(off)
            off += (n + 63) // 64 * 64
        self.total = off
        self.grads = torch.zeros(off)
        self.touch_log = None
        self.generation = 1

    def write(self, i, value):
        if self.touch_log is not None:
            self.touch_log.append((self.offsets[i], self.sizes[i]))
        self.grads[self.offsets[i]:self.offsets[i] + self.sizes[i]] = value


def _run_backward(sync, store, order, values):
    """`order`: one list of parameter indices per tape entry (the engine's backward protocol, engine.UNetEngine.backward)."""
    store.grads.zero_()
    sync.begin(store, len(order))
    try:
        for done, params in enumerate(order, 1):
            for i in params:
                store.write(i, values[i])
            sync.after_entry(done)
    finally:
        sync.finish()


def _sync_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pokemon_sprite_generator_b200.parallel import GradSync
    sizes = [300, 70, 1000, 64, 513, 2000, 5, 900]
    store = _FakeStore(sizes)
    # backward finalises the tail first; parameter 1 is written twice (entries 2 and 6); parameter 0 last
    order = [[7], [6, 1], [5], [4, 3], [], [2, 1], [0]]
    reserved = []
    sync = GradSync(None, bucket_bytes=4 * 1024, reserve_sms=8, reserve_hook=reserved.append, prescaled=False)
    # the engine's hooks around a bucket's all-reduce: before_issue (a stream join) and issue_ctx (a stream made current)
    import contextlib
    hook_log = []

    @contextlib.contextmanager
    def issue_ctx():
        hook_log.append("enter")
        yield
        hook_log.append("exit")

    sync.before_issue = lambda: hook_log.append("before")
    sync.issue_ctx = issue_ctx
    results = []
    for step in range(3):
        values = [float((rank + 1) * (i + 1) + step) for i in range(len(sizes))]
        _run_backward(sync, store, order, values)
        expect = torch.zeros(store.total)
        for i, n in enumerate(sizes):
            expect[store.offsets[i]:store.offsets[i] + n] = sum((r + 1) * (i + 1) + step for r in range(world)) / world
        results.append(torch.allclose(store.grads, expect, atol=1e-6))
    st = dict(sync.stats)
    ok_plan = st["calibrations"] == 1 and st["overlapped_buckets"] > 0 and store.touch_log is None
    # steps 2 and 3 issue most buckets from inside backward; the SM reservation is raised then and always released
    ok_hook = reserved.count(8) >= 2 and reserved[-1] == 0
    # every all-reduce went out inside the context, after the before-hook: (before, enter, exit) per issued bucket
    n_issued = len(hook_log) // 3
    ok_hook = ok_hook and n_issued >= len(sync.bounds) and hook_log == ["before", "enter", "exit"] * n_issued
    # a tape that writes a bucket after it was reduced must be caught, not silently mis-reduced
    caught = False
    try:
        _run_backward(sync, store, [[7], [6, 1], [5], [4, 3], [], [2], [0, 7]], [1.0] * len(sizes))
    except RuntimeError:
        caught = True
    # a different tape length recalibrates instead of reusing the schedule
    _run_backward(sync, store, order[:-1] + [[], [0]], [2.0] * len(sizes))
    ok_recal = sync.stats["calibrations"] == 2 and bool(torch.all(store.grads[:300] == 2.0))
    q.put((rank, all(results), ok_plan, ok_hook, caught, ok_recal))
    dist.destroy_process_group()


def test_two_rank_gloo_gradsync_overlapped_buckets():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(all(r[1:]) for r in results), results


def test_gradsync_plan_head_bucket_and_bounds():
    """GradSync._plan (no process group needed): the leading ranges finalised in the last tape entries form the head bucket,
    the rest is cut into ~bucket_bytes slices that tile the buffer exactly, and each bucket is issued after its last writer."""
    from pokemon_sprite_generator_b200.parallel import GradSync
    sync = GradSync(None, bucket_bytes=4 * 1000)
    sync.flat = torch.zeros(10_000)
    sync.n_entries = 100
    # offset -> (numel, last entry): [0, 1500) written at the very end (conditioning / time MLP), the tail first
    sync.range_last = {0: (1000, 100), 1000: (500, 99), 1500: (2500, 60), 4000: (3000, 40), 7000: (2990, 10)}
    sync._plan()
    assert sync.bounds[0] == (0, 1500)
    assert sync.bounds[-1][1] == 10_000 and all(a[1] == b[0] for a, b in zip(sync.bounds, sync.bounds[1:]))
    assert all(e - s <= 1000 + 1 for s, e in sync.bounds[1:])
    issued_at = {b: at for at, bs in sync.ready_at.items() for b in bs}
    assert 0 not in issued_at                                   # the head bucket goes out after the last entry (finish())
    for b, (s, e) in enumerate(sync.bounds[1:], 1):
        writers = [at for off, (n, at) in sync.range_last.items() if off < e and off + n > s]
        assert issued_at[b] == max(writers)
