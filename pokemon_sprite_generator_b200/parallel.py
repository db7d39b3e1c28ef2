"""Host-side data-parallel helpers (one process per GPU, torch.distributed for plumbing).

Training shards the batch; the only exchange step is the gradient all-reduce over the flat fp32 gradient buffer
(SURVEY.md 8e).  Sampling shards prompts with no communication.  These helpers are pure host logic so the N>1 path is
covered by world_size-2 gloo tests on CPU (tests/test_parallel.py)."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def world_info(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, end) slice of n items for `rank`; the first n % world ranks get one extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_prompts(text_emb: torch.Tensor, group=None) -> torch.Tensor:
    """This rank's slice of a global prompt batch (DDPM sampling: no communication)."""
    rank, world = world_info(group)
    s, e = shard_range(text_emb.shape[0], rank, world)
    return text_emb[s:e]


def allreduce_mean_(flat: torch.Tensor, group=None, prescaled: bool = False, buckets: int = 1) -> torch.Tensor:
    """In-place mean of a flat gradient buffer across ranks.  `prescaled`: each rank already multiplied its gradient by
    1/world (the loss kernel does), so a plain SUM finishes the mean.  `buckets` > 1 splits the message into equal
    contiguous slices issued back to back (async), tail first -- backward finalises the tail of the buffer first."""
    rank, world = world_info(group)
    if world == 1:
        return flat
    n = flat.numel()
    works = []
    for b in reversed(range(buckets)):
        s, e = shard_range(n, b, buckets)
        if e > s:
            works.append(dist.all_reduce(flat[s:e], op=dist.ReduceOp.SUM, group=group, async_op=True))
    for w in works:
        w.wait()
    if not prescaled:
        flat.div_(world)
    return flat


def seed_for_rank(base_seed: int, rank: int) -> int:
    """Per-rank data seed (1234 + rank)."""
    return base_seed + rank


class GradSync:
    """Bucketed gradient all-reduce overlapped with backward (the reference has no DDP: this is the new
    path's only exchange step).

    The engine's backward tape reports, through `store.touch_log`, which gradient ranges each tape entry writes.  The first
    backward of a given tape shape is a calibration pass (all-reduce after backward) that records the LAST entry touching
    each range and then cuts the flat gradient buffer into contiguous buckets: one head bucket holding the ranges that
    only become final in the last entries of backward (the all-blocks conditioning projections and the time MLP, laid out
    first: ~1 % of the bytes -- the only exposed communication), then ~`bucket_bytes` slices.  From then on a bucket's
    all-reduce is issued (async, on the process group's own stream, ordered after the kernels launched so far) right
    after the entry that finalises it.  A touch of an already-issued bucket raises: the schedule is verified on every
    step, not trusted.

    While a bucket is in flight the persistent GEMM kernels leave `reserve_sms` SMs to the collective's CTAs
    (`reserve_hook(n)`; cap the collective with NCCL_MAX_CTAS = reserve_sms): a persistent grid that cannot co-reside with
    them would otherwise wait a whole wave for the few SMs they hold.  "In flight" is counted in tape entries
    (`window_entries` per bucket): kernels enqueued in the entries right after the issue point are exactly the ones that
    run next to the collective on the device, however far the host runs ahead."""

    def __init__(self, group=None, bucket_bytes: int = 128 << 20, reserve_sms: int = 0,
                 reserve_hook: Optional[Callable[[int], None]] = None, prescaled: bool = True, window_entries: int = 4):
        self.group = group
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.reserve_sms, self.reserve_hook = reserve_sms, reserve_hook
        self.window_entries = window_entries
        self.prescaled = prescaled
        self.plan_key = None
        self.ready_at: Dict[int, List[int]] = {}        # entries done -> buckets to issue
        self.bounds: List[Tuple[int, int]] = []
        self.stats = {"calibrations": 0, "overlapped_buckets": 0, "tail_buckets": 0}
        self._reserved = 0
        self.before_issue: Optional[Callable[[], None]] = None     # called before a bucket goes out (e.g. a stream join)
        # context manager factory entered around the all-reduce call: the engine makes its weight-gradient stream current there, so
        # that the collective is ordered after BOTH streams without stalling the main (dX) stream at every bucket
        self.issue_ctx: Optional[Callable[[], object]] = None

    def _reserve(self, n: int) -> None:
        if self.reserve_hook is not None and n != self._reserved:
            self.reserve_hook(n)
            self._reserved = n

    # ---- per-backward protocol (called by UNetEngine.backward) ---------------------------------------------------
    def begin(self, store, n_entries: int) -> None:
        self.store = store
        self.rank, self.world = world_info(self.group)
        self.flat = store.grads
        key = (id(store), getattr(store, "generation", 0), n_entries, self.flat.numel())
        if key != self.plan_key:
            self.plan_key, self.calibrating = key, True
            self.bounds = [(0, self.flat.numel())]
            self.range_last: Dict[int, Tuple[int, int]] = {}      # offset -> (numel, last entry that touched it)
            self.stats["calibrations"] += 1
        else:
            self.calibrating = False
        self.n_entries = n_entries
        self.issued = [False] * len(self.bounds)
        self.works = []
        self.reserve_until = 0
        store.touch_log = []

    def _plan(self) -> None:
        """Bucket bounds and issue points from the calibration pass's last-touch record."""
        n = self.flat.numel()
        ranges = sorted((off, numel, at) for off, (numel, at) in self.range_last.items())
        tail_entries = max(2, self.n_entries // 50)
        head_end = 0
        for off, numel, at in ranges:                      # the leading run of ranges finalised at the very end
            if at <= self.n_entries - tail_entries:
                break
            head_end = min(n, off + numel)
        bounds = [(0, head_end)] if head_end > 0 else []
        rest = n - head_end
        nb = max(1, -(-rest // self.bucket_elems)) if rest > 0 else 0
        for b in range(nb):
            s, e = shard_range(rest, b, nb)
            bounds.append((head_end + s, head_end + e))
        ready = [0] * len(bounds)
        for off, numel, at in ranges:
            for b, (s, e) in enumerate(bounds):
                if off < e and off + numel > s:
                    ready[b] = max(ready[b], at)
        self.bounds = bounds
        self.ready_at = {}
        for b, at in enumerate(ready):
            # never-touched buckets (padding only) and the last entry's buckets go out at the end
            if 0 < at < self.n_entries:
                self.ready_at.setdefault(at, []).append(b)

    def _buckets_of(self, off: int, numel: int):
        for b, (s, e) in enumerate(self.bounds):
            if off < e and off + max(numel, 1) > s:
                yield b

    def _issue(self, b: int) -> None:
        s, e = self.bounds[b]
        self.issued[b] = True
        if self.world > 1 and e > s:
            if self.before_issue is not None:
                self.before_issue()
            if self.issue_ctx is not None:
                with self.issue_ctx():
                    self.works.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            else:
                self.works.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def after_entry(self, done: int) -> None:
        log = self.store.touch_log
        if self.calibrating:
            for off, numel in log:
                prev = self.range_last.get(off)
                self.range_last[off] = (max(numel, prev[0]) if prev else numel, done)
        else:
            for off, numel in log:
                for b in self._buckets_of(off, numel):
                    if self.issued[b]:
                        raise RuntimeError(f"GradSync: gradient range [{off}, {off + numel}) written after its bucket {b} was reduced")
        log.clear()
        if not self.calibrating:
            for b in self.ready_at.get(done, ()):
                self._issue(b)
                self.stats["overlapped_buckets"] += 1
                self.reserve_until = max(self.reserve_until, done) + self.window_entries
            if self.world > 1:
                self._reserve(self.reserve_sms if done < self.reserve_until else 0)

    def finish(self) -> None:
        self.store.touch_log = None
        try:
            self._reserve(0)
            for b in reversed(range(len(self.bounds))):       # whatever is left: tail first
                if not self.issued[b]:
                    self._issue(b)
                    self.stats["tail_buckets"] += 1
            for w in self.works:
                w.wait()
            self.works = []
            if not self.prescaled and self.world > 1:
                self.flat.div_(self.world)
            if self.calibrating:
                self._plan()
        finally:
            self._reserve(0)
