"""Host-side data-parallel helpers (one process per GPU, torch.distributed for plumbing).

Training shards the batch; the only exchange step is the gradient all-reduce over the flat fp32 gradient buffer
(SURVEY.md 8e).  Sampling shards prompts with no communication.  These helpers are pure host logic so the N>1 path is
covered by world_size-2 gloo tests on CPU (tests/test_parallel.py)."""
from __future__ import annotations

from typing import Optional, Tuple

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
