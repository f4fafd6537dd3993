"""Page-sharded multi-GPU execution: pages are independent units, so ranks take
contiguous page ranges, run the whole page -> text path on their own GPU, and the per-page results
(Python objects: masks, Line / OCRLine lists) are gathered on the HOST by rank 0.  There is no
collective on the data path and nothing crosses NVLink; torch.distributed is only the rendezvous
and the object gather (NCCL process groups delegate object gathers to the CPU as well)."""
from __future__ import annotations

import os
from typing import Callable, Sequence

import torch.distributed as dist


def shard_range(n_items: int, world: int, rank: int) -> range:
    """Contiguous shard [rank*n/world, (rank+1)*n/world): sizes differ by at most one."""
    return range(rank * n_items // world, (rank + 1) * n_items // world)


def rank_world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def run_sharded(n_pages: int, load_page: Callable[[int], object], process: Callable[[Sequence[object]], list],
                chunk: int = 16, group=None):
    """Every rank processes its page range in chunks of `chunk` pages through `process` (e.g.
    OCRPipeline.run_ocr_batch) and rank 0 returns the results in page order (others return None)."""
    rank, world = rank_world()
    mine = shard_range(n_pages, world, rank)
    local = []
    for c0 in range(mine.start, mine.stop, chunk):
        idx = range(c0, min(c0 + chunk, mine.stop))
        res = process([load_page(i) for i in idx])
        assert len(res) == len(idx)
        local.extend(zip(idx, res))
    if world == 1:
        return [r for _, r in local]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0, group=group)
    if rank != 0:
        return None
    flat = sorted((p for part in gathered for p in part), key=lambda t: t[0])
    assert [i for i, _ in flat] == list(range(n_pages))
    return [r for _, r in flat]
