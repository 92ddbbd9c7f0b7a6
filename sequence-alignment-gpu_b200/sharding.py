"""Multi-GPU sharding of a batch of independent pairs (BASELINE config 4).

One process per GPU.  Pairs are independent, so the batch is cut into `world` contiguous,
cell-balanced ranges (sa_partition_batch) and every rank aligns its own range on its own GPU:
there is NO data-path collective.  torch.distributed is only used to (optionally) gather the small
per-pair results for a caller that wants them on one rank.
"""
from __future__ import annotations

import numpy as np


def shard_csr(partition, text, text_off, pattern, pattern_off, rank: int):
    """Slice a CSR batch to the pair range of `rank`; offsets are rebased to start at 0.
    `partition` is the world+1 array returned by partition_batch()."""
    first, last = int(partition[rank]), int(partition[rank + 1])
    toff = np.asarray(text_off[first:last + 1], dtype=np.int64)
    poff = np.asarray(pattern_off[first:last + 1], dtype=np.int64)
    t = text[toff[0]:toff[-1]] if last > first else text[:0]
    p = pattern[poff[0]:poff[-1]] if last > first else pattern[:0]
    return dict(first=first, count=last - first, text=np.ascontiguousarray(t), text_off=toff - toff[0],
                pattern=np.ascontiguousarray(p), pattern_off=poff - poff[0])


def align_batch_sharded(align_fn, partition, text, text_off, pattern, pattern_off, rank: int, world: int,
                        gather: bool = False):
    """Runs `align_fn(text, text_off, pattern, pattern_off) -> list of per-pair results` on this rank's
    shard.  With gather=True the per-pair results of all ranks are all-gathered (torch.distributed must
    be initialised) and returned in the original pair order on every rank."""
    sh = shard_csr(partition, text, text_off, pattern, pattern_off, rank)
    local = align_fn(sh["text"], sh["text_off"], sh["pattern"], sh["pattern_off"]) if sh["count"] else []
    if not gather or world == 1:
        return sh["first"], local
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, (sh["first"], local))
    parts.sort(key=lambda x: x[0])
    merged = []
    for _, res in parts:
        merged.extend(res)
    return 0, merged
