"""Batch-sharded data parallelism: host-side logic (one process per GPU, torch.distributed).

The SR model is sample-independent (LayerNorm is per pixel, the loss is a mean), so each rank
runs the same step on its contiguous slice of the global batch with the loss gradient scaled by
1/world, and the per-step exchange is ONE sum all-reduce of the flat fp32 gradient buffer
(SURVEY section 8e).  The buffer is cut into buckets that are reduced as soon as the backward pass
has written them (reverse layer order), on NCCL's stream, overlapping the rest of backward.

Nothing here touches CUDA directly; the same code runs under gloo on CPU tensors (tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a global batch of n samples owned by `rank` (equal shards)."""
    if n % world != 0:
        raise ValueError(f"global batch {n} is not divisible by world size {world}: equal shards are "
                         "required for mean-of-shard-means == global mean")
    per = n // world
    return rank * per, (rank + 1) * per


def plan_buckets(total: int, writes: Sequence[Sequence[Tuple[int, int]]], bucket_elems: int) -> List[dict]:
    """Cut [0, total) into buckets of about `bucket_elems` elements, filled from the END of the buffer
    (backward writes the last layers first), never splitting a written range, and find for each
    bucket the index of the last backward step that writes into it.

    writes[i] = [(offset, count), ...] ranges of the flat gradient buffer written by backward step i.
    Returns [{"lo", "hi", "ready_after"}] ordered by readiness.
    """
    ranges = sorted({(o, c) for step in writes for (o, c) in step})
    last_writer = {}
    for i, step in enumerate(writes):
        for rng in step:
            last_writer[rng] = max(last_writer.get(rng, -1), i)
    buckets = []
    hi = total
    cur_lo, cur_ready = total, -1
    for (o, c) in reversed(ranges):
        cur_lo = o
        cur_ready = max(cur_ready, last_writer[(o, c)])
        if hi - cur_lo >= bucket_elems:
            buckets.append({"lo": cur_lo, "hi": hi, "ready_after": cur_ready})
            hi, cur_ready = cur_lo, -1
    if hi > 0:
        buckets.append({"lo": 0, "hi": hi, "ready_after": max(cur_ready, 0) if ranges else 0})
    buckets.sort(key=lambda b: b["ready_after"])
    return buckets


def allreduce_buckets(dist, flat, buckets, group=None, async_op=False):
    """Sum-all-reduce every bucket of `flat` (in readiness order).  Returns the work handles."""
    works = []
    for b in buckets:
        w = dist.all_reduce(flat[b["lo"]:b["hi"]], op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works
