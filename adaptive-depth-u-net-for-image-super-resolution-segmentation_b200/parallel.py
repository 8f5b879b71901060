"""Batch-sharded data parallelism: host-side logic (one process per GPU, torch.distributed).

The SR model is sample-independent (LayerNorm is per pixel, the loss is a mean), so each rank
runs the same step on its contiguous slice of the global batch with the loss gradient scaled by
1/world, and the per-step exchange is a sum over ranks of the flat fp32 gradient buffer
(SURVEY section 8e).  The buffer is cut into buckets that are exchanged as soon as the backward pass
has written them (reverse layer order), on NCCL's stream, overlapping the rest of backward.

Two exchange schemes share the bucket plan:
  * all-reduce + replicated Adam (every rank updates every parameter), and
  * sharded optimizer: reduce-scatter of each bucket -> Adam on the owned shard only -> all-gather of
    the compute-dtype shadow.  For a depth-5 net (138 M parameters) this halves the gradient traffic,
    moves a quarter of it (bf16) after the update and cuts the Adam pass by the world size.

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


def complement_ranges(done: Sequence[Tuple[int, int]], total: int) -> List[Tuple[int, int]]:
    """[lo, hi) ranges of [0, total) NOT covered by `done` (half-open ranges, any order, may touch or overlap)."""
    rest, pos = [], 0
    for lo, hi in sorted(done):
        if lo > pos:
            rest.append((pos, min(lo, total)))
        pos = max(pos, hi)
        if pos >= total:
            break
    if pos < total:
        rest.append((pos, total))
    return [(lo, hi) for lo, hi in rest if hi > lo]


def plan_buckets(total: int, writes: Sequence[Sequence[Tuple[int, int]]], bucket_elems: int, align: int = 1,
                 head_elems: int = 0) -> List[dict]:
    """Cut [0, total) into buckets of about `bucket_elems` elements, filled from the END of the buffer
    (backward writes the last layers first), and find for each bucket the index of the last backward
    step that writes into it.

    writes[i] = [(offset, count), ...] ranges of the flat gradient buffer written by backward step i
    (ranges at or beyond `total` belong to another region and are ignored).  With align == 1 a bucket
    never splits a written range; with align > 1 (sharded optimizer: every bucket must divide evenly
    over the ranks) interior boundaries are snapped down to multiples of `align`, so a range may straddle
    two buckets and counts for the readiness of both.  `total` must be a multiple of `align`.

    head_elems > 0 splits the bucket at offset 0 once more: the FIRST layers' gradients are the last ones backward
    produces, so whatever bucket holds them is exchanged with nothing left to hide it behind, and the next forward
    pass waits for it first.  Cutting a head of at most about `head_elems` elements off that bucket leaves only a
    latency-sized exchange exposed; the remainder becomes ready as soon as its own last writer is done.  The head
    boundary is snapped UP to `align`, so a straddling range delays the (late anyway) head, not the remainder.
    Returns [{"lo", "hi", "ready_after"}] ordered by readiness (the head bucket, if one was cut, also has "head": True).
    """
    if total % align:
        raise ValueError(f"total {total} is not a multiple of align {align}")
    ranges = sorted({(o, c) for step in writes for (o, c) in step if o < total})
    last_writer = {}
    for i, step in enumerate(writes):
        for rng in step:
            if rng[0] < total:
                last_writer[rng] = max(last_writer.get(rng, -1), i)
    cuts, head_cut = [total], None
    for (o, c) in reversed(ranges):
        lo = o // align * align
        if cuts[-1] - lo >= bucket_elems and lo < cuts[-1]:
            cuts.append(lo)
    if head_elems > 0:
        heads = [-(-o // align) * align for (o, c) in ranges]
        heads = [h for h in heads if 0 < h <= head_elems and h < cuts[-1]]
        if heads and cuts[-1] - max(heads) >= head_elems:      # not worth it when the last bucket is small already
            cuts.append(max(heads))
            head_cut = max(heads)
    if cuts[-1] != 0:
        cuts.append(0)
    buckets = []
    for hi, lo in zip(cuts, cuts[1:]):
        ready = max([last_writer[(o, c)] for (o, c) in ranges if o < hi and o + c > lo], default=-1)
        buckets.append({"lo": lo, "hi": hi, "ready_after": ready})
        if hi == head_cut:
            buckets[-1]["head"] = True
    if buckets and all(b["ready_after"] < 0 for b in buckets):
        buckets[-1]["ready_after"] = 0
    for b in buckets:
        b["ready_after"] = max(b["ready_after"], 0)
    buckets.sort(key=lambda b: b["ready_after"])
    return buckets


def shard_of(bucket: dict, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the equal shard of `bucket` that `rank` owns (bucket length must divide by world)."""
    n = bucket["hi"] - bucket["lo"]
    if n % world:
        raise ValueError(f"bucket of {n} elements does not divide over {world} ranks")
    per = n // world
    return bucket["lo"] + rank * per, bucket["lo"] + (rank + 1) * per


def reduce_scatter_bucket(dist, flat, bucket, group=None, async_op=False):
    """Sum-reduce `bucket` of `flat` so that every rank ends up with the total in ITS shard (in place).
    NCCL: reduce_scatter_tensor; backends without it (gloo on CPU): all_reduce of the whole bucket."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_of(bucket, rank, world)
    if dist.get_backend(group) == "nccl":
        return dist.reduce_scatter_tensor(flat[lo:hi], flat[bucket["lo"]:bucket["hi"]], op=dist.ReduceOp.SUM, group=group,
                                          async_op=async_op)
    return dist.all_reduce(flat[bucket["lo"]:bucket["hi"]], op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def all_gather_bucket(dist, flat, bucket, group=None, async_op=False):
    """Every rank contributes its shard of `bucket`; afterwards the whole bucket is identical everywhere (in place)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_of(bucket, rank, world)
    if dist.get_backend(group) == "nccl":
        return dist.all_gather_into_tensor(flat[bucket["lo"]:bucket["hi"]], flat[lo:hi], group=group, async_op=async_op)
    parts = [flat[bucket["lo"] + r * (hi - lo):bucket["lo"] + (r + 1) * (hi - lo)] for r in range(world)]
    mine = flat[lo:hi].clone()
    return dist.all_gather(parts, mine, group=group, async_op=async_op)


def allreduce_buckets(dist, flat, buckets, group=None, async_op=False):
    """Sum-all-reduce every bucket of `flat` (in readiness order).  Returns the work handles."""
    works = []
    for b in buckets:
        w = dist.all_reduce(flat[b["lo"]:b["hi"]], op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works
