"""Data-parallel host logic on CPU: world_size-2 gloo run of the batch-sharding recipe.

Each rank evaluates the CPU oracle on its contiguous shard with the loss gradient scaled by 1/world,
flattens the gradients, and sum-all-reduces them bucket by bucket with the product's own
``parallel.plan_buckets`` / ``allreduce_buckets``; the result must equal the single-process gradient
of the concatenated batch (mean-of-shard-means == global mean for equal shards)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import keras_ops as K, models as M


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _flat_grads(ws_np, x, t, scale, depth, grad_scale):
    ws = [torch.tensor(w, requires_grad=True) for w in ws_np]
    loss = K.charbonnier_loss(t, M.sr_unet_forward(ws, x, scale, depth)) * grad_scale
    gs = torch.autograd.grad(loss, ws)
    offs, off = [], 0
    for g in gs:
        offs.append((off, g.numel())); off += (g.numel() + 63) // 64 * 64
    flat = torch.zeros(off)
    for (o, n), g in zip(offs, gs):
        flat[o:o + n] = g.reshape(-1)
    return flat, offs


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200unet.parallel import allreduce_buckets, plan_buckets, shard_range
    torch.set_num_threads(1)
    scale, depth, P, B = 0.5, 1, 16, 4
    ws_np = M.init_weights(M.sr_unet_spec(depth), seed=3, jitter=0.05)
    rng = np.random.default_rng(0)
    hr = torch.from_numpy(rng.random((B, P, P, 3), dtype=np.float32))
    lr = torch.clamp(hr + 0.05 * torch.randn(hr.shape, generator=torch.Generator().manual_seed(1)), 0, 1)
    lo, hi = shard_range(B, rank, world)
    flat, offs = _flat_grads(ws_np, lr[lo:hi], hr[lo:hi], scale, depth, 1.0 / world)
    # backward writes the last parameters first: one "step" per parameter, in reverse order
    writes = [[rng_] for rng_ in reversed(offs)]
    buckets = plan_buckets(flat.numel(), writes, bucket_elems=5000)
    assert sum(b["hi"] - b["lo"] for b in buckets) == flat.numel()
    assert [b["ready_after"] for b in buckets] == sorted(b["ready_after"] for b in buckets)
    works = allreduce_buckets(dist, flat, buckets, async_op=True)
    for w in works:
        w.wait()
    if rank == 0:
        ref, _ = _flat_grads(ws_np, lr, hr, scale, depth, 1.0)
        out.put(((flat - ref).norm() / ref.norm()).item())
    dist.barrier()
    dist.destroy_process_group()


def _worker_sharded(rank, world, port, out):
    """Sharded optimizer on CPU tensors: reduce-scatter (gloo: emulated by all-reduce) -> Adam on the owned shard ->
    all-gather; every rank must end with the parameters of a single-process Adam step on the summed gradient."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200unet.parallel import all_gather_bucket, plan_buckets, reduce_scatter_bucket, shard_of
    torch.manual_seed(0)
    total, align = 64 * world * 37, 64 * world
    p0 = torch.randn(total)
    g_ranks = [torch.randn(total, generator=torch.Generator().manual_seed(10 + r)) for r in range(world)]
    writes = [[(o, 700)] for o in reversed(range(0, total - 700, 700))] + [[(total - total % 700 if total % 700 else total - 700, 10)]]
    buckets = plan_buckets(total, writes, bucket_elems=3000, align=align)
    assert sum(b["hi"] - b["lo"] for b in buckets) == total and all((b["hi"] - b["lo"]) % align == 0 for b in buckets)
    g = g_ranks[rank].clone()
    for b in buckets:
        reduce_scatter_bucket(dist, g, b)
    p = p0.clone()
    m = torch.zeros(total); v = torch.zeros(total)
    for b in buckets:
        lo, hi = shard_of(b, rank, world)
        pn, _, _ = K.adam_step(p[lo:hi], g[lo:hi], m[lo:hi], v[lo:hi], 1, 1e-3)
        p[lo:hi] = pn
    for b in buckets:
        all_gather_bucket(dist, p, b)
    ref, _, _ = K.adam_step(p0, sum(g_ranks), torch.zeros(total), torch.zeros(total), 1, 1e-3)
    out.put((rank, ((p - ref).norm() / ref.norm()).item()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_optimizer_equivalence_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    errs = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert set(errs) == {0, 1} and max(errs.values()) < 1e-6, errs


def test_dp_gradient_equivalence_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-5, err


def test_shard_range_and_bucket_plan():
    from b200unet.parallel import plan_buckets, shard_range
    assert shard_range(512, 3, 8) == (192, 256)
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)
    writes = [[(900, 100)], [(800, 50), (850, 50)], [(0, 800)]]
    b = plan_buckets(1000, writes, 120)
    assert b == [{"lo": 850, "hi": 1000, "ready_after": 1}, {"lo": 0, "hi": 850, "ready_after": 2}]
    # aligned plan (sharded optimizer): boundaries are multiples of `align`, a straddling range counts for both buckets
    a = plan_buckets(1024, [[(960, 64)], [(500, 460)], [(0, 500)]], 200, align=128)
    assert all(x["lo"] % 128 == 0 and x["hi"] % 128 == 0 for x in a) and sum(x["hi"] - x["lo"] for x in a) == 1024
    lows = sorted(x["lo"] for x in a)
    assert lows[0] == 0 and {x["ready_after"] for x in a if x["lo"] <= 500 < x["hi"]} == {2}
    # head bucket: the first layers (written last) get a small bucket of their own, boundary snapped UP so that the
    # straddling range delays the head and not the remainder
    wr = [[(900, 124)], [(300, 600)], [(100, 200)], [(0, 100)]]
    h = plan_buckets(1024, wr, 2000, align=64, head_elems=150)
    assert h == [{"lo": 128, "hi": 1024, "ready_after": 2}, {"lo": 0, "hi": 128, "ready_after": 3, "head": True}]
    assert plan_buckets(1024, wr, 2000, align=64, head_elems=1000) == [{"lo": 0, "hi": 1024, "ready_after": 3}]
    assert plan_buckets(1024, wr, 2000, align=64) == [{"lo": 0, "hi": 1024, "ready_after": 3}]
    # every element is covered exactly once whatever the bucket size
    for size in (1, 64, 10_000):
        bs = sorted(plan_buckets(1000, writes, size), key=lambda d: d["lo"])
        assert bs[0]["lo"] == 0 and bs[-1]["hi"] == 1000
        assert all(a["hi"] == c["lo"] for a, c in zip(bs, bs[1:]))


def test_complement_ranges():
    """The flat-buffer ranges left for the final optimizer pass after some were updated early (Model._apply_optimizer)."""
    from b200unet.parallel import complement_ranges as C
    assert C([], 10) == [(0, 10)]
    assert C([(0, 10)], 10) == []
    assert C([(2, 4), (6, 8)], 10) == [(0, 2), (4, 6), (8, 10)]
    assert C([(6, 8), (2, 4), (3, 7)], 10) == [(0, 2), (8, 10)]            # unordered, overlapping
    assert C([(0, 3), (3, 5)], 5) == []                                     # touching
    assert C([(4, 20)], 10) == [(0, 4)]                                     # past the end
    rng = np.random.default_rng(0)
    for _ in range(50):
        total = int(rng.integers(1, 60))
        done = [tuple(sorted(rng.integers(0, total + 5, 2).tolist())) for _ in range(int(rng.integers(0, 6)))]
        mask = np.zeros(total, bool)
        for lo, hi in done:
            mask[lo:hi] = True
        got = np.zeros(total, bool)
        for lo, hi in C(done, total):
            assert 0 <= lo < hi <= total and not got[lo:hi].any()
            got[lo:hi] = True
        assert np.array_equal(got, ~mask)
