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
    # every element is covered exactly once whatever the bucket size
    for size in (1, 64, 10_000):
        bs = sorted(plan_buckets(1000, writes, size), key=lambda d: d["lo"])
        assert bs[0]["lo"] == 0 and bs[-1]["hi"] == 1000
        assert all(a["hi"] == c["lo"] for a, c in zip(bs, bs[1:]))
