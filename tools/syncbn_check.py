"""Synchronised BatchNorm under data parallelism: N ranks x B/N samples must train like one rank x B samples.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/syncbn_check.py
B200_DP_CHECK_ONE_GPU=1 runs every rank on cuda:0 over gloo (a one-GPU box; NCCL refuses two ranks on one device);
use it with B200_DP_SHARD=0 (gloo has no reduce-scatter on CUDA tensors).

The BatchNorm segmentation net (Segmenation/code/train_adaptive_unet.py:325-362 of the reference) on a global batch:
the sharded run (per-channel sums of z, z^2 / g, g*xhat all-reduced inside the forward / backward pass) against the
same model on the whole batch in one process -- loss, every weight gradient, moving statistics, weights after Adam."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if os.environ.get("B200_DP_CHECK_ONE_GPU") == "1":
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from b200unet import builders as B
    from b200unet.keras import clear_session, losses as LS, mixed_precision
    from b200unet.keras.optimizers import Adam
    from b200unet.parallel import shard_range
    from oracle import models as M

    depth, base, P, GB = 2, 64, 32, 4 * world
    ws_np = M.init_weights(M.seg_adaptive_spec(depth, base), seed=7, jitter=0.05)
    rng = np.random.default_rng(3)
    x = rng.random((GB, P, P, 3), dtype=np.float32)
    t = (rng.random((GB, P, P, 1)) > 0.5).astype(np.float32)

    def make(distributed):
        clear_session()
        mixed_precision.set_global_policy("float32")
        model = B.build_adaptive_depth_unet(P, base, depth)
        model.set_weights(ws_np)
        model.compile(optimizer=Adam(1e-3), loss=LS.make_hybrid_ce_dice_loss(0.4, 0.6), metrics=[LS.dice_metric, LS.iou_metric])
        if distributed:
            model.distribute()
        return model

    lo, hi = shard_range(GB, rank, world)
    m_dp = make(True)
    logs = m_dp.train_on_batch(x[lo:hi], t[lo:hi])
    torch.cuda.synchronize()
    assert m_dp._train_state(hi - lo)["plan"].sync_bn, "the distributed BatchNorm model must run synchronised BatchNorm"
    g_dp = m_dp.gathered_gradients()
    m_dp._sync_master()
    w_dp = m_dp.get_weights()
    loss_dp = torch.tensor([logs["loss"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(loss_dp)

    m_1 = make(False)
    logs1 = m_1.train_on_batch(x, t)
    torch.cuda.synchronize()
    g_1 = m_1.G.clone()
    w_1 = m_1.get_weights()

    def rel(a, b):
        a, b = a.double().flatten(), b.double().flatten()
        return float((a - b).norm() / (b.norm() + 1e-30))

    worst = 0.0
    for ly in m_1.layers:
        for w in ly.weight_specs:
            if not w["trainable"]:
                continue
            off, n = m_1._grad_range(ly, w["name"].split("/", 1)[1])
            if g_1[off:off + n].abs().max() < 1e-9:
                continue
            e = rel(g_dp[off:off + n], g_1[off:off + n])
            worst = max(worst, e)
            # BatchNorm backward is ill-conditioned in fp32 (the fp32 torch oracle itself sits 3e-3 from fp64): summation
            # order differs between 2 x 4 and 1 x 8 samples
            assert e < 2e-2, (w["name"], e)
    werr = max(rel(torch.from_numpy(a), torch.from_numpy(b)) for a, b in zip(w_dp, w_1))
    # per-replica statistics (no synchronisation) would put the moving means several percent apart
    stats = [(a, b) for (a, b), sp in zip(zip(w_dp, w_1), [w for ly in m_1.layers for w in ly.weight_specs]) if not sp["trainable"]]
    serr = max(rel(torch.from_numpy(a), torch.from_numpy(b)) for a, b in stats)
    print(f"rank {rank}: loss dp(mean over ranks) {loss_dp.item() / world:.7f} vs single {logs1['loss']:.7f}; "
          f"worst gradient rel-L2 {worst:.2e}; weights after Adam {werr:.2e}; moving statistics {serr:.2e}", flush=True)
    # the hybrid loss has a per-sample Dice term: mean of shard means == global mean for equal shards
    assert abs(loss_dp.item() / world - logs1["loss"]) < 1e-5 * max(1.0, abs(logs1["loss"]))
    # (the first Adam step is sign-like: gradients that differ in their last bits move the weights by up to 2 lr)
    assert serr < 1e-5 and werr < 5e-3
    m_dp.release_graphs()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("SYNCBN_OK")


if __name__ == "__main__":
    main()
