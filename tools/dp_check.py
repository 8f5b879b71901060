"""DP equivalence on real GPUs (run under torchrun, N >= 2): the all-reduced gradient of the sharded
global batch must equal the single-GPU gradient of the whole batch, and the weights after one Adam
step must agree on every rank.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py
B200_DP_CHECK_ONE_GPU=1 (with B200_DP_SHARD=0) runs every rank on cuda:0 over the gloo backend: the same segmented
capture / bucket exchange / fork-join logic, checkable on a one-GPU box (NCCL refuses two ranks on one device)."""
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
    from b200unet.keras import clear_session, mixed_precision
    from b200unet.keras.optimizers import Adam
    from b200unet.parallel import shard_range
    from oracle import models as M

    scale, depth, P, GB = 0.5, 3, 64, 8 * world
    ws_np = M.init_weights(M.sr_unet_spec(depth), seed=1234, jitter=0.05)
    rng = np.random.default_rng(7)
    hr = rng.random((GB, P, P, 3), dtype=np.float32)
    lr = np.clip(hr + 0.05 * rng.standard_normal(hr.shape).astype(np.float32), 0, 1)

    def make(distributed):
        clear_session()
        mixed_precision.set_global_policy("mixed_bfloat16")
        model, _ = B.build_super_resolution_unet(scale, depth_override=depth, input_size=P)
        model.set_weights(ws_np)
        loss, metrics = B.build_losses_and_metrics("charbonnier")
        model.compile(optimizer=Adam(1e-3), loss=loss, metrics=metrics)
        if distributed:
            model.distribute()
        return model

    lo, hi = shard_range(GB, rank, world)
    m_dp = make(True)
    logs = m_dp.train_on_batch(lr[lo:hi], hr[lo:hi])
    torch.cuda.synchronize()
    g_dp = m_dp.gathered_gradients()      # collective: under the sharded optimizer a rank holds only its shards' sums
    m_dp._sync_master()                   # collective: all-gather the fp32 master of the kernel region
    p_dp = m_dp.P.clone()
    s_dp = m_dp.S.clone().float()
    s_ref = s_dp.clone()
    dist.broadcast(s_ref, src=0)
    assert bool((s_ref == s_dp).all().item()), "compute-dtype shadow differs between ranks"
    # every rank must hold identical weights after the step
    p_ref = p_dp.clone()
    dist.broadcast(p_ref, src=0)
    same = bool((p_ref == p_dp).all().item())
    # a few more steps (the exchange's cross-step ordering: gradient buffer reuse, shadow pulls overlapping the forward pass)
    more = int(os.environ.get("B200_DP_CHECK_STEPS", "4"))
    for _ in range(more):
        m_dp.train_on_batch(lr[lo:hi], hr[lo:hi])
    torch.cuda.synchronize()
    m_dp._sync_master()
    p_more = m_dp.P.clone()
    p_ref = p_more.clone()
    dist.broadcast(p_ref, src=0)
    same_more = bool((p_ref == p_more).all().item())
    peer = getattr(m_dp, "_peer", None) is not None
    if rank == 0:
        m_1 = make(False)
        m_1.train_on_batch(lr, hr)
        torch.cuda.synchronize()
        rel = ((g_dp - m_1.G).norm() / m_1.G.norm()).item()
        relp = ((p_dp - m_1.P).norm() / m_1.P.norm()).item()
        for _ in range(more):
            m_1.train_on_batch(lr, hr)
        torch.cuda.synchronize()
        relm = ((p_more - m_1.P).norm() / m_1.P.norm()).item()
        print(f"DP check world={world} (exchange: {'peer memory' if peer else 'NCCL'}): grad rel-L2 vs single-GPU full batch "
              f"{rel:.3e}; weights after Adam {relp:.3e}; after {more} more steps {relm:.3e}; ranks identical: "
              f"{same and same_more}; loss(rank0 shard) {logs['loss']:.6f}", flush=True)
        assert rel < 2e-2 and same and same_more and relm < 2e-3
    else:
        assert same and same_more
    m_dp.release_graphs()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
