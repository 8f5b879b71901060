"""Where the data-parallel step spends its time (run under torchrun, N >= 2):

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/dp_breakdown.py [c2|c3]

  replica      the same model, one captured graph, no exchange (what perfect scaling would give)
  segments     the data-parallel capture (backward cut at the bucket boundaries, Adam in its own graph) replayed WITHOUT the
               collectives: what the segmentation itself costs
  dp           the full step: reduce-scatter per bucket behind its segment, sharded Adam, all-gather of the bf16 shadow
  collectives  every bucket's reduce-scatter and all-gather timed alone, back to back (CUDA events, max over ranks)
Exposed exchange = dp - segments; segmentation = segments - replica."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from b200unet.parallel import all_gather_bucket, reduce_scatter_bucket
    depth, scale, patch, batch, desc = bench.CONFIGS[cfg]
    lr_np, hr_np = bench.synth_batch(batch, patch, 1234 + rank)
    x, y = torch.from_numpy(lr_np).pin_memory(), torch.from_numpy(hr_np).pin_memory()
    steps, out = 20, {"config": desc, "world": world, "per_gpu_batch": batch}

    model = bench._build_sr(cfg, 1)
    model.train_on_batch(x, y)
    e = model._train_state(batch)
    ms, _ = bench._timed_steps(lambda: model._run_step(e), steps, 5, world, local)
    out["replica_ms"] = ms / steps
    model.release_graphs(); del model, e
    torch.cuda.empty_cache()

    model = bench._build_sr(cfg, world)
    model.train_on_batch(x, y)
    e = model._train_state(batch)
    plan = e["plan"]
    buckets = model._buckets(plan)
    out["buckets"] = [{"mb_fp32": (b["hi"] - b["lo"]) * 4 / 2**20, "sharded": b["sharded"], "ready_after_step": b["ready_after"],
                       "of_steps": len(plan.bwd_steps)} for b in buckets]
    out["segments"] = len(e["segments"])
    out["forward_segments"] = len(e["fsegs"])
    out["pipelined_optimizer"] = "adam_buckets" in e
    out["exchange"] = "peer memory (copy engines)" if getattr(model, "_peer", None) is not None else "NCCL"

    def no_comm():
        for g, _w in e["fsegs"]:
            g.replay()
        for g, _b in e["segments"]:
            g.replay()
        e["adam_graph"].replay()

    ms, _ = bench._timed_steps(no_comm, steps, 5, world, local)
    out["segments_ms"] = ms / steps
    ms, _ = bench._timed_steps(lambda: model._run_step(e), steps, 5, world, local)
    out["dp_ms"] = ms / steps
    # the collectives alone
    coll = []
    dgroup = model._dist[1]
    for b in buckets:
        def rs(b=b):
            if b["sharded"]:
                reduce_scatter_bucket(dist, model.G, b, group=dgroup)
            else:
                dist.all_reduce(model.G[b["lo"]:b["hi"]], group=dgroup)
        t_rs, _ = bench._timed_steps(rs, 10, 3, world, local)
        t_ag = None
        if b["sharded"]:
            t_ag, _ = bench._timed_steps(lambda b=b: all_gather_bucket(dist, model.S, b, group=dgroup), 10, 3, world, local)
        coll.append({"mb_fp32": (b["hi"] - b["lo"]) * 4 / 2**20, "reduce_scatter_us": t_rs / 10 * 1e3,
                     "all_gather_bf16_us": None if t_ag is None else t_ag / 10 * 1e3})
    out["collectives"] = coll
    out["exposed_exchange_ms"] = out["dp_ms"] - out["segments_ms"]
    out["segmentation_ms"] = out["segments_ms"] - out["replica_ms"]
    out["efficiency_vs_replica"] = out["replica_ms"] / out["dp_ms"]
    if rank == 0:
        print(json.dumps(out))
    model.release_graphs()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
