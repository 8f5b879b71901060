#!/usr/bin/env python
"""Headline benchmark: adaptive-depth SR U-Net training throughput (patches/s) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2|c2alt|c3|c1]

A "step" is one full training step (zero-grad, forward, Charbonnier loss, backward, gradient
all-reduce when N>1, Adam) on one batch of synthetic patches of the named shape.  N=1 runs
BASELINE.json's configs[1] (C2: depth 4, scale 0.25 ("x4"), 128x128 patches, batch 64, bf16); for
N>1 every rank runs the same per-GPU batch (weak scaling, batch-sharded data parallel, NCCL
all-reduce inside the captured step).  One JSON line is printed by rank 0.

  value     whole-job patches/s with inputs already resident in HBM (CUDA-graph replay)
  e2e       the same metric through the public API (Model.train_on_batch) with HOST pinned buffers:
            H2D of the batch and D2H of the loss inside the timed region, every step
  roofline  the dominant kernel (conv3x3 tcgen05 implicit GEMM, fprop+dgrad launches of one step),
            algorithmic FLOPs / CUDA-event time, against the measured sustained bf16 peak
  cpu_baseline / --impl reference
            the torch-CPU restatement of the Keras graph (oracle/), the stand-in for the reference's
            TF/Keras CPU path (TensorFlow is not installable here), on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (depth, scale, patch, per-GPU batch, description)
    "c1": (3, 0.5, 64, 8, "C1: depth 3, scale 0.5 (x2), 64x64, batch 8"),
    "c2": (4, 0.25, 128, 64, "C2: U-Net depth 4, scale 0.25 (x4 SR), 128x128 patches, batch 64 per GPU, bf16 training"),
    "c2alt": (4, 0.5, 128, 64, "C2-alt: depth 4, scale 0.5, 128x128 patches, batch 64 per GPU, bf16 training"),
    "c3": (5, 0.25, 128, 64, "C3: U-Net depth 5, scale 0.25 (x4 SR), 128x128 patches, batch 64 per GPU (global 512 at 8 GPUs)"),
}


def synth_batch(batch, patch, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    hr = rng.random((batch, patch, patch, 3), dtype=np.float32)
    lr = np.clip(hr + 0.05 * rng.standard_normal(hr.shape).astype(np.float32), 0.0, 1.0)
    return lr, hr


# --------------------------------------------------------------------------- CPU (oracle) arm
def cpu_reference_steps(depth, scale, patch, batch, steps, warmup):
    """Time the torch-CPU restatement of the Keras graph: fwd + Charbonnier + bwd + Adam."""
    import torch
    from oracle import keras_ops as K, models as M
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = M.sr_unet_spec(depth)
    ws = [torch.tensor(w, requires_grad=True) for w in M.init_weights(spec, seed=1234)]
    ms = [torch.zeros_like(w) for w in ws]
    vs = [torch.zeros_like(w) for w in ws]
    lr_np, hr_np = synth_batch(batch, patch, 1234)
    x, t = torch.from_numpy(lr_np), torch.from_numpy(hr_np)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        y = M.sr_unet_forward(ws, x, scale, depth)
        loss = K.charbonnier_loss(t, y)
        grads = torch.autograd.grad(loss, ws, allow_unused=True)
        with torch.no_grad():
            for i, (w, g) in enumerate(zip(ws, grads)):
                g = torch.zeros_like(w) if g is None else g
                p, ms[i], vs[i] = K.adam_step(w, g, ms[i], vs[i], it + 1, 1e-4)
                w.copy_(p)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    depth, scale, patch, batch, desc = CONFIGS[args.config]
    sample_batch = min(batch, 8)
    times, cores = cpu_reference_steps(depth, scale, patch, sample_batch, args.steps, args.warmup)
    total = sum(times)
    value = sample_batch * len(times) / total
    line = {
        "impl": "reference", "metric": "U-Net train patches/s", "value": value, "unit": "patches/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "depth": depth, "scale": scale, "patch": patch, "sample_batch": sample_batch,
                   "note": "torch-CPU restatement of the Keras graph (TensorFlow/Keras not installable in this image)"},
        "cpu_baseline": {"value": value, "unit": "patches/s", "cores": cores, "kind": "port",
                         "sample": f"{len(times)} steps of batch {sample_batch} (same shapes as the workload; patches/s is per sample)"},
        "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=OUT, flush=True)


# --------------------------------------------------------------------------- GPU arm
class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled every ~5 ms through NVML while the timed region runs
    (the region is only tens of milliseconds long at N=1; `nvidia-smi -lms 100` would see one sample or none)."""
    NAMES = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.rows, self.mask, self.max_mhz = index, [], 0, None
        self._stop = threading.Event()
        self.th = None

    def _handle(self):
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.index
        if vis:
            ent = [v.strip() for v in vis.split(",") if v.strip()]
            if idx < len(ent) and ent[idx].isdigit():
                idx = int(ent[idx])
            elif idx < len(ent):
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(ent[idx].encode() if hasattr(ent[idx], "encode") else ent[idx])
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def __enter__(self):
        try:
            nv, h = self._handle()
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self._stop.is_set():
                    try:
                        self.rows.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        self.mask |= int(reasons(h))
                    except Exception:
                        pass
                    time.sleep(0.005)

            self.th = threading.Thread(target=loop, daemon=True)
            self.th.start()
        except Exception:
            self.th = None
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.th is not None:
            self.th.join(timeout=1.0)

    def summary(self):
        sm = sorted(self.rows)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        reasons = [n for n, bit in self.NAMES.items() if self.mask & bit]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm)}


def conv_flops(op, batch):
    x, y = op.inputs[0], op.output
    k = op.layer.kernel_size[0]
    return 2.0 * batch * y.h * y.w * x.c * y.c * k * k


def profile_step(model, plan, st):
    """One eager step with CUDA events around every launch closure: per-kernel-class time, and the
    algorithmic FLOPs / time of the tcgen05 conv kernel (fprop + dgrad launches)."""
    import torch
    from b200unet import ops
    rec = []

    def timed(tag, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        rec.append((tag, a, b))

    model.G.zero_()
    for f in plan.pre_steps:
        timed("repack", f)
    for tag, f in zip(plan.step_tags, plan.steps):
        timed("fwd:" + tag, f)
    timed("loss", lambda: model.loss.launch(plan, st, grad_scale=1.0))
    for i, f in enumerate(plan.bwd_steps):
        timed("bwd:" + plan.bwd_tags[i], f)
    timed("adam", lambda: model.optimizer.apply(model))
    torch.cuda.synchronize()
    out = {}
    for tag, a, b in rec:
        out[tag] = out.get(tag, 0.0) + a.elapsed_time(b)
    return out


def _is_tc(op):
    from b200unet.keras.engine import Plan
    return Plan.is_tc(op)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import b200unet  # noqa: F401  (fails loudly if the CUDA library is missing)
    from b200unet import builders as B, ops
    from b200unet.keras import mixed_precision
    from b200unet.keras.optimizers import Adam

    depth, scale, patch, batch, desc = CONFIGS[args.config]
    mixed_precision.set_global_policy("mixed_bfloat16")
    model, info = B.build_super_resolution_unet(scale, depth_override=depth, input_size=patch)
    # zero-initialised head => every upstream gradient is 0; randomise it as SURVEY 8d prescribes
    import numpy as np
    rng = np.random.default_rng(1234)
    head = model.get_layer("residual_rgb")
    lim = (6.0 / (64 + 3)) ** 0.5
    head.weight_specs[0]["value"] = rng.uniform(-lim, lim, size=(1, 1, 64, 3)).astype(np.float32)
    loss, metrics = B.build_losses_and_metrics("charbonnier")
    model.compile(optimizer=Adam(learning_rate=1e-4), loss=loss, metrics=metrics, jit_compile=False)
    if world > 1:
        model.distribute()

    lr_np, hr_np = synth_batch(batch, patch, 1234 + rank)
    x_pin = torch.from_numpy(lr_np).pin_memory()
    y_pin = torch.from_numpy(hr_np).pin_memory()

    # ---- device-resident timing: graph replay only -------------------------------------------
    ops.launch_count(reset=True)
    model.train_on_batch(x_pin, y_pin)          # builds the plan, warm-up + capture
    entry = model._train_state(batch)
    launches_per_step = ops.launch_count() // 2  # body ran once eagerly (warm-up) and once under capture
    plan, st = entry["plan"], entry["state"]

    class graph:   # one step on the device-resident batch: CUDA-graph replay (segmented around the NCCL
        @staticmethod   # all-reduces when N>1), or eager launches under B200_NO_CUDA_GRAPH=1 (ncu launch list)
        def replay():
            model._run_step(entry)

    if entry["graph"] is None and entry["segments"] is None:
        launches_per_step = ops.launch_count()
    for _ in range(max(args.warmup, 3)):
        graph.replay()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        ev0.record()
        for _ in range(args.steps):
            graph.replay()
        ev1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    value = world * batch * args.steps / (total_ms / 1000.0)

    # ---- end to end through the public API: pinned host buffers, H2D + D2H every step -------------
    for _ in range(3):
        model.train_on_batch_async(x_pin, y_pin).result()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    last = prev = None
    for _ in range(args.steps):
        # every step copies ITS batch from pinned host memory (copy stream, overlapping the previous step's kernels)
        # and its loss/metric is read back to the host; one step in flight while the next is enqueued
        h = model.train_on_batch_async(x_pin, y_pin)
        if prev is not None:
            last = prev.result()                    # D2H read + wait of the previous step
        prev = h
    last = prev.result()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * batch * args.steps / float(e2e_s.item())
    h2d = x_pin.numel() * 4 + y_pin.numel() * 4
    d2h = 4 * len(last)

    if rank != 0:
        if world > 1:
            model.release_graphs()
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- per-kernel breakdown and roofline of the dominant kernel (rank 0, eager, CUDA events) ----
    prof = profile_step(model, plan, st)
    for _ in range(2):
        p2 = profile_step(model, plan, st)
        prof = {k: min(v, p2[k]) for k, v in prof.items()}
    step_ms_eager = sum(prof.values())
    conv_ops = [op for op in plan.ops if op.kind == "conv" and _is_tc(op)]
    flops_fprop = sum(conv_flops(op, batch) for op in conv_ops)
    flops_dgrad = sum(conv_flops(op, batch) for op in conv_ops if op.inputs[0].needs_grad)
    # every launch of conv3x3_tc_kernel: plain fprop, fprop with the fused LayerNorm epilogue, dgrad
    t_conv = prof.get("fwd:conv:tc", 0.0) + prof.get("fwd:conv+ln:tc", 0.0) + prof.get("bwd:dgrad:tc", 0.0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    ach = (flops_fprop + flops_dgrad) / (t_conv / 1000.0) / 1e12 if t_conv > 0 else 0.0
    t_wg = prof.get("bwd:wgrad:tc", 0.0)
    ach_wg = flops_fprop / (t_wg / 1000.0) / 1e12 if t_wg > 0 else 0.0
    n_launch = sum(1 for op in conv_ops) + sum(1 for op in conv_ops if op.inputs[0].needs_grad)
    # algorithmic HBM bytes of those launches: input read once + output written once (bf16), x2 outputs with fused LN
    def _act_bytes(v):
        return 2.0 * batch * v.h * v.w * v.c
    bytes_conv = sum(_act_bytes(op.inputs[0]) + _act_bytes(op.output) * (2 if getattr(op, "fused_into_ln", False) else 1)
                     for op in conv_ops)
    bytes_conv += sum(_act_bytes(op.inputs[0]) + _act_bytes(op.output) for op in conv_ops if op.inputs[0].needs_grad)
    # the same figure over the launches that are conv3x3_tc_kernel proper (images larger than 4x4; the deep levels run
    # the split-K conv_gemm_kernel, whose traffic is weights) -- the population of the ncu `traffic` number
    win_ops = [op for op in conv_ops if max(op.output.h, op.output.w) > 4]
    bytes_win = sum(_act_bytes(op.inputs[0]) + _act_bytes(op.output) * (2 if getattr(op, "fused_into_ln", False) else 1)
                    for op in win_ops)
    bytes_win += sum(_act_bytes(op.inputs[0]) + _act_bytes(op.output) for op in win_ops if op.inputs[0].needs_grad)
    n_win = len(win_ops) + sum(1 for op in win_ops if op.inputs[0].needs_grad)
    traffic = None            # DRAM bytes per launch of the same kernel, from the committed `ncu --set full` capture
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_v9_conv_traffic.json")))["dram_bytes_per_launch_mean"]
    except Exception:
        pass
    roofline = {
        "kernel": "conv3x3_tc_kernel (tcgen05 implicit GEMM; every fprop / fprop+LayerNorm / dgrad launch of one step, incl. the split-K conv_gemm_kernel launches of the deep levels)",
        "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
        "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1400 (of fallback)",
        "traffic": traffic, "traffic_source": "profiles/r01_v9_conv_traffic.json (ncu dram__bytes_read+write, mean over the 33 conv3x3_tc_kernel launches of one step)",
        "algorithmic_bytes_per_launch": bytes_win / max(n_win, 1),
        "launches_per_step": n_launch, "avg_launch_ms": t_conv / max(n_launch, 1),
        "wgrad_kernel": {"achieved": ach_wg, "frac": ach_wg / peak_tf, "ms_per_step": t_wg},
    }
    # algorithmic FLOPs of one step: 3 x the forward FLOPs of every Conv2D of the model's own graph
    step_tflop = 3.0 * sum(conv_flops(op, batch) for op in plan.ops if op.kind == "conv") / 1e12
    ms_per_step = total_ms / args.steps

    # ---- CPU baseline: the oracle, bounded sample ---------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        sb = 8                                   # ~10 s of host work on the box's cores (0.7 s per batch-8 step at 16 cores)
        times, cores = cpu_reference_steps(depth, scale, patch, sb, 12, 1)
        cpu = {"value": sb * len(times) / sum(times), "unit": "patches/s", "cores": cores, "kind": "port",
               "sample": f"{len(times)} steps of batch {sb} of the same workload shapes, torch-CPU restatement of the Keras graph"}

    line = {
        "metric": "U-Net train patches/s", "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": desc, "depth": depth, "scale": scale, "patch": patch, "per_gpu_batch": batch,
                   "global_batch": batch * world, "loss": "charbonnier", "optimizer": "adam",
                   "parallelism": f"dp{world}", "streams": "dgrad / norm backward on the main stream, wgrad forked to a second stream inside the captured step" if getattr(model, "overlap_wgrad", False) else "single stream", "l2": "per-step working set (>2 GB of activations) far exceeds the 126 MB L2",
                   "params": model.count_params(), "step_tflop_algorithmic": step_tflop,
                   "step_frac_of_conv_roofline": step_tflop / (ms_per_step / 1000.0) / peak_tf},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "patches/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "breakdown_ms": {k: round(v, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1])},
        "eager_step_ms": step_ms_eager,
    }
    print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        model.release_graphs()
        dist.barrier()
        dist.destroy_process_group()


def _claim_stdout():
    """Route everything libraries write to fd 1 (NCCL prints its version banner there) to stderr and return a
    file object on the real stdout for the ONE JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    global OUT
    OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
