#!/usr/bin/env python
"""Headline benchmark: adaptive-depth SR U-Net training throughput (patches/s) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2|c2alt|c3|c1|c4|c5]

A "step" is one full training step (zero-grad, forward, Charbonnier loss, backward, gradient
all-reduce when N>1, Adam) on one batch of synthetic patches of the named shape.  N=1 runs
BASELINE.json's configs[1] (C2: depth 4, scale 0.25 ("x4"), 128x128 patches, batch 64, bf16); for
N>1 every rank runs the same per-GPU batch (weak scaling, batch-sharded data parallel, NCCL
all-reduce inside the captured step).  One JSON line is printed by rank 0.

  value     whole-job patches/s with inputs already resident in HBM (CUDA-graph replay)
  e2e       the same metric through the public API (Model.train_on_batch) with HOST pinned buffers:
            H2D of the batch and D2H of the loss inside the timed region, every step
  roofline  the dominant kernel (conv3x3 tcgen05 implicit GEMM, fprop+dgrad launches of one step),
            algorithmic FLOPs / CUDA-event time per launch (each launch timed alone, SM clock at boost), against the
            measured BURST bf16 peak; `frac_of_sustained` gives the same against the sustained peak
  sustained >= 3 s of back-to-back graph replays with the NVML clock trace: patches/s, and the step's algorithmic
            FLOP/s against the measured SUSTAINED bf16 peak (the 20..50-step timed region of `value` lasts ~0.1 s)
  c3        (N > 1 only) BASELINE configs[2] -- depth 5, scale 0.25, global batch 512 split over the N ranks -- measured in
            the same run: data-parallel value, and the same per-GPU batch as N independent replicas (no exchange)
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
    # BASELINE configs[3] / configs[4]: different model families, their own lines (run_c4 / run_c5)
    "c4": (4, None, 256, 32, "C4: segmentation U-Net depth 4 (build_unet, base 32), 256x256 RGB, 21 classes, softmax-CE training, batch 32 per GPU"),
    "c5": (5, 0.25, 1024, 16, "C5: U-Net depth 5 SR inference, scale 0.25, 1024x1024 full images, batch 16 per GPU, forward only"),
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


def cpu_reference_aux(config, steps, warmup):
    """CPU arm of the two auxiliary configs: C4 = one training step of the segmentation net (softmax-CE + Adam) on a
    batch of 2; C5 = one forward pass of the depth-5 SR net on ONE 512x512 image (a quarter of a 1024x1024 one)."""
    import numpy as np
    import torch
    from oracle import keras_ops as K, models as M
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.default_rng(1234)
    times = []
    if config == "c4":
        classes, sb, size = 21, 2, 256
        ws = [torch.tensor(w, requires_grad=True) for w in M.init_weights(M.seg_vanilla_spec(4, 32, classes), seed=1234)]
        ms, vs = [torch.zeros_like(w) for w in ws], [torch.zeros_like(w) for w in ws]
        x = torch.from_numpy(rng.random((sb, size, size, 3), dtype=np.float32))
        t = torch.nn.functional.one_hot(torch.from_numpy(rng.integers(0, classes, (sb, size, size))), classes).float()
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            loss = K.categorical_crossentropy(t, M.seg_vanilla_forward(ws, x, 4, classes))
            grads = torch.autograd.grad(loss, ws, allow_unused=True)
            with torch.no_grad():
                for i, (w, g) in enumerate(zip(ws, grads)):
                    p, ms[i], vs[i] = K.adam_step(w, torch.zeros_like(w) if g is None else g, ms[i], vs[i], it + 1, 1e-4)
                    w.copy_(p)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        return times, cores, sb, f"{steps} training steps of batch {sb} at 256x256 (same model, same loss)"
    ws = [torch.tensor(w) for w in M.init_weights(M.sr_unet_spec(5), seed=1234)]
    x = torch.from_numpy(rng.random((1, 512, 512, 3), dtype=np.float32))
    with torch.no_grad():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            M.sr_unet_forward(ws, x, 0.25, 5)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    return times, cores, 0.25, f"{steps} forward passes of one 512x512 image = a quarter of a 1024x1024 image each"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    depth, scale, patch, batch, desc = CONFIGS[args.config]
    if args.config in ("c4", "c5"):
        times, cores, per_step, sample = cpu_reference_aux(args.config, min(args.steps, 5), min(args.warmup, 1))
        metric = "segmentation U-Net train images/s" if args.config == "c4" else "SR U-Net inference images/s"
        unit, sample_batch = "images/s", per_step
    else:
        sample_batch = min(batch, 8)
        times, cores = cpu_reference_steps(depth, scale, patch, sample_batch, args.steps, args.warmup)
        metric, unit = "U-Net train patches/s", "patches/s"
        sample = f"{len(times)} steps of batch {sample_batch} (same shapes as the workload; patches/s is per sample)"
    total = sum(times)
    value = sample_batch * len(times) / total
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "depth": depth, "scale": scale, "patch": patch, "sample_batch": sample_batch,
                   "note": "torch-CPU restatement of the Keras graph (TensorFlow/Keras not installable in this image)"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    algorithmic FLOPs / time of the tcgen05 conv kernel (fprop + dgrad launches).  The stream is first held busy by a
    ~10 ms spin kernel while the host enqueues the whole step, so that the events bracket DEVICE time only -- an eager
    launch otherwise shows a ~10 us floor per launch that is the host's enqueue latency, not the kernel."""
    import torch
    from b200unet import ops
    rec = []
    torch.cuda.synchronize()
    torch.cuda._sleep(20_000_000)

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


def _build_sr(config, world, per_gpu_batch=None):
    """The SR U-Net of a named config, compiled (Charbonnier + Adam) and, when world > 1, set up for data parallelism."""
    import numpy as np
    from b200unet import builders as B
    from b200unet.keras import clear_session, mixed_precision
    from b200unet.keras.optimizers import Adam
    depth, scale, patch, batch, desc = CONFIGS[config]
    clear_session()
    mixed_precision.set_global_policy("mixed_bfloat16")
    model, info = B.build_super_resolution_unet(scale, depth_override=depth, input_size=patch)
    # zero-initialised head => every upstream gradient is 0; randomise it as SURVEY 8d prescribes
    rng = np.random.default_rng(1234)
    head = model.get_layer("residual_rgb")
    lim = (6.0 / (64 + 3)) ** 0.5
    head.weight_specs[0]["value"] = rng.uniform(-lim, lim, size=(1, 1, 64, 3)).astype(np.float32)
    loss, metrics = B.build_losses_and_metrics("charbonnier")
    model.compile(optimizer=Adam(learning_rate=1e-4), loss=loss, metrics=metrics, jit_compile=False)
    if world > 1:
        model.distribute()
    return model


def _timed_steps(step_fn, steps, warmup, world, local):
    """W untimed + K timed calls of `step_fn`, bracketed by barrier + synchronize, CUDA events on the launching
    stream, MAX over ranks; SM clock / throttle reasons sampled during the timed region.  -> (total_ms, clocks)"""
    import torch
    import torch.distributed as dist
    for _ in range(max(warmup, 3)):
        step_fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        ev0.record()
        for _ in range(steps):
            step_fn()
        ev1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), clocks.summary()


def _c3_subrecord(args, world, rank, local):
    """BASELINE configs[2] in the same run: depth 5, scale 0.25, 128x128, GLOBAL batch 512 split over the N ranks --
    (a) batch-sharded data parallel (the exchange inside the step), (b) the same per-GPU batch as N independent replicas
    running at the same time with no exchange; (a) / (b) is what the gradient exchange costs at this config."""
    import torch
    per_gpu = 512 // world
    steps = min(args.steps, 20)
    depth, scale, patch, _, desc = CONFIGS["c3"]
    lr_np, hr_np = synth_batch(per_gpu, patch, 4321 + rank)
    x_pin, y_pin = torch.from_numpy(lr_np).pin_memory(), torch.from_numpy(hr_np).pin_memory()
    out = {}
    for mode in ("dp", "replicas"):
        model = _build_sr("c3", world if mode == "dp" else 1)
        model.train_on_batch(x_pin, y_pin)
        entry = model._train_state(per_gpu)
        total_ms, clocks = _timed_steps(lambda: model._run_step(entry), steps, args.warmup, world, local)
        out[mode] = (total_ms, clocks)
        model.release_graphs()
        del model, entry
        torch.cuda.empty_cache()
    dp_ms, rep_ms = out["dp"][0], out["replicas"][0]
    return {"workload": "C3: U-Net depth 5, scale 0.25 (x4 SR), 128x128 patches, global batch 512, data parallel",
            "global_batch": per_gpu * world, "per_gpu_batch": per_gpu, "steps": steps, "params": 138427843,
            "value": per_gpu * world * steps / (dp_ms / 1000.0), "unit": "patches/s", "ms_per_step": dp_ms / steps,
            "replicas_value": per_gpu * world * steps / (rep_ms / 1000.0), "replicas_ms_per_step": rep_ms / steps,
            "efficiency_vs_replicas": rep_ms / dp_ms, "clocks": out["dp"][1],
            "note": "replicas = the same model and per-GPU batch on every rank at the same time, no gradient exchange "
                    "(max over ranks); efficiency_vs_replicas = replicas_ms / dp_ms"}


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
    from b200unet import ops

    if args.config in ("c4", "c5"):
        line = (run_c4 if args.config == "c4" else run_c5)(args, world, rank, local)
        if rank == 0:
            print(json.dumps(line), file=OUT, flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    depth, scale, patch, batch, desc = CONFIGS[args.config]
    model = _build_sr(args.config, world)
    lr_np, hr_np = synth_batch(batch, patch, 1234 + rank)
    x_pin = torch.from_numpy(lr_np).pin_memory()
    y_pin = torch.from_numpy(hr_np).pin_memory()

    # ---- device-resident timing: graph replay only -------------------------------------------
    ops.launch_count(reset=True)
    model.train_on_batch(x_pin, y_pin)          # builds the plan, warm-up + capture
    entry = model._train_state(batch)
    launches_per_step = ops.launch_count() // 2  # body ran once eagerly (warm-up) and once under capture
    plan, st = entry["plan"], entry["state"]
    if entry["graph"] is None and entry["segments"] is None:
        launches_per_step = ops.launch_count()

    def replay():   # one step on the device-resident batch: CUDA-graph replay (segmented around the NCCL exchanges when
        model._run_step(entry)   # N>1), or eager launches under B200_NO_CUDA_GRAPH=1 (ncu launch list)

    total_ms, clocks = _timed_steps(replay, args.steps, args.warmup, world, local)
    value = world * batch * args.steps / (total_ms / 1000.0)
    ms_per_step = total_ms / args.steps

    # ---- sustained: >= 3 s of back-to-back steps (the region above lasts ~0.1 s at boost clock) ------------
    sustained = None
    if not args.no_sustained:
        n_sus = int(min(max(args.steps, 3000.0 / max(ms_per_step, 1e-3) + 1), 4000))
        sus_ms, sus_clocks = _timed_steps(replay, n_sus, 0, world, local)
        sustained = {"seconds": sus_ms / 1000.0, "steps": n_sus, "value": world * batch * n_sus / (sus_ms / 1000.0),
                     "unit": "patches/s", "ms_per_step": sus_ms / n_sus, "clocks": sus_clocks}

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

    # ---- BASELINE configs[2] (depth 5, global batch 512) measured in the same multi-GPU run --------------
    c3 = None
    if world > 1 and args.config == "c2" and not args.no_c3:
        # per-kernel profile of the main config first (rank 0), while its plan is alive
        prof = _profile(model, plan, st) if rank == 0 else None
        model.release_graphs()
        del entry
        c3 = _c3_subrecord(args, world, rank, local)
    else:
        prof = _profile(model, plan, st) if rank == 0 else None

    if rank != 0:
        if world > 1:
            model.release_graphs()
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rank 0, eager, CUDA events around every launch) ----
    step_ms_eager = sum(prof.values())
    conv_ops = [op for op in plan.ops if op.kind == "conv" and _is_tc(op)]
    flops_fprop = sum(conv_flops(op, batch) for op in conv_ops)
    flops_dgrad = sum(conv_flops(op, batch) for op in conv_ops if op.inputs[0].needs_grad)
    # every launch of conv3x3_tc_kernel: plain fprop, fprop with the fused LayerNorm epilogue, dgrad
    t_conv = (prof.get("fwd:conv:tc", 0.0) + prof.get("fwd:conv+ln:tc", 0.0) + prof.get("bwd:dgrad:tc", 0.0)
              + prof.get("bwd:dgrad+ln:tc", 0.0))      # (dgrad launches with the fused LayerNorm backward epilogue)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # B200_PROFILING.md fallbacks when the driver-written file is absent
    peak_burst = peaks.get("bf16_tflops", 1650.0)
    peak_sus = peaks.get("bf16_tflops_sustained", 1400.0)
    ach = (flops_fprop + flops_dgrad) / (t_conv / 1000.0) / 1e12 if t_conv > 0 else 0.0
    t_wg = prof.get("bwd:wgrad:tc", 0.0)
    ach_wg = flops_fprop / (t_wg / 1000.0) / 1e12 if t_wg > 0 else 0.0
    n_launch = sum(1 for op in conv_ops) + sum(1 for op in conv_ops if op.inputs[0].needs_grad)
    # algorithmic HBM bytes of those launches: input read once + output written once (bf16), x2 outputs with fused LN
    def _act_bytes(v):
        return 2.0 * batch * v.h * v.w * v.c
    # the launches that are conv3x3_tc_kernel proper (images larger than 4x4; the deep levels run the split-K
    # conv_gemm_kernel, whose traffic is weights) -- the population of the ncu `traffic` number
    win_ops = [op for op in conv_ops if max(op.output.h, op.output.w) > 4]
    bytes_win = sum(_act_bytes(op.inputs[0]) + _act_bytes(op.output) * (2 if getattr(op, "fused_into_ln", False) else 1)
                    for op in win_ops)
    bytes_win += sum(_act_bytes(op.inputs[0]) + _act_bytes(op.output) for op in win_ops if op.inputs[0].needs_grad)
    n_win = len(win_ops) + sum(1 for op in win_ops if op.inputs[0].needs_grad)
    traffic, traffic_file = None, None   # DRAM bytes per launch of the same kernel, from the committed `ncu --set full` capture
    for name in ("r02_conv_traffic.json", "r01_v9_conv_traffic.json"):
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", name)))["dram_bytes_per_launch_mean"]
            traffic_file = name
            break
        except Exception:
            pass
    at_boost = bool(clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.97 * clocks["sm_max_mhz"])
    roofline = {
        "kernel": "conv3x3_tc_kernel (tcgen05 implicit GEMM; every fprop / fprop+LayerNorm / dgrad launch of one step, incl. the split-K conv_gemm_kernel launches of the deep levels)",
        "bound": "tensor", "achieved": ach, "peak": peak_burst, "unit": "TFLOP/s", "frac": ach / peak_burst,
        "peak_source": ("MEASURED_PEAKS.json bf16_tflops (burst, of measured): each launch is timed alone with CUDA events"
                        if peaks else "fallback 1650 burst (of fallback, B200_PROFILING.md)"),
        "frac_of_sustained": ach / peak_sus, "peak_sustained": peak_sus, "sm_clock_at_boost_during_value": at_boost,
        "traffic": traffic, "traffic_static_from_profile": True,
        "traffic_source": f"profiles/{traffic_file} (ncu dram__bytes_read+write, mean over the conv3x3_tc_kernel launches of one step; "
                          "a committed capture, NOT measured in this run)" if traffic_file else None,
        "algorithmic_bytes_per_launch": bytes_win / max(n_win, 1),
        "launches_per_step": n_launch, "avg_launch_ms": t_conv / max(n_launch, 1),
        "wgrad_kernel": {"achieved": ach_wg, "frac": ach_wg / peak_burst, "frac_of_sustained": ach_wg / peak_sus,
                         "ms_per_step": t_wg},
    }
    # algorithmic FLOPs of one step: 3 x the forward FLOPs of every Conv2D of the model's own graph
    step_tflop = 3.0 * sum(conv_flops(op, batch) for op in plan.ops if op.kind == "conv") / 1e12
    if sustained is not None:
        sustained["step_tflops_algorithmic"] = step_tflop / (sustained["ms_per_step"] / 1000.0)
        sustained["step_frac_of_sustained_peak"] = sustained["step_tflops_algorithmic"] / peak_sus

    # ---- CPU baseline: the oracle, bounded sample; rank 0 at N = 1 only (the other ranks of a multi-GPU run would idle) --
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        sb = 8                                   # ~10 s of host work on the box's cores (0.7 s per batch-8 step at 16 cores)
        times, cores = cpu_reference_steps(depth, scale, patch, sb, 12, 1)
        cpu = {"value": sb * len(times) / sum(times), "unit": "patches/s", "cores": cores, "kind": "port",
               "sample": f"{len(times)} steps of batch {sb} of the same workload shapes, torch-CPU restatement of the Keras graph"}
    elif world > 1:
        cpu = {"value": None, "unit": "patches/s", "cores": None, "kind": "port",
               "sample": "not run at N > 1 (rank 0 at N = 1 only; see the N = 1 line or `--impl reference`)"}

    line = {
        "metric": "U-Net train patches/s", "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": desc, "depth": depth, "scale": scale, "patch": patch, "per_gpu_batch": batch,
                   "global_batch": batch * world, "loss": "charbonnier", "optimizer": "adam",
                   "parallelism": f"dp{world}", "streams": "dgrad / norm backward on the main stream, wgrad forked to a second stream inside the captured step" if getattr(model, "overlap_wgrad", False) else "single stream", "l2": "per-step working set (>2 GB of activations) far exceeds the 126 MB L2",
                   "params": model.count_params(), "step_tflop_algorithmic": step_tflop,
                   "step_frac_of_conv_roofline": step_tflop / (ms_per_step / 1000.0) / peak_burst,
                   "step_frac_of_sustained_peak": step_tflop / (ms_per_step / 1000.0) / peak_sus},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "patches/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "sustained": sustained,
        "breakdown_ms": {k: round(v, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1])},
        "eager_step_ms": step_ms_eager,
    }
    if c3 is not None:
        line["c3"] = c3
    print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        model.release_graphs()
        dist.barrier()
        dist.destroy_process_group()


def _profile(model, plan, st):
    prof = profile_step(model, plan, st)
    for _ in range(2):
        p2 = profile_step(model, plan, st)
        prof = {k: min(v, p2[k]) for k, v in prof.items()}
    return prof


def _aux_line(metric, unit, value, ms, world, args, desc, extra, clocks, e2e, launches):
    return {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict({"workload": desc, "parallelism": f"replicas x{world}" if world > 1 else "1 GPU"}, **extra),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": None, "cpu_baseline": None}


def run_c4(args, world, rank, local):
    """BASELINE configs[3]: segmentation U-Net (build_unet, base 32, depth 4), 256x256 RGB, 21 classes, softmax-CE
    training, batch 32 -- images/s.  N > 1: data parallel like the SR net."""
    import numpy as np
    import torch
    from b200unet import builders as B, ops
    from b200unet.keras import clear_session, mixed_precision
    from b200unet.keras.losses import CategoricalCrossentropy
    from b200unet.keras.optimizers import Adam
    _, _, size, batch, desc = CONFIGS["c4"]
    classes = 21
    clear_session(); mixed_precision.set_global_policy("mixed_bfloat16")
    model = B.build_unet(size, num_classes=classes, base_channels=32, depth=4)
    model.compile(optimizer=Adam(learning_rate=1e-4), loss=CategoricalCrossentropy())
    if world > 1:
        model.distribute()
    rng = np.random.default_rng(1234 + rank)
    x = torch.from_numpy(rng.random((batch, size, size, 3), dtype=np.float32)).pin_memory()
    y = torch.from_numpy(rng.integers(0, classes, (batch, size, size)).astype(np.int32)).pin_memory()
    ops.launch_count(reset=True)
    model.train_on_batch(x, y)
    e = model._train_state(batch)
    launches = ops.launch_count() // 2
    total_ms, clocks = _timed_steps(lambda: model._run_step(e), args.steps, args.warmup, world, local)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model.train_on_batch(x, y)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    fwd_gflop = 24.23
    line = _aux_line("segmentation U-Net train images/s", "images/s", world * batch * args.steps / (total_ms / 1000.0),
                     total_ms / args.steps, world, args, desc,
                     {"classes": classes, "base_channels": 32, "params": model.count_params(),
                      "step_tflops_algorithmic": 3 * fwd_gflop * batch / (total_ms / args.steps)},
                     clocks, {"value": world * batch * args.steps / e2e_s, "unit": "images/s",
                              "h2d_bytes_per_step": x.numel() * 4 + y.numel() * 4, "d2h_bytes_per_step": 4},
                     launches * args.steps)
    if world > 1:
        model.release_graphs()
    return line


def run_c5(args, world, rank, local):
    """BASELINE configs[4]: depth-5 SR U-Net inference on 1024x1024 images, batch 16 per GPU, forward only -- images/s.
    N > 1: N independent replicas (inference has no exchange step)."""
    import numpy as np
    import torch
    from b200unet import builders as B, ops
    from b200unet.keras import clear_session, mixed_precision
    depth, scale, size, batch, desc = CONFIGS["c5"]
    clear_session(); mixed_precision.set_global_policy("mixed_bfloat16")
    model, _ = B.build_super_resolution_unet(scale, depth_override=depth, input_size=size)
    rng = np.random.default_rng(1234 + rank)
    x = torch.from_numpy(rng.random((batch, size, size, 3), dtype=np.float32)).pin_memory()
    ops.launch_count(reset=True)
    y = model(x)
    launches = ops.launch_count() // 2
    e = model._eval_state(batch, False)
    total_ms, clocks = _timed_steps(lambda: (e["graph"].replay() if e["graph"] is not None else e["body"]()),
                                    args.steps, args.warmup, world, local)
    host = torch.empty(y.shape, dtype=torch.float32).pin_memory()
    t0 = time.perf_counter()
    for _ in range(max(args.steps // 5, 2)):
        host.copy_(model(x))                       # H2D of the images, forward, D2H of the restored images
    torch.cuda.synchronize()
    n_e2e = max(args.steps // 5, 2)
    e2e_s = time.perf_counter() - t0
    fwd_gflop = 789.46
    return _aux_line("SR U-Net inference images/s", "images/s", world * batch * args.steps / (total_ms / 1000.0),
                     total_ms / args.steps, world, args, desc,
                     {"depth": depth, "scale": scale, "tflops_algorithmic": fwd_gflop * batch / (total_ms / args.steps),
                      "output_finite": bool(torch.isfinite(y.float()).all().item()),
                      "mem_gb": torch.cuda.max_memory_allocated() / 2**30},
                     clocks, {"value": world * batch * n_e2e / e2e_s, "unit": "images/s",
                              "h2d_bytes_per_step": x.numel() * 4, "d2h_bytes_per_step": host.numel() * 4},
                     launches * args.steps)


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
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 3 s sustained block")
    ap.add_argument("--no-c3", action="store_true", help="N > 1: skip the configs[2] (depth 5, global batch 512) sub-record")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
