"""BASELINE configs 4 and 5 at their full sizes (the bench line covers config 2; these are run once per round as a
scale check and reported in DESIGN.md):
  C4  segmentation U-Net depth 4 (build_unet, base 32), 256x256 RGB, 21 classes, softmax-CE training, batch 32
  C5  SR U-Net depth 5, scale 0.25, 1024x1024 full images, batch 16, inference (forward only)
Usage (GPU box):  python tools/config_sweep.py [c4] [c5]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200unet import builders as B  # noqa: E402
from b200unet.keras import clear_session, mixed_precision  # noqa: E402
from b200unet.keras.losses import CategoricalCrossentropy  # noqa: E402
from b200unet.keras.optimizers import Adam  # noqa: E402


def timed(fn, iters):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def c4(base=32):
    clear_session(); mixed_precision.set_global_policy("mixed_bfloat16")
    batch, size, classes = 32, 256, 21
    model = B.build_unet(size, num_classes=classes, base_channels=base, depth=4)
    model.compile(optimizer=Adam(learning_rate=1e-4), loss=CategoricalCrossentropy())
    rng = np.random.default_rng(1234)
    x = torch.from_numpy(rng.random((batch, size, size, 3), dtype=np.float32)).pin_memory()
    y = torch.from_numpy(rng.integers(0, classes, (batch, size, size)).astype(np.int32)).pin_memory()
    l0 = model.train_on_batch(x, y)["loss"]
    e = model._train_state(batch)
    ms = timed(lambda: model._run_step(e), 10)
    l1 = model.train_on_batch(x, y)["loss"]
    prof = {}
    try:
        plan, st = e["plan"], e["state"]
        rec = []
        def tm(tag, fn):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); rec.append((tag, a, b))
        model.G.zero_()
        for tag, f in zip(plan.step_tags, plan.steps):
            tm("fwd:" + tag, f)
        tm("loss", lambda: model.loss.launch(plan, st, grad_scale=1.0))
        for i, f in enumerate(plan.bwd_steps):
            tm("bwd:" + plan.bwd_tags[i], f)
        torch.cuda.synchronize()
        for tag, a, b in rec:
            prof[tag] = round(prof.get(tag, 0.0) + a.elapsed_time(b), 3)
    except Exception as ex:
        prof = {"error": str(ex)}
    return {"config": f"C4: seg U-Net depth 4 (build_unet base {base}), 256x256, 21 classes, softmax-CE, batch 32, bf16",
            "breakdown_ms": dict(sorted(prof.items(), key=lambda kv: -kv[1] if isinstance(kv[1], float) else 0)),
            "ms_per_step": ms, "images_per_s": batch / ms * 1e3, "loss_first": l0, "loss_after": l1,
            "params": model.count_params()}


def c5():
    clear_session(); mixed_precision.set_global_policy("mixed_bfloat16")
    batch, size = 16, 1024
    model, info = B.build_super_resolution_unet(0.25, depth_override=5, input_size=size)
    rng = np.random.default_rng(1234)
    x = torch.from_numpy(rng.random((batch, size, size, 3), dtype=np.float32)).pin_memory()
    y = model(x)
    e = model._eval_state(batch, False)
    ms = timed(lambda: (e["graph"].replay() if e["graph"] is not None else e["body"]()), 5)
    fwd_gflop = 789.46 * batch
    return {"config": "C5: SR U-Net depth 5, scale 0.25, 1024x1024, batch 16, forward only, bf16",
            "ms_per_batch": ms, "images_per_s": batch / ms * 1e3, "tflops_algorithmic": fwd_gflop / ms,
            "output_finite": bool(torch.isfinite(y.float()).all().item()),
            "mem_gb": torch.cuda.max_memory_allocated() / 2**30}


if __name__ == "__main__":
    which = sys.argv[1:] or ["c4", "c5"]
    for w in which:
        print(json.dumps({"c4": c4, "c4b64": lambda: c4(64), "c5": c5}[w]()), flush=True)
