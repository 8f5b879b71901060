#!/usr/bin/env python
"""Achieved HBM bandwidth of the patch-pipeline and eval-metric kernels (CUDA events, 20 timed launches after 5
warm-up launches, inputs larger than nothing in particular: these kernels run once per batch, cold).

    python tools/pipeline_bench.py            -> one JSON line per kernel (algorithmic bytes / time vs measured HBM peak)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200unet import metrics as MT, ops  # noqa: E402
from b200unet.shared.pipeline import DevicePatchDataset  # noqa: E402


def timed(fn, flush, reps=20, warm=5):
    for _ in range(warm):
        fn()
    evs = []
    for _ in range(reps):
        flush.zero_()                       # 256 MB write: evicts the 126 MB L2 between launches
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs]))


def main():
    peak = 6548.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak)
    except Exception:
        pass
    flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    B, P = 64, 128
    rng = np.random.default_rng(0)
    img = torch.from_numpy(rng.integers(0, 256, (1024, 1024, 3), dtype=np.uint8)).cuda()
    org = torch.from_numpy(rng.integers(0, 1024 - P, (B, 2)).astype(np.int32)).cuda()
    hr = torch.empty((B, P, P, 3), dtype=torch.float32, device="cuda")
    rows = []
    for scale in (0.25, 0.5):
        ds = DevicePatchDataset([], P, scale, B)
        s = ds.small
        small = torch.empty((B, s, s, 3), dtype=torch.float32, device="cuda")
        lr = torch.empty_like(hr)
        ms = timed(lambda: ops.patch_extract(img, org, hr), flush)
        rows.append(("patch_extract_kernel", f"B{B} P{P} uint8 source", B * P * P * 3 * (1 + 4), ms))
        ms = timed(lambda: ops.gather2d(hr, small, ds._area, ds._area, True), flush)
        rows.append(("gather2d_kernel (INTER_AREA)", f"{P}->{s}", 4 * 3 * B * (P * P + s * s), ms))
        ms = timed(lambda: ops.gather2d(small, lr, ds._cubic, ds._cubic, False), flush)
        rows.append(("gather2d_kernel (INTER_CUBIC)", f"{s}->{P}", 4 * 3 * B * (P * P + s * s), ms))
    pool = torch.empty((1024, P, P, 3), dtype=torch.float32, device="cuda")
    ids = torch.from_numpy(rng.permutation(1024)[:B].astype(np.int32)).cuda()
    ms = timed(lambda: ops.copy_rows(pool, ids, hr, None, B), flush)
    rows.append(("copy_rows_kernel", f"{B} rows of {P * P * 3} floats out of a 1024-row pool", 2 * 4 * B * P * P * 3, ms))
    n, h = 16, 256
    pred = torch.rand((n, h, h, 3), device="cuda").to(torch.bfloat16)
    tgt = torch.rand((n, h, h, 3), device="cuda")
    py = torch.empty((n, h - 8, h - 8), dtype=torch.float32, device="cuda")
    hy, sse, out = torch.empty_like(py), torch.empty(n, device="cuda"), torch.empty((n, 2), device="cuda")
    ms = timed(lambda: ops.luma_pair(pred, tgt, 4, py, hy, sse), flush)
    rows.append(("luma_pair_kernel", f"{n}x{h}x{h} bf16 pred + f32 target, shave 4", n * (h - 8) ** 2 * (6 + 12 + 8), ms))
    py4, hy4 = py[..., None], hy[..., None]
    ms = timed(lambda: ops.ssim_planes(py4, hy4, out), flush)
    rows.append(("ssim_kernel", f"{n} planes of {h - 8}x{h - 8}", 2 * 4 * n * (h - 8) ** 2, ms))
    p2 = torch.empty((n, (h - 7) // 2, (h - 7) // 2, 1), dtype=torch.float32, device="cuda")
    ms = timed(lambda: ops.avgpool2_planes(py4, p2), flush)
    rows.append(("avgpool2_planes_kernel", f"{n} planes", 4 * n * ((h - 8) ** 2 + p2.shape[1] ** 2), ms))
    for name, what, nbytes, ms in rows:
        gbps = nbytes / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": name, "case": what, "algorithmic_bytes": nbytes, "ms": round(ms, 5),
                          "achieved_gbps": round(gbps, 1), "peak_gbps": peak, "frac": round(gbps / peak, 4)}))


if __name__ == "__main__":
    main()
