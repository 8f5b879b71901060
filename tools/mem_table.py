"""Per-kernel table of the HBM-bound kernels at C2's full-resolution shapes (CUDA events, eager launches):
microseconds and achieved GB/s against the ALGORITHMIC bytes (each input read once, each output written once).
Run on the GPU box:  python tools/mem_table.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200unet.ops as ops  # noqa: E402

PEAK = 6548.8
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def row(name, us, nbytes):
    gbs = nbytes / us / 1e3
    print(f"{name:44s} {us:8.1f} us  {nbytes / 1e6:8.1f} MB  {gbs:7.0f} GB/s  {gbs / PEAK * 100:5.1f}% of HBM peak", flush=True)


def bf(*shape):
    return torch.randn(*shape, device="cuda").bfloat16()


def main():
    B, P = 64, 128
    dev = "cuda"
    for C, S in ((64, 128), (128, 32), (256, 8), (512, 2), (1024, 1)):
        npix = B * S * S
        dy, z, dz = bf(B, S, S, C), bf(B, S, S, C), bf(B, S, S, C)
        y = torch.empty_like(z)
        mean = torch.zeros(npix, device=dev); rstd = torch.ones(npix, device=dev)
        g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
        dg, db, dbias = torch.zeros(C, device=dev), torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        t = timeit(lambda: ops.layernorm_fwd(z, g, b, 1e-3, True, y, mean, rstd))
        row(f"ln_fwd  C={C} {S}x{S}", t, 2 * z.numel() * 2)
        t = timeit(lambda: ops.layernorm_bwd(dy, z, mean, rstd, g, b, True, dz, dg, db, dbias))
        row(f"ln_bwd  C={C} {S}x{S}", t, 3 * z.numel() * 2)
        t = timeit(lambda: ops.bias_act_bwd(dy, z, 1, dy, dbias))
        row(f"bias_relu_bwd (in place) C={C} {S}x{S}", t, 3 * z.numel() * 2)
    # resize: encoder down-sampling (antialiased, 8 taps/axis) and decoder up-sampling (2 taps/axis), fwd + bwd
    for C, big, small in ((64, 128, 32), (128, 32, 8)):
        ph = ops.ResamplePlan(big, small, True, dev)
        x, ys = bf(B, big, big, C), bf(B, small, small, C)
        t = timeit(lambda: ops.resample2d(x, ys, ph, ph))
        row(f"resize down {big}->{small} C={C} fwd", t, (x.numel() + ys.numel()) * 2)
        t = timeit(lambda: ops.resample2d_bwd(ys, x, ph, ph, True))
        row(f"resize down {big}->{small} C={C} bwd (accumulate)", t, (2 * x.numel() + ys.numel()) * 2)
        pu = ops.ResamplePlan(small, big, True, dev)
        xs, yb = bf(B, small, small, 2 * C), bf(B, big, big, 2 * C)
        t = timeit(lambda: ops.resample2d(xs, yb, pu, pu))
        row(f"resize up {small}->{big} C={2 * C} fwd", t, (xs.numel() + yb.numel()) * 2)
        t = timeit(lambda: ops.resample2d_bwd(yb, xs, pu, pu, False))
        row(f"resize up {small}->{big} C={2 * C} bwd", t, (xs.numel() + yb.numel()) * 2)
    # stem 3->64 and 1x1 head 64->3 at full resolution
    x3 = bf(B, P, P, 3)
    xcol = torch.empty(B, P, P, 64, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: ops.im2col3x3(x3, xcol))
    row("stem im2col 3 -> 64 (27 live)", t, (x3.numel() + xcol.numel()) * 2)
    y64, dy64 = bf(B, P, P, 64), bf(B, P, P, 64)
    f = ops.ConvFilter((torch.randn(3, 3, 3, 64, device=dev) * 0.1).bfloat16())
    bias = torch.zeros(64, device=dev)
    t = timeit(lambda: ops.conv2d_fprop(x3, f, bias, y64, 0))
    row("stem fprop 3->64", t, (x3.numel() + y64.numel()) * 2)
    dw = torch.zeros(3 * 3 * 3 * 64, device=dev)
    t = timeit(lambda: ops.conv2d_wgrad(x3, dy64, 3, 3, dw))
    row("stem wgrad 3->64", t, (x3.numel() + y64.numel()) * 2)
    fh = ops.ConvFilter((torch.randn(1, 1, 64, 3, device=dev) * 0.1).bfloat16())
    y3, dy3 = bf(B, P, P, 3), bf(B, P, P, 3)
    b3 = torch.zeros(3, device=dev)
    t = timeit(lambda: ops.conv2d_fprop(y64, fh, b3, y3, 0))
    row("head fprop 64->3", t, (y3.numel() + y64.numel()) * 2)
    t = timeit(lambda: ops.conv2d_dgrad(dy3, fh, dy64, False))
    row("head dgrad 3->64", t, (y3.numel() + y64.numel()) * 2)
    dwh = torch.zeros(64 * 3, device=dev)
    t = timeit(lambda: ops.conv2d_wgrad(y64, dy3, 1, 1, dwh))
    row("head wgrad", t, (y3.numel() + y64.numel()) * 2)
    t = timeit(lambda: ops.bias_act_bwd(dy3, y3, 0, dy3, b3))
    row("head bias_bwd C=3", t, 2 * y3.numel() * 2)


if __name__ == "__main__":
    main()
