"""Where does the N=64 conv kernel spend its time?  Runs the 64->64 layer at C2's full resolution with the
epilogue progressively disabled (B200_CONV_DEBUG) and the MMA-rate probe in the convolution's access pattern."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200unet.ops as ops  # noqa: E402
from b200unet._ffi import check  # noqa: E402


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


B, S = 64, 128
for ci, co in ((64, 64), (128, 64), (64, 128)):
    x = torch.randn(B, S, S, ci, device="cuda").bfloat16()
    w = ops.ConvFilter((torch.randn(3, 3, ci, co, device="cuda") * 0.05).bfloat16())
    y = torch.empty(B, S, S, co, device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(co, device="cuda")
    g = torch.ones(co, device="cuda"); be = torch.zeros(co, device="cuda")
    z = torch.empty_like(y); mean = torch.empty(B * S * S, device="cuda"); rstd = torch.empty_like(mean)
    fl = 2.0 * B * S * S * ci * co * 9
    for dbg in ("0", "2", "1"):
        os.environ["B200_CONV_DEBUG"] = dbg
        t = timeit(lambda: ops.conv2d_fprop(x, w, bias, y, 1))
        print(f"{ci}->{co} fprop debug={dbg}: {t:7.1f} us  {fl / t / 1e6:7.1f} TFLOP/s", flush=True)
    os.environ["B200_CONV_DEBUG"] = "0"
    t = timeit(lambda: ops.conv2d_ln_fprop(x, w, bias, g, be, 1e-3, True, z, y, mean, rstd))
    print(f"{ci}->{co} conv+LN fused: {t:7.1f} us  {fl / t / 1e6:7.1f} TFLOP/s", flush=True)
    t = timeit(lambda: ops.conv2d_ln_fprop(x, w, bias, g, be, 1e-3, True, None, y, mean, rstd))
    print(f"{ci}->{co} conv+LN fused (no z): {t:7.1f} us", flush=True)

L = ops.lib()
sms = torch.cuda.get_device_properties(0).multi_processor_count
for n, stride in ((64, 1024), (64, 1280), (128, 1024), (128, 1280)):
    cyc = torch.zeros(sms, dtype=torch.int64, device="cuda")
    iters = 1800
    check(L.b200_debug_umma_rate(n, iters, stride, cyc.data_ptr(), sms, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    print(f"mma rate N={n} sbo={stride}: {cyc.float().mean().item() / (iters * 4):.2f} cycles/MMA (ideal {128 * n / 256})")
