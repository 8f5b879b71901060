"""Where does the fused conv+LayerNorm epilogue spend its time?  1x1 (stem-like, K=64) and 3x3 64->64 layers at
C2's full resolution under the B200_CONV_DEBUG / B200_CONV_SLOTS switches (timing aid, results are not checked)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200unet.ops as ops  # noqa: E402


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
for ks, ci, co in ((1, 64, 64), (3, 64, 64), (3, 128, 64), (3, 64, 128)):
    x = torch.randn(B, S, S, ci, device="cuda").bfloat16()
    w = ops.ConvFilter((torch.randn(ks, ks, ci, co, device="cuda") * 0.05).bfloat16())
    y = torch.empty(B, S, S, co, device="cuda", dtype=torch.bfloat16)
    z = torch.empty_like(y)
    bias = torch.zeros(co, device="cuda"); g = torch.ones(co, device="cuda"); be = torch.zeros(co, device="cuda")
    mean = torch.empty(B * S * S, device="cuda"); rstd = torch.empty_like(mean)
    for dbg in ("0", "1", "3", "4", "6"):
        os.environ["B200_CONV_DEBUG"] = dbg
        t = timeit(lambda: ops.conv2d_ln_fprop(x, w, bias, g, be, 1e-3, True, z, y, mean, rstd))
        t2 = timeit(lambda: ops.conv2d_ln_fprop(x, w, bias, g, be, 1e-3, True, None, y, mean, rstd))
        t3 = timeit(lambda: ops.conv2d_fprop(x, w, bias, y, 1))
        print(f"k{ks} {ci}->{co} debug={dbg} slots={os.environ.get('B200_CONV_SLOTS', '4')}: conv+LN {t:6.1f} us, no z {t2:6.1f} us, plain conv+relu {t3:6.1f} us", flush=True)
