"""conv + fused LayerNorm epilogue vs the plain bias+ReLU epilogue at C2's full-resolution shapes, CTA pairs vs single CTA
(CUDA-graph timing): how much of the layer is epilogue once the MMA stream runs as pairs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200unet.ops as ops  # noqa: E402
from conv_table import timeit_graph  # noqa: E402

B, S = 64, 128
for ci, co in ((64, 64), (128, 64), (64, 128), (128, 128)):
    s = S if co == 64 else 32
    x = torch.randn(B, s, s, ci, device="cuda").bfloat16()
    w = ops.ConvFilter((torch.randn(3, 3, ci, co, device="cuda") * 0.05).bfloat16(), packed=True)
    y = torch.empty(B, s, s, co, device="cuda", dtype=torch.bfloat16)
    z = torch.empty_like(y)
    bias = torch.zeros(co, device="cuda"); g = torch.ones(co, device="cuda"); be = torch.zeros(co, device="cuda")
    mean = torch.empty(B * s * s, device="cuda"); rstd = torch.empty_like(mean)
    fl = 2.0 * B * s * s * ci * co * 9
    for name, algo in (("pairs", ops.ALGO_TCGEN05), ("1cta", ops.ALGO_TCGEN05_1CTA)):
        t = timeit_graph(lambda: ops.conv2d_ln_fprop(x, w, bias, g, be, 1e-3, True, z, y, mean, rstd, algo)) * 1e3
        t2 = timeit_graph(lambda: ops.conv2d_ln_fprop(x, w, bias, g, be, 1e-3, True, None, y, mean, rstd, algo)) * 1e3
        t3 = timeit_graph(lambda: ops.conv2d_fprop(x, w, bias, y, 1, algo)) * 1e3
        print(f"{s:4d}^2 {ci:3d}->{co:3d} [{name:5s}] conv+LN {t:6.1f} us ({fl / t / 1e6:6.0f} TF), no z {t2:6.1f} us, plain conv+relu {t3:6.1f} us "
              f"({fl / t3 / 1e6:6.0f} TF)", flush=True)
