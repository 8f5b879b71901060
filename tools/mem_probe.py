"""One launch of each HBM-bound kernel variant at C2's full-resolution shapes -- the target of an
`ncu --set full` capture (tools/mem_table.py times the same kernels without the profiler)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200unet.ops as ops  # noqa: E402

B, S, C = 64, 128, 64
dev = "cuda"
bf = lambda *s: torch.randn(*s, device=dev).bfloat16()
dy, z, dz = bf(B, S, S, C), bf(B, S, S, C), bf(B, S, S, C)
npix = B * S * S
mean = torch.zeros(npix, device=dev); rstd = torch.ones(npix, device=dev)
g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
dg, db, dbias = torch.zeros(C, device=dev), torch.zeros(C, device=dev), torch.zeros(C, device=dev)
y = torch.empty_like(z)
ops.layernorm_bwd(dy, z, mean, rstd, g, b, True, dz, dg, db, dbias)
ops.layernorm_fwd(z, g, b, 1e-3, True, y, mean, rstd)
ph = ops.ResamplePlan(128, 32, True, dev)
x, ys = bf(B, 128, 128, 64), bf(B, 32, 32, 64)
ops.resample2d(x, ys, ph, ph)
ops.resample2d_bwd(ys, x, ph, ph, True)
pu = ops.ResamplePlan(32, 128, True, dev)
xs, yb = bf(B, 32, 32, 128), bf(B, 128, 128, 128)
ops.resample2d(xs, yb, pu, pu)
ops.resample2d_bwd(yb, xs, pu, pu, False)
torch.cuda.synchronize()
print("ok")
