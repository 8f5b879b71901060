"""Launch each hot kernel of config C2 at its full-resolution shape a few times, for `ncu --set full`.
    ncu --set full --clock-control none --import-source on -o gpurun_out/prof python tools/profile_kernels.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200unet.ops as ops  # noqa: E402

B, S = 64, 128
dt = torch.bfloat16
dev = "cuda"


def t(*shape, dtype=dt):
    return torch.randn(*shape, device=dev).to(dtype)


x64, x128 = t(B, S, S, 64), t(B, S, S, 128)
y64, dy64 = torch.empty_like(x64), t(B, S, S, 64)
w66, w126 = ops.ConvFilter((t(3, 3, 64, 64) * 0.05)), ops.ConvFilter((t(3, 3, 128, 64) * 0.05))
bias = torch.zeros(64, device=dev)
dw66 = torch.empty(3, 3, 64, 64, device=dev); dw126 = torch.empty(3, 3, 128, 64, device=dev)
ws = torch.empty(max(ops.conv2d_wgrad_workspace(x128, dy64, 3, 3), 16) // 4, device=dev)
dx64, dx128 = torch.empty_like(x64), torch.empty_like(x128)
g, b_ = torch.ones(64, device=dev), torch.zeros(64, device=dev)
mean = torch.empty(B * S * S, device=dev); rstd = torch.empty_like(mean)
dg, db, dbias = (torch.zeros(64, device=dev) for _ in range(3))
small = t(B, 32, 32, 64); up = torch.empty(B, S, S, 128, device=dev, dtype=dt); s128 = t(B, 32, 32, 128)
pd = ops.ResamplePlan(S, 32, True, dev); pu = ops.ResamplePlan(32, S, True, dev)
x3 = t(B, S, S, 3); w3 = ops.ConvFilter(t(3, 3, 3, 64) * 0.2, packed=False); dw3 = torch.empty(3, 3, 3, 64, device=dev)
wh = ops.ConvFilter(t(1, 1, 64, 3) * 0.2, packed=False); y3 = torch.empty(B, S, S, 3, device=dev, dtype=dt)
dwh = torch.empty(1, 1, 64, 3, device=dev)
n = 34_599_363
p, gr, m, v = (torch.zeros(n, device=dev) for _ in range(4)); sh = torch.empty(n, device=dev, dtype=dt)
hyper = torch.tensor([1e-4, 0.9, 0.999, 1e-7, 0.1, 0.001], device=dev); step = torch.ones(1, dtype=torch.int32, device=dev)

for rep in range(2):
    ops.conv2d_fprop(x64, w66, bias, y64, 0)
    ops.conv2d_fprop(x128, w126, bias, y64, 1)
    ops.conv2d_dgrad(dy64, w66, dx64)
    ops.conv2d_dgrad(dy64, w126, dx128)
    ops.conv2d_wgrad(x64, dy64, 3, 3, dw66, ws)
    ops.conv2d_wgrad(x128, dy64, 3, 3, dw126, ws)
    ops.layernorm_fwd(x64, g, b_, 1e-3, True, y64, mean, rstd)
    ops.layernorm_bwd(dy64, x64, mean, rstd, g, b_, True, dx64, dg, db, dbias)
    ops.resample2d(x64, small, pd, pd)
    ops.resample2d(s128, up, pu, pu)
    ops.resample2d_bwd(small, dx64, pd, pd)
    ops.resample2d_bwd(up, s128, pu, pu)
    ops.bias_act_bwd(dy64, y64, 1, dy64, dbias)
    ops.conv2d_fprop(x3, w3, bias, y64, 0)
    ops.conv2d_wgrad(x3, dy64, 3, 3, dw3, None)
    ops.conv2d_fprop(x64, wh, None, y3, 0)
    ops.conv2d_dgrad(y3, wh, dx64)
    ops.conv2d_wgrad(x64, y3, 1, 1, dwh, None)
    ops.adam_step(p, gr, m, v, hyper, step, sh)
torch.cuda.synchronize()
print("ok")
