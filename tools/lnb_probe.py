"""Fused dgrad + LayerNorm backward at C2's full-resolution shape (64 x 128 x 128 x 64, K = 64): time against the
unfused pair; also the command ncu profiles."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200unet.ops as ops
from conv_table import timeit_graph

B, S, C = 64, 128, 64
z = torch.randn(B, S, S, C, device="cuda").bfloat16()
dy = torch.randn(B, S, S, C, device="cuda").bfloat16()
w = ops.ConvFilter((torch.randn(3, 3, C, C, device="cuda") * 0.05).bfloat16())
g = torch.ones(C, device="cuda"); be = torch.zeros(C, device="cuda")
y = torch.empty_like(z); mean = torch.empty(B * S * S, device="cuda"); rstd = torch.empty_like(mean)
ops.layernorm_fwd(z, g, be, 1e-3, True, y, mean, rstd)
dz = torch.empty_like(z); dx = torch.empty_like(z)
dg = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda"); dbias = torch.zeros(C, device="cuda")
fused = lambda: ops.conv2d_dgrad_ln_bwd(dy, w, z, mean, rstd, g, be, True, dz, dg, db, dbias)
dgrad = lambda: ops.conv2d_dgrad(dy, w, dx, False)
lnb = lambda: ops.layernorm_bwd(dx, z, mean, rstd, g, be, True, dz, dg, db, dbias)
for _ in range(3):
    fused(); dgrad(); lnb()
torch.cuda.synchronize()
if os.environ.get("PROBE_TIMING", "1") == "1":
    print(f"fused {timeit_graph(fused) * 1e3:.1f} us | dgrad {timeit_graph(dgrad) * 1e3:.1f} us + LN bwd {timeit_graph(lnb) * 1e3:.1f} us")
