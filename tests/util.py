"""Shared helpers for the parity tests."""
import numpy as np
import torch


def relerr(a, b):
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def rand(shape, seed, dtype=torch.float32, scale=1.0, device="cuda"):
    g = torch.Generator().manual_seed(seed)
    t = (torch.rand(shape, generator=g) * 2 - 1) * scale
    return t.to(dtype).to(device)


def f32(t):
    return t.detach().float().cpu()


TOL = {torch.float32: 2e-5, torch.bfloat16: 1e-2}


def conv_ws(ops, x, filt, dy=None):
    """Per-call split-K scratch of a convolution (None when the shapes do not take that path), NaN-filled."""
    need = ops.conv2d_workspace(x, filt, False)
    if dy is not None:
        need = max(need, ops.conv2d_workspace(dy, filt, True))
    ws = ops.new_workspace(need, "cuda")
    if ws is not None:
        ws.fill_(float("nan"))
    return ws
