"""Evaluation metrics of the reference's eval loops (not on the training hot path; torch ops).

BT.601 luma, PSNR, SSIM and MS-SSIM with tf.image semantics
(/root/reference/Super_resolution/code/train_adaptive_unet.py:144-157, 673-721): 11x11 Gaussian
window (sigma 1.5), K1 0.01, K2 0.03, "valid" filtering, per-image mean; MS-SSIM over 5 scales with
the standard weights and 2x2 average pooling between scales.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_MS_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def rgb_to_luma_bt601(image: torch.Tensor) -> torch.Tensor:
    """RGB in [0,1], (N,H,W,3) or (H,W,3) -> BT.601 luma in [0,1] with a trailing singleton channel."""
    image = image.float()
    coeffs = torch.tensor([65.481, 128.553, 24.966], device=image.device)
    y = (image * coeffs).sum(dim=-1, keepdim=True) + 16.0
    return (y / 255.0).clamp(0.0, 1.0)


def psnr(a: torch.Tensor, b: torch.Tensor, max_val: float = 1.0) -> torch.Tensor:
    mse = ((a.float() - b.float()) ** 2).mean(dim=(1, 2, 3))
    return 10.0 * torch.log10(max_val * max_val / mse)


def _gauss(size=11, sigma=1.5, device="cpu"):
    x = torch.arange(size, dtype=torch.float32, device=device) - (size - 1) / 2.0
    g = torch.exp(-(x * x) / (2 * sigma * sigma))
    g = g / g.sum()
    return (g[:, None] * g[None, :])[None, None]


def _ssim_cs(a, b, max_val=1.0):
    """a, b: (N,H,W,C).  Returns per-image (ssim, cs)."""
    a, b = a.float().permute(0, 3, 1, 2), b.float().permute(0, 3, 1, 2)
    c = a.shape[1]
    k = _gauss(device=a.device).expand(c, 1, 11, 11)
    f = lambda t: F.conv2d(t, k, groups=c)
    mu_a, mu_b = f(a), f(b)
    var_a, var_b, cov = f(a * a) - mu_a * mu_a, f(b * b) - mu_b * mu_b, f(a * b) - mu_a * mu_b
    c1, c2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    cs = (2 * cov + c2) / (var_a + var_b + c2)
    lum = (2 * mu_a * mu_b + c1) / (mu_a * mu_a + mu_b * mu_b + c1)
    return (lum * cs).mean(dim=(1, 2, 3)), cs.mean(dim=(1, 2, 3))


def ssim(a, b, max_val=1.0):
    return _ssim_cs(a, b, max_val)[0]


def ssim_multiscale(a, b, max_val=1.0):
    """tf.image.ssim_multiscale; NaN when the image is too small for 5 scales (< 176 px), where TF raises."""
    if min(a.shape[1], a.shape[2]) < 11 * 2 ** 4:
        return torch.full((a.shape[0],), float("nan"), device=a.device)
    vals = []
    for i, w in enumerate(_MS_WEIGHTS):
        s, cs = _ssim_cs(a, b, max_val)
        vals.append(torch.relu(s if i == len(_MS_WEIGHTS) - 1 else cs) ** w)
        if i < len(_MS_WEIGHTS) - 1:
            a = F.avg_pool2d(a.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
            b = F.avg_pool2d(b.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    return torch.stack(vals, dim=0).prod(dim=0)
