"""torch-CPU restatement of the reference's evaluation metrics.

Oracle code (test infrastructure, see ``oracle/__init__.py``); the product path is ``b200unet.metrics``
(CUDA kernels of ``csrc/metrics.cu``).

BT.601 luma (``/root/reference/Super_resolution/code/train_adaptive_unet.py:144-157``) and the
``tf.image.psnr / ssim / ssim_multiscale`` calls of the eval loops (``:673-721``,
``evaluate_model.py:94-163``).  tf.image semantics (TF 2.16.1, absent here -- parity unpinned, pinned only
by the independent numpy restatement in ``tests/test_metrics_cpu.py``): 11x11 Gaussian window (sigma 1.5)
applied as a "VALID" depthwise filter, K1 0.01, K2 0.03, luminance x contrast-structure averaged per image;
MS-SSIM over 5 scales with weights (0.0448, 0.2856, 0.3001, 0.2363, 0.1333), 2x2 average pooling between
scales with odd extents padded symmetrically first, negative factors clamped to 0 before the powers.
Pass float64 tensors for a float64 evaluation.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_MS_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def rgb_to_luma_bt601(image: torch.Tensor) -> torch.Tensor:
    """RGB in [0,1], (N,H,W,3) or (H,W,3) -> BT.601 luma in [0,1] with a trailing singleton channel."""
    image = image if image.dtype == torch.float64 else image.float()
    coeffs = torch.tensor([65.481, 128.553, 24.966], device=image.device, dtype=image.dtype)
    y = (image * coeffs).sum(dim=-1, keepdim=True) + 16.0
    return (y / 255.0).clamp(0.0, 1.0)


def psnr(a: torch.Tensor, b: torch.Tensor, max_val: float = 1.0) -> torch.Tensor:
    a, b = _fp(a), _fp(b)
    mse = ((a - b) ** 2).mean(dim=(1, 2, 3))
    return 10.0 * torch.log10(max_val * max_val / mse)


def _fp(t):
    return t if t.dtype == torch.float64 else t.float()


def _gauss(size=11, sigma=1.5, device="cpu", dtype=torch.float32):
    x = torch.arange(size, dtype=dtype, device=device) - (size - 1) / 2.0
    g = torch.exp(-(x * x) / (2 * sigma * sigma))
    g = g / g.sum()
    return (g[:, None] * g[None, :])[None, None]


def _ssim_cs(a, b, max_val=1.0):
    """a, b: (N,H,W,C).  Returns per-image, per-channel (ssim, cs) of shape (N,C) (tf _ssim_per_channel)."""
    a, b = _fp(a).permute(0, 3, 1, 2), _fp(b).permute(0, 3, 1, 2)
    c = a.shape[1]
    k = _gauss(device=a.device, dtype=a.dtype).expand(c, 1, 11, 11)
    f = lambda t: F.conv2d(t, k, groups=c)
    mu_a, mu_b = f(a), f(b)
    var_a, var_b, cov = f(a * a) - mu_a * mu_a, f(b * b) - mu_b * mu_b, f(a * b) - mu_a * mu_b
    c1, c2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    cs = (2 * cov + c2) / (var_a + var_b + c2)
    lum = (2 * mu_a * mu_b + c1) / (mu_a * mu_a + mu_b * mu_b + c1)
    return (lum * cs).mean(dim=(2, 3)), cs.mean(dim=(2, 3))


def ssim(a, b, max_val=1.0):
    """tf.image.ssim: the per-channel SSIM averaged over the channels."""
    return _ssim_cs(a, b, max_val)[0].mean(dim=1)


def ssim_multiscale(a, b, max_val=1.0):
    """tf.image.ssim_multiscale; NaN when the image is too small for 5 scales (< 176 px), where TF raises."""
    if min(a.shape[1], a.shape[2]) < 11 * 2 ** 4:
        return torch.full((a.shape[0],), float("nan"), device=a.device)
    a, b = _fp(a), _fp(b)
    vals = []
    for i, w in enumerate(_MS_WEIGHTS):
        s, cs = _ssim_cs(a, b, max_val)
        vals.append(torch.relu(s if i == len(_MS_WEIGHTS) - 1 else cs) ** w)
        if i < len(_MS_WEIGHTS) - 1:
            a, b = _pool2(a), _pool2(b)
    return torch.stack(vals, dim=0).prod(dim=0).mean(dim=1)     # product over scales per channel, then channel mean


def _pool2(t):
    """2x2 average pooling of (N,H,W,C); an odd extent is first padded by repeating its last row / column
    (tf ``pad(mode="SYMMETRIC")`` in ssim_multiscale's do_pad)."""
    t = t.permute(0, 3, 1, 2)
    t = F.pad(t, (0, t.shape[3] % 2, 0, t.shape[2] % 2), mode="replicate")
    return F.avg_pool2d(t, 2).permute(0, 2, 3, 1)
