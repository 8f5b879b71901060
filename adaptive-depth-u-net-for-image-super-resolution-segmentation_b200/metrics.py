"""Evaluation metrics of the reference's eval loops on the GPU (kernels of ``csrc/metrics.cu``).

BT.601 luma, border shave, MSE(Y), PSNR(Y), SSIM(Y) and MS-SSIM(Y) with tf.image semantics
(/root/reference/Super_resolution/code/train_adaptive_unet.py:144-157, 673-721;
evaluate_model.py:94-163).  One fused kernel turns the RGB prediction and target into the two shaved luma
planes and the per-image squared error; the SSIM kernel filters a, b, a^2, b^2, ab with the separable 11x11
Gaussian in shared memory and reduces the SSIM / contrast-structure maps per image; MS-SSIM chains it over
five scales with the 2x2 pooling kernel.  Only the last step (a few floats per image: log10, the five-factor
product) runs on the host.  There is no CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import _ffi, ops

_MS_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)
_WIN = 11


def _cuda(t, dtype=None) -> torch.Tensor:
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(t)
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise ops._ffi.B200Error("b200unet.metrics needs a CUDA device: there is no CPU fallback")
        t = t.cuda()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def luma_planes(pred_rgb, hr_rgb, shave: int = 0):
    """(pred_y, hr_y, sse): shaved BT.601 luma planes [n,h',w'] (fp32; prediction clipped to [0,1] first) and
    the per-image sum of squared luma differences."""
    pred = _cuda(pred_rgb)
    if pred.dtype not in (torch.float32, torch.bfloat16):
        pred = pred.float()
    hr = _cuda(hr_rgb, torch.float32)
    n, h, w, _ = pred.shape
    oh, ow = h - 2 * shave, w - 2 * shave
    py = torch.empty((n, oh, ow), dtype=torch.float32, device=pred.device)
    hy = torch.empty_like(py)
    sse = torch.empty((n,), dtype=torch.float32, device=pred.device)
    ops.luma_pair(pred, hr, shave, py, hy, sse)
    return py, hy, sse


def rgb_to_luma_bt601(image) -> torch.Tensor:
    """RGB in [0,1], (N,H,W,3) or (H,W,3) -> BT.601 luma in [0,1] with a trailing singleton channel.

    Note: the input is clipped to [0,1] first (the eval loops clip the prediction before the call; an HR image
    is in range already)."""
    img = _cuda(image, torch.float32)
    single = img.dim() == 3
    if single:
        img = img[None]
    _, y, _ = luma_planes(img, img, 0)
    y = y[..., None]
    return y[0] if single else y


def _images(t) -> torch.Tensor:
    """(N,H,W) / (N,H,W,C) -> contiguous fp32 CUDA [N,H,W,C]."""
    t = _cuda(t, torch.float32)
    if t.dim() == 3:
        t = t[..., None]
    return t.contiguous()


def _ssim_cs(a: torch.Tensor, b: torch.Tensor, max_val: float) -> np.ndarray:
    """Per-image, per-channel means of the SSIM map and of its contrast-structure factor: float64 numpy [n,ch,2]."""
    n, h, w, ch = a.shape
    out = torch.empty((n * ch, 2), dtype=torch.float32, device=a.device)
    ops.ssim_planes(a, b, out, max_val)
    return out.cpu().numpy().astype(np.float64).reshape(n, ch, 2) / float((h - _WIN + 1) * (w - _WIN + 1))


def ssim(a, b, max_val: float = 1.0) -> np.ndarray:
    """tf.image.ssim per image (float32 numpy [n]): per-channel SSIM averaged over the channels."""
    return _ssim_cs(_images(a), _images(b), max_val)[:, :, 0].mean(axis=1).astype(np.float32)


def ssim_multiscale(a, b, max_val: float = 1.0) -> np.ndarray:
    """tf.image.ssim_multiscale per image; NaN when the image is too small for 5 scales (< 176 px), where TF raises."""
    a, b = _images(a), _images(b)
    n, h, w, ch = a.shape
    if min(h, w) < _WIN * 2 ** (len(_MS_WEIGHTS) - 1):
        return np.full((n,), np.nan, dtype=np.float32)
    factors = []
    for i in range(len(_MS_WEIGHTS)):
        sc = _ssim_cs(a, b, max_val)
        factors.append(sc[:, :, 0] if i == len(_MS_WEIGHTS) - 1 else sc[:, :, 1])
        if i < len(_MS_WEIGHTS) - 1:
            h2, w2 = (a.shape[1] + 1) // 2, (a.shape[2] + 1) // 2
            a2 = torch.empty((n, h2, w2, ch), dtype=torch.float32, device=a.device)
            b2 = torch.empty_like(a2)
            ops.avgpool2_planes(a, a2)
            ops.avgpool2_planes(b, b2)
            a, b = a2, b2
    f = np.maximum(np.stack(factors, axis=2), 0.0)                       # [n,ch,5]: a few floats per image
    return np.prod(f ** np.asarray(_MS_WEIGHTS)[None, None, :], axis=2).mean(axis=1).astype(np.float32)


def psnr_rgb(a, b, max_val: float = 1.0) -> np.ndarray:
    """tf.image.psnr per image over all channels (u-net-vinillia.py:226): the SR-loss kernel's squared-error sum."""
    a, b = _images(a), _images(b)
    n = a.shape[0]
    out = np.empty((n,), dtype=np.float32)
    res = torch.empty((2,), dtype=torch.float32, device=a.device)
    ws = torch.empty((2 + n,), dtype=torch.float32, device=a.device)
    for i in range(n):      # the loss kernel reduces a whole tensor: one launch per image (evaluation only)
        ops.sr_loss(a[i:i + 1], b[i:i + 1], _ffi.LOSS_MSE, 0.0, 1.0, res, None, ws)
        out[i] = 10.0 * np.log10(max_val * max_val / max(float(res[0]), 1e-30))
    return out


def eval_luma_metrics(pred_rgb, hr_rgb, shave: int = 0) -> Dict[str, np.ndarray]:
    """One eval-loop iteration (evaluate_model.py:104-121): per-image psnr / ssim / msssim / mse of the luma planes."""
    py, hy, sse = luma_planes(pred_rgb, hr_rgb, shave)
    n, h, w = py.shape
    mse_v = sse.double().cpu().numpy() / float(h * w)
    with np.errstate(divide="ignore"):
        psnr_v = 10.0 * np.log10(1.0 / mse_v)
    ssim_v = ssim(hy, py) if min(h, w) >= _WIN else np.full((n,), np.nan, dtype=np.float32)
    return {"psnr": psnr_v.astype(np.float32), "ssim": ssim_v, "msssim": ssim_multiscale(hy, py),
            "mse": mse_v.astype(np.float32)}
