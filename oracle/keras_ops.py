"""torch-CPU restatement of the Keras 3.3.3 / TF 2.16.1 ops on the hot path.

Oracle code (test infrastructure, see ``oracle/__init__.py``; parity unpinned).
Tensors are NHWC ``torch`` tensors (fp32 or fp64) on the CPU; gradients come from
torch autograd, which is the gradient truth the CUDA backward kernels are checked
against.  Every function names the reference call site whose Keras layer it
restates (paths relative to ``/root/reference``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import resize_np


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


# --------------------------------------------------------------------------- #
# Convolutions
# --------------------------------------------------------------------------- #
def conv2d_same(x, kernel_hwio, bias=None):
    """keras ``Conv2D(nf, k, padding="same", use_bias=True)`` stride 1.

    Cross-correlation with an HWIO kernel and symmetric zero padding (odd k).
    Call sites: Super_resolution/code/train_adaptive_unet.py:202,207,259,267;
    Segmenation/code/train_adaptive_unet.py:326,329,361; unet_vinillia.py:44,49,90.
    """
    kh, kw = kernel_hwio.shape[0], kernel_hwio.shape[1]
    w = kernel_hwio.permute(3, 2, 0, 1)  # OIHW
    y = F.conv2d(_nchw(x), w, bias, stride=1, padding=(kh // 2, kw // 2))
    return _nhwc(y)


def conv2d_transpose_2x2(x, kernel_hwoi, bias=None):
    """keras ``Conv2DTranspose(nf, 2, strides=2, padding="same")``.

    Kernel layout [kh, kw, Cout, Cin]; windows do not overlap, so
    out[2i+a, 2j+b, o] = sum_c in[i, j, c] * K[a, b, o, c] + bias[o].
    Call site: Segmenation/code/unet_vinillia.py:67.
    """
    n, h, w, c = x.shape
    co = kernel_hwoi.shape[2]
    y = torch.einsum("nhwc,aboc->nhawbo", x, kernel_hwoi).reshape(n, 2 * h, 2 * w, co)
    if bias is not None:
        y = y + bias
    return y


# --------------------------------------------------------------------------- #
# Normalisation / activation
# --------------------------------------------------------------------------- #
def layer_norm(x, gamma, beta, eps=1e-3):
    """keras ``LayerNormalization(axis=-1)``: biased variance over C, eps 1e-3.

    Call sites: train_adaptive_unet.py:203,208; unet_vinillia.py:45,50.
    """
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    inv = torch.rsqrt(var + eps)
    return (x - mean) * inv * gamma + beta


def batch_norm_train(x, gamma, beta, moving_mean, moving_var, momentum=0.99, eps=1e-3):
    """keras ``BatchNormalization()`` in training mode.

    Batch moments over (N, H, W), biased variance; moving statistics updated as
    ``moving = moving * momentum + batch * (1 - momentum)`` (Keras 3 uses the
    biased variance there too).  Returns (y, new_moving_mean, new_moving_var).
    Call sites: Segmenation/code/train_adaptive_unet.py:327,330.
    """
    mean = x.mean(dim=(0, 1, 2))
    var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
    y = (x - mean) * torch.rsqrt(var + eps) * gamma + beta
    new_mean = moving_mean * momentum + mean.detach() * (1.0 - momentum)
    new_var = moving_var * momentum + var.detach() * (1.0 - momentum)
    return y, new_mean, new_var


def batch_norm_infer(x, gamma, beta, moving_mean, moving_var, eps=1e-3):
    """keras ``BatchNormalization()`` with ``training=False``."""
    return (x - moving_mean) * torch.rsqrt(moving_var + eps) * gamma + beta


def relu(x):
    """keras ``Activation("relu")``."""
    return torch.relu(x)


# --------------------------------------------------------------------------- #
# Resampling
# --------------------------------------------------------------------------- #
def resize_bilinear(x, out_h, out_w, antialias=True):
    """``tf.image.resize(x, [out_h, out_w], "bilinear", antialias=antialias)``.

    Separable: columns, then rows, with the ScaleAndTranslate span weights of
    ``oracle/resize_np.py``.  The backward pass is the exact transpose.
    Call sites: shared/custom_layers.py:102 and :124.
    """
    n, h, w, c = x.shape
    if h == out_h and w == out_w:
        return x
    rw = torch.from_numpy(resize_np.resize_matrix(w, out_w, antialias)).to(x.dtype)
    rh = torch.from_numpy(resize_np.resize_matrix(h, out_h, antialias)).to(x.dtype)
    t = torch.einsum("pw,nhwc->nhpc", rw, x)
    return torch.einsum("qh,nhpc->nqpc", rh, t)


def resize_by_scale(x, scale, antialias=True):
    """``ResizeByScale.call`` -- shared/custom_layers.py:93-103."""
    nh = resize_np.resized_extent(x.shape[1], scale)
    nw = resize_np.resized_extent(x.shape[2], scale)
    return resize_bilinear(x, nh, nw, antialias)


def resize_to_match(x, ref, antialias=True):
    """``ResizeToMatch.call`` -- shared/custom_layers.py:121-125."""
    return resize_bilinear(x, ref.shape[1], ref.shape[2], antialias)


def upsample2_bilinear(x):
    """keras ``UpSampling2D(2, interpolation="bilinear")``: half-pixel bilinear,
    no antialias.  Call site: Segmenation/code/train_adaptive_unet.py:357."""
    return resize_bilinear(x, 2 * x.shape[1], 2 * x.shape[2], antialias=False)


def max_pool2(x):
    """keras ``MaxPooling2D(2)``: valid, stride 2, floor.
    Call sites: Segmenation/code/train_adaptive_unet.py:351; unet_vinillia.py:62."""
    return _nhwc(F.max_pool2d(_nchw(x), 2, 2))


# --------------------------------------------------------------------------- #
# Head
# --------------------------------------------------------------------------- #
def clipped_residual_add(inp, residual):
    """``ClippedResidualAdd.call`` -- shared/custom_layers.py:136-139.

    ``tf.clip_by_value``'s gradient passes through on the closed interval
    [0, 1]; torch.clamp does the same.
    """
    return torch.clamp(inp + residual, 0.0, 1.0)


# --------------------------------------------------------------------------- #
# Losses and metrics
# --------------------------------------------------------------------------- #
def charbonnier_loss(y_true, y_pred, eps=1e-3):
    """train_adaptive_unet.py:313-320."""
    d = y_true - y_pred
    return torch.sqrt(d * d + eps * eps).mean()


def l1_loss(y_true, y_pred):
    """train_adaptive_unet.py:326-330."""
    return (y_true - y_pred).abs().mean()


def mse_loss(y_true, y_pred):
    """train_adaptive_unet.py:345-348."""
    return ((y_true - y_pred) ** 2).mean()


def psnr_metric(y_true, y_pred):
    """train_adaptive_unet.py:308-311: mean over the batch of tf.image.psnr."""
    p = torch.clamp(y_pred, 0.0, 1.0)
    mse = ((y_true - p) ** 2).mean(dim=(1, 2, 3))
    return (10.0 * torch.log10(1.0 / mse)).mean()


_EPS7 = 1e-7


def binary_crossentropy(y_true, y_pred):
    """keras ``BinaryCrossentropy()`` on probabilities: clip to [1e-7, 1-1e-7],
    mean over every element.  Segmenation/code/train_adaptive_unet.py:284,296."""
    p = torch.clamp(y_pred, _EPS7, 1.0 - _EPS7)
    bce = -(y_true * torch.log(p) + (1.0 - y_true) * torch.log(1.0 - p))
    return bce.mean()


def dice_coefficient(y_true, y_pred, smooth=1e-6):
    """Per-sample dice, batch mean.  Segmenation/code/train_adaptive_unet.py:258-265."""
    p = torch.clamp(y_pred, _EPS7, 1.0 - _EPS7)
    inter = (y_true * p).sum(dim=(1, 2, 3))
    union = (y_true + p).sum(dim=(1, 2, 3))
    return ((2.0 * inter + smooth) / (union + smooth)).mean()


def dice_coefficient_global(y_true, y_pred, smooth=1e-6):
    """The baseline trainer's Dice: ONE ratio over the whole batch, no clipping.  Segmenation/code/unet_vinillia.py:94-99."""
    return (2.0 * (y_true * y_pred).sum() + smooth) / ((y_true + y_pred).sum() + smooth)


def iou_score(y_true, y_pred, smooth=1e-6):
    """Segmenation/code/train_adaptive_unet.py:272-280."""
    p = torch.clamp(y_pred, _EPS7, 1.0 - _EPS7)
    inter = (y_true * p).sum(dim=(1, 2, 3))
    total = (y_true + p).sum(dim=(1, 2, 3))
    return ((inter + smooth) / (total - inter + smooth)).mean()


def bce_dice_loss(y_true, y_pred, bce_weight, dice_weight):
    """``make_hybrid_ce_dice_loss`` / ``make_bce_dice_loss`` -- :283-304."""
    return bce_weight * binary_crossentropy(y_true, y_pred) + dice_weight * (
        1.0 - dice_coefficient(y_true, y_pred)
    )


def categorical_crossentropy(y_true_onehot, y_pred):
    """keras ``CategoricalCrossentropy(from_logits=False)``: renormalise by the
    class sum, clip to [1e-7, 1-1e-7], mean over batch*pixels.  NOT in the
    reference; the extrapolated loss for config C4."""
    p = y_pred / y_pred.sum(dim=-1, keepdim=True)
    p = torch.clamp(p, _EPS7, 1.0 - _EPS7)
    return -(y_true_onehot * torch.log(p)).sum(dim=-1).mean()


# --------------------------------------------------------------------------- #
# Optimiser
# --------------------------------------------------------------------------- #
def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-7):
    """keras ``Adam`` update (train_adaptive_unet.py:490): ``step`` is 1-based.

    alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t);  m += (g - m)(1 - beta1);
    v += (g^2 - v)(1 - beta2);  p -= alpha * m / (sqrt(v) + eps).
    Returns the new (p, m, v).
    """
    alpha = lr * (1.0 - beta2 ** step) ** 0.5 / (1.0 - beta1 ** step)
    m = m + (g - m) * (1.0 - beta1)
    v = v + (g * g - v) * (1.0 - beta2)
    p = p - alpha * m / (torch.sqrt(v) + eps)
    return p, m, v
