"""numpy restatement of TensorFlow's ``ScaleAndTranslate`` span/weight rule.

Oracle code (test infrastructure, see ``oracle/__init__.py``).

``tf.image.resize(x, size, method="bilinear", antialias=True)`` is what the
reference calls in ``shared/custom_layers.py:102`` (ResizeByScale) and
``shared/custom_layers.py:124`` (ResizeToMatch).  In TF 2.16.1 that call lowers to
the ``ScaleAndTranslate`` op with a triangle kernel (radius 1), zero translation
and ``scale = float32(out) / float32(in)`` per axis.  The op is separable: it
computes, per output index, a contiguous span of source indices and normalised
weights, and applies them first along one axis and then along the other, all in
float32.  This file restates that published rule; the third-party source is not in
``/root/reference`` (pinned versions: tensorflow==2.16.1, keras==3.3.3).
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


def resized_extent(extent: int, scale: float) -> int:
    """``max(1, ceil(float32(extent) * scale))`` -- custom_layers.py:97-100.

    TF multiplies a float32 tensor by the Python float ``self.scale``; the
    constant is converted to float32 first, so the product is a float32 product.
    """
    prod = F32(extent) * F32(scale)
    return max(1, int(math.ceil(float(prod))))


def triangle_spans(in_size: int, out_size: int, antialias: bool = True):
    """Return (starts[int32 out], weights[float32 out x span]) for one axis.

    Follows ScaleAndTranslate's ComputeSpansCore with a triangle kernel:
      inv_scale    = 1 / (float32(out)/float32(in))
      kernel_scale = max(inv_scale, 1) if antialias else 1
      span_size    = min(2*ceil(kernel_scale) + 1, in)
      sample       = (x + 0.5) * inv_scale
      span         = [ceil(sample - kernel_scale - 0.5), floor(sample + kernel_scale - 0.5)]
                     clamped to [0, in-1]
      weight(j)    = max(0, 1 - |j + 0.5 - sample| / kernel_scale), normalised to sum 1.
    All arithmetic in float32, as in the op.
    """
    scale = F32(out_size) / F32(in_size)
    inv_scale = F32(1.0) / scale
    kernel_scale = max(inv_scale, F32(1.0)) if antialias else F32(1.0)
    span_size = min(2 * int(math.ceil(float(kernel_scale))) + 1, in_size)
    starts = np.zeros(out_size, dtype=np.int32)
    weights = np.zeros((out_size, span_size), dtype=F32)
    one_over_ks = F32(1.0) / kernel_scale
    for x in range(out_size):
        sample = F32(F32(x) + F32(0.5)) * inv_scale
        if sample < 0 or sample > in_size:
            continue
        lo = int(math.ceil(float(F32(F32(sample - kernel_scale) - F32(0.5)))))
        hi = int(math.floor(float(F32(F32(sample + kernel_scale) - F32(0.5)))))
        lo = min(max(lo, 0), in_size - 1)
        hi = min(max(hi, 0), in_size - 1) + 1
        total = F32(0.0)
        tmp = []
        for src in range(lo, hi):
            pos = F32(F32(F32(src) + F32(0.5)) - sample)
            wgt = max(F32(0.0), F32(F32(1.0) - abs(F32(pos * one_over_ks))))
            total = F32(total + wgt)
            tmp.append(wgt)
        if abs(total) >= F32(1000.0) * np.finfo(F32).tiny:
            inv_total = F32(1.0) / total
            for k, wgt in enumerate(tmp):
                weights[x, k] = F32(wgt * inv_total)
        starts[x] = lo
    return starts, weights


def resize_matrix(in_size: int, out_size: int, antialias: bool = True) -> np.ndarray:
    """Dense [out, in] float32 resampling matrix for one axis (identity if equal).

    ``tf.image.resize`` returns its input unchanged when the size is unchanged
    (the short-cut in ``_resize_images_common``), hence the identity.
    """
    if in_size == out_size:
        return np.eye(in_size, dtype=F32)
    starts, weights = triangle_spans(in_size, out_size, antialias)
    mat = np.zeros((out_size, in_size), dtype=F32)
    for x in range(out_size):
        for k in range(weights.shape[1]):
            j = starts[x] + k
            if j < in_size and weights[x, k] != 0:
                mat[x, j] += weights[x, k]
    return mat


def size_chain(extent: int, scale: float, depth: int):
    """[extent, after 1 down, ... after ``depth`` downs] (ResizeByScale chain)."""
    out = [extent]
    for _ in range(depth):
        out.append(resized_extent(out[-1], scale))
    return out
