"""Drop-in interface of the reference's ``shared/custom_layers.py``.

Same public names, arguments and serialisation keys as
/root/reference/shared/custom_layers.py (ResizeByScale :85-111, ResizeToMatch :114-132,
ClippedResidualAdd :134-139, ClipAdd alias :142, depth rules :10-82); the layers are
symbolic graph nodes here and execute as the library's resampling / clip-add kernels.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np

from ..keras.layers import Layer

_SERIALIZABLE: Dict[str, type] = {}


def register_keras_serializable(package: str):
    """Record ``package>ClassName`` keys the way keras.saving does (custom_layers.py:85,114,134)."""

    def deco(cls):
        _SERIALIZABLE[f"{package}>{cls.__name__}"] = cls
        return cls

    return deco


def get_custom_objects() -> Dict[str, type]:
    return dict(_SERIALIZABLE)


def _check_scale(scale: float) -> None:
    if not (0.05 < scale < 1.0):
        raise ValueError("Scale should be between 0 and 1 (exclusive).")


def infer_depth_from_scale(scale: float, min_depth: int = 1, max_depth: int = 4) -> int:
    """Design-table depth: <=0.25 -> 1, <=0.45 -> 2, else 3, clamped to [min_depth, max_depth]."""
    _check_scale(scale)
    table = ((0.25, 1), (0.45, 2))
    depth = next((d for lim, d in table if scale <= lim), 3)
    return max(min_depth, min(depth, max_depth))


def depth_and_sizes(scale, min_res=21, max_depth=7):
    """Size chain from 256 while the extent stays above ``min_res`` (diagnostic helper)."""
    sizes = [256]
    while sizes[-1] > min_res and len(sizes) < max_depth:
        sizes.append(math.ceil(sizes[-1] * scale))
    return min(len(sizes), max_depth), sizes


def custom_depth_from_scale(scale: float, min_depth: int = 1, max_depth: int = 7, *, base_resolution: int = 256,
                            min_feature: int = 21) -> int:
    """Encoder depth: keep shrinking by ``ceil(extent * scale)`` while the next extent is still at
    least ``min_feature`` pixels and the depth limit is not reached."""
    _check_scale(scale)
    for label, val, lo in (("min_depth", min_depth, 1), ("max_depth", max_depth, 1)):
        if val < lo:
            raise ValueError(f"{label} must be at least 1.")
    if base_resolution <= 0:
        raise ValueError("base_resolution must be positive.")
    if min_feature < 1:
        raise ValueError("min_feature must be at least 1 pixel.")
    depth, extent = max(min_depth, 1), base_resolution
    while depth < max_depth:
        nxt = math.ceil(extent * scale)
        if nxt < min_feature:
            break
        extent, depth = nxt, depth + 1
    return max(min_depth, min(depth, max_depth))


def estimate_bottleneck_size(hr: int, scale: float, depth: int) -> int:
    """Diagnostic bottleneck extent (uses round(), unlike the layer's ceil())."""
    size = hr
    for _ in range(depth):
        size = max(1, int(round(size * scale)))
    return size


def resized_extent(extent: int, scale: float) -> int:
    """max(1, ceil(float32(extent) * float32(scale))) -- the float32 product TF evaluates (:97-100)."""
    return max(1, int(math.ceil(float(np.float32(extent) * np.float32(scale)))))


@register_keras_serializable(package="resize")
class ResizeByScale(Layer):
    """Shrink H and W by ``scale`` with an antialiased bilinear filter."""

    def __init__(self, scale: float, method: str = "bilinear", antialias: bool = True, name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        if method != "bilinear":
            raise NotImplementedError("ResizeByScale: only method='bilinear'")
        self.scale, self.method, self.antialias = float(scale), method, bool(antialias)

    def compute_output_shape(self, s):
        n, h, w, c = s[0]
        return (n, resized_extent(h, self.scale), resized_extent(w, self.scale), c)

    def get_config(self):
        return {**super().get_config(), "scale": self.scale, "method": self.method, "antialias": self.antialias}


@register_keras_serializable(package="resize")
class ResizeToMatch(Layer):
    """Resize the first input to the spatial size of the second (reference) input."""

    def __init__(self, method: str = "bilinear", antialias: bool = True, name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        if method != "bilinear":
            raise NotImplementedError("ResizeToMatch: only method='bilinear'")
        self.method, self.antialias = method, bool(antialias)

    def compute_output_shape(self, s):
        x, ref = s
        return (x[0], ref[1], ref[2], x[3])

    def get_config(self):
        return {**super().get_config(), "method": self.method, "antialias": self.antialias}


@register_keras_serializable(package="utils")
class ClippedResidualAdd(Layer):
    """clip(input + residual, 0, 1) evaluated in fp32, returned in the input's dtype."""

    def compute_output_shape(self, s):
        return s[0]


# legacy checkpoints and configs refer to the layer under this name
ClipAdd = ClippedResidualAdd
_SERIALIZABLE["utils>ClipAdd"] = ClippedResidualAdd
