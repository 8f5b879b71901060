"""numpy restatement of the two OpenCV resizes behind the reference's LR synthesis.

Oracle code (test infrastructure, see ``oracle/__init__.py``).

``degrade_image`` (``/root/reference/shared/pipeline.py:79-94``) is
``cv2.resize(clip(hr,0,1), (s,s), INTER_AREA)`` followed by ``cv2.resize(., (P,P), INTER_CUBIC)`` on float32
HxWx3 arrays, with ``s = max(1, int(round(P*scale)))``; the result is NOT clipped.  The arithmetic lives in
OpenCV (``opencv-python-headless==4.9.0.80``, ``Super_resolution/requirement.txt:16``), which is not vendored
in the reference; its published algorithm (``modules/imgproc/src/resize.cpp``) is restated here:

* INTER_AREA, shrinking: per axis ``scale = in/out`` (double).  Output ``d`` covers the source interval
  ``[d*scale, (d+1)*scale)``; ``cell = min(scale, in - d*scale)``; the whole source pixels inside get weight
  ``1/cell`` and the two partially covered ones their covered fraction ``/cell`` when that fraction exceeds
  1e-3 (``computeResizeAreaTab``).  Integer ratios take the block-average fast path, which is the same table
  with every weight ``1/scale``.  Rows are reduced horizontally first, then vertically, in float32.
* INTER_CUBIC: ``fx = float32((d+0.5)*in/out - 0.5)``, ``sx = floor(fx)``, ``t = fx - sx``; four taps at
  ``sx-1 .. sx+2`` with the Keys kernel A = -0.75 evaluated in float32 (``interpolateCubic``; the 4th weight is
  ``1 - w0 - w1 - w2``); out-of-range taps are clamped to the border pixel (replicate).  Horizontal pass first,
  then vertical, float32.

UNLIKE the TensorFlow ops this restatement IS pinned: OpenCV is importable in the build container, so
``tests/test_pipeline_cpu.py`` checks it against ``cv2.resize`` itself and against fixtures produced by the
reference's own ``degrade_image`` / ``random_patches`` / ``grid_patches`` (``tests/golden/make_pipeline_golden.py``).
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32
INTER_AREA, INTER_CUBIC = 0, 1


def degraded_extent(side: int, scale: float) -> int:
    """``max(1, int(round(side*scale)))`` -- pipeline.py:89-90 (Python round: half to even)."""
    return max(1, int(round(side * scale)))


def area_table(in_size: int, out_size: int):
    """(idx[int32 out x taps], w[float32 out x taps]) of cv2 INTER_AREA for in_size >= out_size.

    Unused slots repeat the last valid index with weight 0."""
    if out_size > in_size:
        raise ValueError("INTER_AREA table is for shrinking (cv2 falls back to bilinear when enlarging)")
    scale = in_size / out_size
    rows = []
    iscale = int(round(scale))
    if abs(scale - iscale) < np.finfo(np.float64).eps:            # is_area_fast: plain block average
        for d in range(out_size):
            rows.append([(d * iscale + k, F32(1.0 / iscale)) for k in range(iscale)])
    else:
        for d in range(out_size):
            f1 = d * scale
            f2 = f1 + scale
            cell = min(scale, in_size - f1)
            s1, s2 = math.ceil(f1), math.floor(f2)
            s2 = min(s2, in_size - 1)
            s1 = min(s1, s2)
            row = []
            if s1 - f1 > 1e-3:
                row.append((s1 - 1, F32((s1 - f1) / cell)))
            for s in range(s1, s2):
                row.append((s, F32(1.0 / cell)))
            if f2 - s2 > 1e-3:
                row.append((s2, F32(min(min(f2 - s2, 1.0), cell) / cell)))
            rows.append(row)
    taps = max(len(r) for r in rows)
    idx = np.zeros((out_size, taps), dtype=np.int32)
    w = np.zeros((out_size, taps), dtype=F32)
    for d, row in enumerate(rows):
        for k in range(taps):
            if k < len(row):
                idx[d, k], w[d, k] = row[k]
            else:
                idx[d, k] = row[-1][0]
    return idx, w


def cubic_coeffs(t) -> np.ndarray:
    """interpolateCubic: Keys kernel with A = -0.75, float32 arithmetic in OpenCV's operation order."""
    a = F32(-0.75)
    t = F32(t)
    one = F32(1.0)
    c0 = ((a * (t + one) - F32(5) * a) * (t + one) + F32(8) * a) * (t + one) - F32(4) * a
    c1 = ((a + F32(2)) * t - (a + F32(3))) * t * t + one
    u = one - t
    c2 = ((a + F32(2)) * u - (a + F32(3))) * u * u + one
    c3 = one - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], dtype=F32)


def cubic_table(in_size: int, out_size: int):
    """(idx[int32 out x 4], w[float32 out x 4]) of cv2 INTER_CUBIC, border taps clamped (replicate)."""
    scale = in_size / out_size
    idx = np.zeros((out_size, 4), dtype=np.int32)
    w = np.zeros((out_size, 4), dtype=F32)
    for d in range(out_size):
        fx = F32((d + 0.5) * scale - 0.5)
        sx = int(math.floor(float(fx)))
        t = fx - F32(sx)
        w[d] = cubic_coeffs(t)
        for k in range(4):
            idx[d, k] = min(max(sx - 1 + k, 0), in_size - 1)
    return idx, w


def table(in_size: int, out_size: int, interp: int):
    return area_table(in_size, out_size) if interp == INTER_AREA else cubic_table(in_size, out_size)


def apply_tables(x: np.ndarray, yt, xt) -> np.ndarray:
    """out[..., oy, ox, c] = sum_j wy[oy,j] * (sum_k wx[ox,k] * x[..., iy[oy,j], ix[ox,k], c]), float32,
    taps accumulated in table order (horizontal pass first, as OpenCV does)."""
    (iy, wy), (ix, wx) = yt, xt
    x = np.asarray(x, dtype=F32)
    rows = np.zeros(x.shape[:-2] + (ix.shape[0], x.shape[-1]), dtype=F32)
    for k in range(ix.shape[1]):
        rows = rows + x[..., ix[:, k], :] * wx[:, k][:, None]
    rows = rows.astype(F32)
    out = np.zeros(x.shape[:-3] + (iy.shape[0],) + rows.shape[-2:], dtype=F32)
    for j in range(iy.shape[1]):
        out = out + rows[..., iy[:, j], :, :] * wy[:, j][:, None, None]
    return out.astype(F32)


def resize(x: np.ndarray, out_h: int, out_w: int, interp: int) -> np.ndarray:
    """cv2.resize(x, (out_w, out_h), interpolation=INTER_AREA|INTER_CUBIC) for float32 (..., H, W, C)."""
    h, w = x.shape[-3], x.shape[-2]
    return apply_tables(x, table(h, out_h, interp), table(w, out_w, interp))


def degrade_image(image: np.ndarray, scale: float, output_size: int) -> np.ndarray:
    """pipeline.py:79-94 on (..., H, W, 3) float32; leading axes are a batch of patches."""
    if not 0 < scale < 1:
        raise ValueError("Scale must be between 0 and 1 for degradation.")
    hr = np.clip(np.asarray(image, dtype=F32), 0.0, 1.0)
    side = output_size if output_size > 0 else max(hr.shape[-3], hr.shape[-2])
    small = degraded_extent(side, scale)
    return resize(resize(hr, small, small, INTER_AREA), side, side, INTER_CUBIC)


def patch_origins(height: int, width: int, patch_size: int, count: int, rng: np.random.Generator) -> np.ndarray:
    """The (top, left) draws of ``random_patches`` (pipeline.py:97-136): per patch one ``rng.integers`` draw for
    the row (skipped when the image is exactly patch-high) and then one for the column."""
    out = np.zeros((count, 2), dtype=np.int32)
    for i in range(count):
        max_y, max_x = height - patch_size, width - patch_size
        out[i, 0] = int(rng.integers(0, max_y + 1)) if max_y > 0 else 0
        out[i, 1] = int(rng.integers(0, max_x + 1)) if max_x > 0 else 0
    return out


def grid_origins(height: int, width: int, patch_size: int, stride: int | None = None) -> np.ndarray:
    """Row-major (top, left) grid of ``grid_patches`` (pipeline.py:139-175, drop_remainder=False)."""
    stride = stride or patch_size
    pts = [(t, l) for t in range(0, height - patch_size + 1, stride) for l in range(0, width - patch_size + 1, stride)]
    if not pts:
        pts.append((height - patch_size, width - patch_size))
    return np.asarray(pts, dtype=np.int32).reshape(-1, 2)


def crop(image: np.ndarray, origins: np.ndarray, patch_size: int) -> np.ndarray:
    """uint8 or float32 HxWx3 -> float32 [n,P,P,3]; uint8 is scaled as load_rgb_image_full does (/255 in float32)."""
    img = np.asarray(image)
    if img.dtype == np.uint8:
        img = img.astype(F32) / F32(255.0)
    return np.stack([img[t:t + patch_size, l:l + patch_size, :] for t, l in origins], axis=0).astype(F32)
