"""Host-side input pipeline with the reference's function names (``shared/pipeline.py``).

The kernels of this repo start where these generators end: they yield ``(lr, hr)`` float32 NHWC
batches in [0, 1], exactly what ``/root/reference/shared/pipeline.py:214-288`` hands to ``Model.fit``
through ``tf.data``.  Here the dataset objects are plain Python iterables (no TensorFlow); image
decoding / resizing uses OpenCV as the reference does (``degrade_image`` :79-94 = INTER_AREA down,
INTER_CUBIC up, not clipped).
"""
from __future__ import annotations

import re
from pathlib import Path
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np


def _cv2():
    import cv2
    return cv2


def sorted_alphanumeric(items: Iterable[str]) -> List[str]:
    """Natural sort: digit runs compare as integers, the rest case-insensitively."""
    def key(text: str):
        return [int(tok) if tok.isdigit() else tok.lower() for tok in re.findall(r"\d+|\D+", text)]
    return sorted(items, key=key)


def load_rgb_image_full(path) -> np.ndarray:
    """Decode an image file to float32 RGB in [0, 1] at its native size."""
    cv2 = _cv2()
    bgr = cv2.imread(str(path), cv2.IMREAD_COLOR)
    if bgr is None:
        raise FileNotFoundError(f"Unable to read image: {path}")
    return cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB).astype(np.float32) / 255.0


def load_rgb_image(path, size: int) -> np.ndarray:
    cv2 = _cv2()
    img = load_rgb_image_full(path)
    return cv2.resize(img, (size, size), interpolation=cv2.INTER_AREA)


def degrade_image(image: np.ndarray, scale: float, output_size: int) -> np.ndarray:
    """LR input synthesis: area-average shrink to round(output_size*scale), bicubic back up."""
    if not 0 < scale < 1:
        raise ValueError("Scale must be between 0 and 1 for degradation.")
    cv2 = _cv2()
    hr = np.clip(np.asarray(image, dtype=np.float32), 0.0, 1.0)
    side = output_size if output_size > 0 else max(hr.shape[:2])
    small = max(1, int(round(side * scale)))
    down = cv2.resize(hr, (small, small), interpolation=cv2.INTER_AREA)
    return cv2.resize(down, (side, side), interpolation=cv2.INTER_CUBIC).astype(np.float32)


def _check_patch_args(image: np.ndarray, patch_size: int):
    if patch_size <= 0:
        raise ValueError("patch_size must be positive.")
    if image.ndim != 3 or image.shape[-1] != 3:
        raise ValueError("image must be an HxWx3 RGB array.")
    if image.shape[0] < patch_size or image.shape[1] < patch_size:
        raise ValueError("patch_size exceeds image dimensions.")


def random_patch(image: np.ndarray, patch_size: int, *, rng: Optional[np.random.Generator] = None) -> np.ndarray:
    _check_patch_args(image, patch_size)
    gen = rng or np.random.default_rng()
    free_y, free_x = image.shape[0] - patch_size, image.shape[1] - patch_size
    top = int(gen.integers(0, free_y + 1)) if free_y > 0 else 0
    left = int(gen.integers(0, free_x + 1)) if free_x > 0 else 0
    return image[top:top + patch_size, left:left + patch_size, :]


def random_patches(image: np.ndarray, patch_size: int, count: int, *, rng: Optional[np.random.Generator] = None):
    if count <= 0:
        raise ValueError("count must be positive.")
    gen = rng or np.random.default_rng()
    return np.stack([random_patch(image, patch_size, rng=gen) for _ in range(count)], axis=0)


def grid_patches(image: np.ndarray, patch_size: int, *, stride: Optional[int] = None, drop_remainder: bool = False):
    _check_patch_args(image, patch_size)
    stride = stride or patch_size
    if stride <= 0:
        raise ValueError("stride must be positive.")
    tops = range(0, image.shape[0] - patch_size + 1, stride)
    lefts = range(0, image.shape[1] - patch_size + 1, stride)
    out = [image[t:t + patch_size, l:l + patch_size, :] for t in tops for l in lefts]
    if not out and not drop_remainder:
        out.append(image[-patch_size:, -patch_size:, :])
    return np.stack(out, axis=0) if out else np.empty((0, patch_size, patch_size, 3), dtype=image.dtype)


class PatchDataset:
    """Re-iterable stream of (lr, hr) batches; stands in for the reference's tf.data.Dataset."""

    def __init__(self, make_iter, batch_size: int, shuffle_buffer: int = 0, seed: int = 0, infinite: bool = False):
        self._make_iter, self.batch_size = make_iter, batch_size
        self.shuffle_buffer, self.seed, self.infinite = shuffle_buffer, seed, infinite
        self._epoch = 0

    def repeat(self):
        return PatchDataset(self._make_iter, self.batch_size, self.shuffle_buffer, self.seed, infinite=True)

    def _pairs(self) -> Iterator[Tuple[np.ndarray, np.ndarray]]:
        while True:
            it = self._make_iter()
            if self.shuffle_buffer > 0:
                rng = np.random.default_rng(self.seed + self._epoch)
                buf = []
                for pair in it:
                    if len(buf) < self.shuffle_buffer:
                        buf.append(pair)
                        continue
                    k = int(rng.integers(0, len(buf)))
                    out, buf[k] = buf[k], pair
                    yield out
                rng.shuffle(buf)
                yield from buf
            else:
                yield from it
            self._epoch += 1
            if not self.infinite:
                return

    def __iter__(self):
        lrs, hrs = [], []
        for lr, hr in self._pairs():
            lrs.append(lr); hrs.append(hr)
            if len(lrs) == self.batch_size:
                yield np.stack(lrs), np.stack(hrs)
                lrs, hrs = [], []
        if lrs:
            yield np.stack(lrs), np.stack(hrs)


def _random_pairs(files, patch_size, per_image, scale, seed):
    rng = np.random.default_rng(seed)
    files = list(files)
    while True:
        rng.shuffle(files)
        for path in files:
            image = load_rgb_image_full(path)
            for hr in random_patches(image, patch_size, count=per_image, rng=rng):
                yield degrade_image(hr, scale, patch_size), np.ascontiguousarray(hr)


def _grid_pairs(files, patch_size, stride, scale):
    for path in files:
        for hr in grid_patches(load_rgb_image_full(path), patch_size, stride=stride):
            yield degrade_image(hr, scale, patch_size), np.ascontiguousarray(hr)


def make_training_patch_dataset(hr_files: Sequence[str], patch_size: int, patches_per_image: int, scale: float,
                                batch_size: int, seed: int, shuffle_buffer: int = 1024):
    """Infinite stream of random (lr, hr) patch batches; returns (dataset, patches per epoch)."""
    hr_files = list(hr_files)
    if not hr_files:
        raise ValueError("hr_files must contain at least one path.")
    if patches_per_image <= 0:
        raise ValueError("patches_per_image must be positive.")
    ds = PatchDataset(lambda: _random_pairs(hr_files, patch_size, patches_per_image, scale, seed), batch_size,
                      shuffle_buffer, seed, infinite=True)
    return ds, len(hr_files) * patches_per_image


def make_eval_patch_dataset(hr_files: Sequence[str], patch_size: int, scale: float, batch_size: int, *,
                            stride: Optional[int] = None):
    """Finite stream of grid patches; returns (dataset, patch count, patch labels)."""
    hr_files = list(hr_files)
    if not hr_files:
        raise ValueError("hr_files must contain at least one path.")
    stride = stride or patch_size
    if stride <= 0:
        raise ValueError("stride must be positive.")
    labels = []
    for path in hr_files:
        n = grid_patches(load_rgb_image_full(path), patch_size, stride=stride).shape[0]
        labels += [f"{Path(path).name}#patch{i:04d}" for i in range(n)]
    ds = PatchDataset(lambda: _grid_pairs(hr_files, patch_size, stride, scale), batch_size)
    return ds, len(labels), labels


def split_indices(n_samples: int, train: float, val: float, test: float, seed: int):
    """Shuffled train/val/test index split with the reference's rounding and minimum-size rules."""
    if not 0 < train < 1:
        raise ValueError("Train fraction should be between 0 and 1.")
    if not 0 <= val < 1 or not 0 <= test < 1:
        raise ValueError("Val/test fractions should be between 0 and 1.")
    total = train + val + test
    order = np.arange(n_samples)
    np.random.default_rng(seed).shuffle(order)
    n_train = int(round(n_samples * train / total))
    n_val = int(round(n_samples * val / total))
    if n_samples > 2:
        n_train = min(n_train, n_samples - 2)
    if n_samples > n_train + 1:
        n_val = min(n_val, n_samples - n_train - 1)
    if n_train <= 0:
        raise ValueError("Train split is empty; adjust fractions.")
    return order[:n_train], order[n_train:n_train + n_val], order[n_train + n_val:]
