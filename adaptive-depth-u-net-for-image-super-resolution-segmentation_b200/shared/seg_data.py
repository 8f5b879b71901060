"""Host-side ISIC-style segmentation data feed (image/mask pairs -> NHWC float32 batches).

Stand-in for the tf.data pipelines of /root/reference/Segmenation/code/train_adaptive_unet.py:71-251
(collect_isic_pairs, load_isic_image/mask, apply_isic_augmentations, build_isic_dataset) and
unet_vinillia.py:100-205 (_discover_pairs, _parse_example, _augment, build_dataset).  Dataset IO is outside the
hot path: this module only has to hand (image [B,S,S,3] in [0,1], mask [B,S,S,1] in {0,1})
batches to ``Model.fit``; decoding/resizing uses cv2 on the host.
"""
from __future__ import annotations

from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .pipeline import _cv2, sorted_alphanumeric

_MASK_TOKENS = ("_segmentation", "_mask", "_leftimg8bit", "_gtfine_labelids", "_gtfine_polygons", "_gtfine_color",
                "_gtfine_instanceids", "_gtcoarse_labelids", "_gtcoarse_color", "_gtcoarse_instanceids", "_instanceids")


def canonical_key(path: Path) -> str:
    """Lower-case stem with the mask suffix tokens removed (train_adaptive_unet.py:71-75, unet_vinillia.py:100-118)."""
    stem = Path(path).stem.lower()
    for tok in _MASK_TOKENS:
        stem = stem.replace(tok, "")
    return stem


def collect_pairs(image_dir, mask_dir, image_suffixes: Sequence[str] = (".jpg", ".jpeg", ".png"),
                  mask_suffixes: Sequence[str] = (".png", ".jpg"), limit: Optional[int] = None,
                  require_segmentation_token: bool = False) -> List[Tuple[str, str]]:
    """Align images with masks by canonical key; error behaviour follows the reference (missing directory ->
    FileNotFoundError, image without mask -> ValueError)."""
    image_dir, mask_dir = Path(image_dir), Path(mask_dir)
    if not image_dir.exists():
        raise FileNotFoundError(f"Image directory does not exist: {image_dir}")
    if not mask_dir.exists():
        raise FileNotFoundError(f"Mask directory does not exist: {mask_dir}")
    images = [p for p in image_dir.rglob("*") if p.is_file() and p.suffix.lower() in image_suffixes
              and "superpixels" not in p.stem.lower()]
    masks = [p for p in mask_dir.rglob("*") if p.is_file() and p.suffix.lower() in mask_suffixes
             and (not require_segmentation_token or p.stem.lower().endswith("_segmentation"))]
    if not images:
        raise FileNotFoundError(f"No image files found in {image_dir}")
    if not masks:
        raise FileNotFoundError(f"No mask files found in {mask_dir}")
    index = {canonical_key(p): p for p in masks}
    pairs, missing = [], []
    for name in sorted_alphanumeric(str(p) for p in images):
        m = index.get(canonical_key(Path(name)))
        if m is None:
            missing.append(Path(name).name)
        else:
            pairs.append((name, str(m)))
    if missing:
        raise ValueError(f"Missing {len(missing)} segmentation masks in {mask_dir}; examples: {', '.join(missing[:5])}")
    return pairs[:limit] if limit else pairs


def load_pair(image_path: str, mask_path: str, size: int, area: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    cv2 = _cv2()
    img = cv2.imread(image_path, cv2.IMREAD_COLOR)
    msk = cv2.imread(mask_path, cv2.IMREAD_GRAYSCALE)
    if img is None or msk is None:
        raise FileNotFoundError(f"cannot decode {image_path} / {mask_path}")
    img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB).astype(np.float32) / 255.0
    img = cv2.resize(img, (size, size), interpolation=cv2.INTER_AREA if area else cv2.INTER_LINEAR)
    msk = cv2.resize(msk, (size, size), interpolation=cv2.INTER_NEAREST).astype(np.float32) / 255.0
    return img, (msk > 0.5).astype(np.float32)[..., None]


def augment(img: np.ndarray, msk: np.ndarray, rng: np.random.Generator, rich: bool) -> Tuple[np.ndarray, np.ndarray]:
    """Flips (both trainers); `rich` adds the adaptive trainer's rot90 + random 1.0-1.15 zoom-and-crop (:139-174)."""
    if rich:
        k = int(rng.integers(0, 4))
        img, msk = np.rot90(img, k), np.rot90(msk, k)
    if rng.random() > 0.5:
        img, msk = img[:, ::-1], msk[:, ::-1]
    if rng.random() > 0.5:
        img, msk = img[::-1], msk[::-1]
    if rich:
        cv2 = _cv2()
        size = img.shape[0]
        s = int(round(float(rng.uniform(1.0, 1.15)) * size))
        img = cv2.resize(np.ascontiguousarray(img), (s, s), interpolation=cv2.INTER_LINEAR)
        msk = cv2.resize(np.ascontiguousarray(msk), (s, s), interpolation=cv2.INTER_NEAREST)[..., None]
        oy, ox = int(rng.integers(0, s - size + 1)), int(rng.integers(0, s - size + 1))
        img, msk = img[oy:oy + size, ox:ox + size], (msk[oy:oy + size, ox:ox + size] > 0.5).astype(np.float32)
    return np.ascontiguousarray(img), np.ascontiguousarray(msk)


class SegDataset:
    """Re-iterable (image, mask) batch stream; reshuffles every epoch like tf.data's shuffle(reshuffle_each_iteration)."""

    def __init__(self, pairs, image_size: int, batch_size: int, shuffle: bool, augment_mode: Optional[str], seed: int,
                 area_resize: bool = True, arrays: Optional[Tuple[np.ndarray, np.ndarray]] = None):
        self.pairs, self.size, self.batch_size = list(pairs), image_size, batch_size
        self.shuffle, self.augment_mode, self.seed, self.area = shuffle, augment_mode, seed, area_resize
        self.arrays = arrays
        self._epoch = 0

    def __len__(self):
        n = len(self.pairs) if self.arrays is None else len(self.arrays[0])
        return (n + self.batch_size - 1) // self.batch_size

    @property
    def samples(self):
        return len(self.pairs) if self.arrays is None else len(self.arrays[0])

    def __iter__(self):
        rng = np.random.default_rng(self.seed + self._epoch)
        self._epoch += 1
        order = np.arange(self.samples)
        if self.shuffle:
            rng.shuffle(order)
        imgs, msks = [], []
        for i in order:
            if self.arrays is not None:
                img, msk = self.arrays[0][i], self.arrays[1][i]
            else:
                img, msk = load_pair(*self.pairs[i], self.size, self.area)
            if self.augment_mode:
                img, msk = augment(img, msk, rng, rich=self.augment_mode == "rich")
            imgs.append(img); msks.append(msk)
            if len(imgs) == self.batch_size:
                yield np.stack(imgs), np.stack(msks)
                imgs, msks = [], []
        if imgs:
            yield np.stack(imgs), np.stack(msks)


def synthetic_arrays(n: int, size: int, seed: int) -> Tuple[np.ndarray, np.ndarray]:
    """Random blobs: a disc of random centre/radius is the lesion; the image is the mask tinted plus noise."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    imgs = np.empty((n, size, size, 3), np.float32)
    msks = np.empty((n, size, size, 1), np.float32)
    for i in range(n):
        cy, cx = rng.uniform(0.3, 0.7, 2) * size
        r = rng.uniform(0.12, 0.3) * size
        m = (((yy - cy) ** 2 + (xx - cx) ** 2) < r * r).astype(np.float32)
        tint = rng.uniform(0.2, 0.8, 3).astype(np.float32)
        imgs[i] = np.clip(0.6 - 0.4 * m[..., None] * tint + 0.05 * rng.standard_normal((size, size, 3)).astype(np.float32), 0, 1)
        msks[i] = m[..., None]
    return imgs, msks
