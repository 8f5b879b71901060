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


def load_image_stack(directory, size: int, limit: Optional[int] = None) -> np.ndarray:
    """All images of a directory (natural order), INTER_AREA-resized to size x size, float32 RGB in [0, 1]: (N, size, size, 3)."""
    directory = Path(directory)
    names = sorted_alphanumeric([p.name for p in directory.iterdir() if p.is_file()])
    if limit is not None:
        names = names[:limit]
    images = [load_rgb_image(directory / name, size) for name in names]
    if not images:
        raise ValueError(f"No images found in {directory}")
    return np.stack(images, axis=0)


def load_rgb_image_full(path) -> np.ndarray:
    """Decode an image file to float32 RGB in [0, 1] at its native size."""
    cv2 = _cv2()
    bgr = cv2.imread(str(path), cv2.IMREAD_COLOR)
    if bgr is None:
        raise FileNotFoundError(f"Unable to read image: {path}")
    return cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB).astype(np.float32) / 255.0


def load_rgb_image(path, size: int) -> np.ndarray:
    """Decode, INTER_AREA-resize the uint8 image to size x size (rounded to integers, as the reference does: the resize
    comes BEFORE the float conversion, pipeline.py:60-67), then scale to [0, 1]."""
    cv2 = _cv2()
    bgr = cv2.imread(str(path), cv2.IMREAD_COLOR)
    if bgr is None:
        raise FileNotFoundError(f"Unable to read image: {path}")
    rgb = cv2.resize(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB), (size, size), interpolation=cv2.INTER_AREA)
    return rgb.astype(np.float32) / 255.0


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
                                batch_size: int, seed: int, shuffle_buffer: int = 1024, *, device=None):
    """Infinite stream of random (lr, hr) patch batches; returns (dataset, patches per epoch).

    ``device`` (not in the reference): a CUDA device puts crop / degrade / shuffle buffer on the GPU
    (``DevicePatchDataset``: same pairs, same order, CUDA tensors); None keeps the host (OpenCV) stream."""
    hr_files = list(hr_files)
    if not hr_files:
        raise ValueError("hr_files must contain at least one path.")
    if patches_per_image <= 0:
        raise ValueError("patches_per_image must be positive.")
    if device is not None:
        ds = DevicePatchDataset(hr_files, patch_size, scale, batch_size, per_image=patches_per_image, seed=seed,
                                shuffle_buffer=shuffle_buffer, infinite=True, device=device)
        return ds, len(hr_files) * patches_per_image
    ds = PatchDataset(lambda: _random_pairs(hr_files, patch_size, patches_per_image, scale, seed), batch_size,
                      shuffle_buffer, seed, infinite=True)
    return ds, len(hr_files) * patches_per_image


def make_eval_patch_dataset(hr_files: Sequence[str], patch_size: int, scale: float, batch_size: int, *,
                            stride: Optional[int] = None, device=None):
    """Finite stream of grid patches; returns (dataset, patch count, patch labels).  ``device``: as above."""
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
    if device is not None:
        ds = DevicePatchDataset(hr_files, patch_size, scale, batch_size, stride=stride, device=device)
    else:
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


# ------------------------------------------------------------------------------------------------
# The same streams with the patch work on the GPU.
#
# Host side: file order, PNG decode (cv2.imread -> uint8 RGB), the (top, left) draws of
# random_patches / grid_patches and the shuffle-buffer bookkeeping -- integers only, with the same
# numpy Generator discipline as the host mirror above, so both yield the SAME pairs in the SAME order.
# Device side (csrc/pipeline.cu): crop + 1/255 scaling, INTER_AREA shrink, INTER_CUBIC enlargement,
# and the shuffle buffer itself (a pool of patch pairs in HBM; batches are row gathers out of it).
# The uint8 image crosses PCIe once (a quarter of the float bytes); no float patch ever exists on the host.
# ------------------------------------------------------------------------------------------------
def load_rgb_image_u8(path) -> np.ndarray:
    """Decode to uint8 RGB; ``load_rgb_image_full`` is this / 255 in float32 (done by the crop kernel)."""
    cv2 = _cv2()
    bgr = cv2.imread(str(path), cv2.IMREAD_COLOR)
    if bgr is None:
        raise FileNotFoundError(f"Unable to read image: {path}")
    return np.ascontiguousarray(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))


def random_patch_origins(height: int, width: int, patch_size: int, count: int, rng: np.random.Generator) -> np.ndarray:
    """(top, left) of ``random_patches``: the draws ``random_patch`` makes, in its order."""
    if patch_size <= 0:
        raise ValueError("patch_size must be positive.")
    if count <= 0:
        raise ValueError("count must be positive.")
    if height < patch_size or width < patch_size:
        raise ValueError("patch_size exceeds image dimensions.")
    out = np.zeros((count, 2), dtype=np.int32)
    free_y, free_x = height - patch_size, width - patch_size
    for i in range(count):
        out[i, 0] = int(rng.integers(0, free_y + 1)) if free_y > 0 else 0
        out[i, 1] = int(rng.integers(0, free_x + 1)) if free_x > 0 else 0
    return out


def grid_patch_origins(height: int, width: int, patch_size: int, stride: Optional[int] = None) -> np.ndarray:
    """(top, left) of ``grid_patches`` (drop_remainder=False), row-major."""
    if patch_size <= 0:
        raise ValueError("patch_size must be positive.")
    if height < patch_size or width < patch_size:
        raise ValueError("patch_size exceeds image dimensions.")
    stride = stride or patch_size
    if stride <= 0:
        raise ValueError("stride must be positive.")
    pts = [(t, l) for t in range(0, height - patch_size + 1, stride) for l in range(0, width - patch_size + 1, stride)]
    if not pts:
        pts.append((height - patch_size, width - patch_size))
    return np.asarray(pts, dtype=np.int32).reshape(-1, 2)


class ShufflePlanner:
    """Index bookkeeping of ``PatchDataset._pairs``'s shuffle buffer, without the payloads.

    ``feed(n)`` announces n new pairs (tmp rows 0..n-1 of the current image) and returns
    ``(emits, stores)``: ``emits`` = ordered list of ("pool", slot) / ("tmp", row) sources to hand out,
    ``stores`` = list of (tmp row, pool slot) copies that must follow the emits.  ``flush()`` returns the
    pool slots left at the end of a finite stream, shuffled as the host mirror shuffles them."""

    def __init__(self, capacity: int, seed: int):
        self.capacity, self.rng, self.fill = int(capacity), np.random.default_rng(seed), 0

    def feed(self, n: int):
        if self.capacity <= 0:
            return [("tmp", i) for i in range(n)], []
        content = {}                      # slot -> tmp row holding its newest content
        emits = []
        for i in range(n):
            if self.fill < self.capacity:
                content[self.fill] = i
                self.fill += 1
                continue
            k = int(self.rng.integers(0, self.fill))
            emits.append(("tmp", content[k]) if k in content else ("pool", k))
            content[k] = i
        return emits, [(row, slot) for slot, row in content.items()]

    def flush(self):
        order = list(range(self.fill))
        self.rng.shuffle(order)
        self.fill = 0
        return [("pool", k) for k in order]


class DevicePatchDataset:
    """(lr, hr) batches as CUDA float32 tensors; drop-in for ``PatchDataset`` in ``Model.fit / evaluate``."""

    def __init__(self, files, patch_size: int, scale: float, batch_size: int, *, per_image: int = 0,
                 stride: Optional[int] = None, seed: int = 0, shuffle_buffer: int = 0, infinite: bool = False,
                 device=None):
        import torch
        from .. import ops
        if not 0 < scale < 1:
            raise ValueError("Scale must be between 0 and 1 for degradation.")
        if not torch.cuda.is_available():
            raise ops._ffi.B200Error("DevicePatchDataset needs a CUDA device: there is no CPU fallback "
                                     "(use make_training_patch_dataset(..., device=None) for the host mirror)")
        self.files, self.patch, self.scale, self.batch_size = list(files), int(patch_size), float(scale), int(batch_size)
        self.per_image, self.stride, self.seed = int(per_image), stride, int(seed)
        self.shuffle_buffer, self.infinite = int(shuffle_buffer), bool(infinite)
        self.device = torch.device(device if device is not None else "cuda")
        self.small = max(1, int(round(self.patch * self.scale)))
        self._area = ops.CvResizePlan(self.patch, self.small, ops.CV_INTER_AREA, self.device)
        self._cubic = ops.CvResizePlan(self.small, self.patch, ops.CV_INTER_CUBIC, self.device)
        self._epoch = 0
        self._pool = None

    def repeat(self):
        return DevicePatchDataset(self.files, self.patch, self.scale, self.batch_size, per_image=self.per_image,
                                  stride=self.stride, seed=self.seed, shuffle_buffer=self.shuffle_buffer, infinite=True,
                                  device=self.device)

    # -- one image -> (lr, hr) patch tensors on the device -------------------------------------------------
    def degrade(self, hr):
        """lr = INTER_CUBIC(INTER_AREA(clip(hr))) for a CUDA fp32 [n,P,P,3] tensor (degrade_image :79-94)."""
        import torch
        from .. import ops
        small = torch.empty((hr.shape[0], self.small, self.small, 3), dtype=torch.float32, device=hr.device)
        lr = torch.empty_like(hr)
        ops.gather2d(hr, small, self._area, self._area, clip01=True)
        ops.gather2d(small, lr, self._cubic, self._cubic, clip01=False)
        return lr

    def patches_of(self, image_u8: np.ndarray, origins: np.ndarray):
        import torch
        from .. import ops
        img = torch.from_numpy(np.ascontiguousarray(image_u8)).to(self.device)
        org = torch.from_numpy(np.ascontiguousarray(origins, dtype=np.int32)).to(self.device)
        hr = torch.empty((origins.shape[0], self.patch, self.patch, 3), dtype=torch.float32, device=self.device)
        ops.patch_extract(img, org, hr)
        return self.degrade(hr), hr

    def _image_stream(self):
        """Yields (image_u8, origins) in the order of ``_random_pairs`` / ``_grid_pairs``."""
        files = list(self.files)
        if self.per_image > 0:
            rng = np.random.default_rng(self.seed)
            while True:
                rng.shuffle(files)
                for path in files:
                    img = load_rgb_image_u8(path)
                    yield img, random_patch_origins(img.shape[0], img.shape[1], self.patch, self.per_image, rng)
        else:
            for path in files:
                img = load_rgb_image_u8(path)
                yield img, grid_patch_origins(img.shape[0], img.shape[1], self.patch, self.stride)

    def __iter__(self):
        import torch
        from .. import ops
        P, B, dev = self.patch, self.batch_size, self.device
        row = (P, P, 3)
        cap = self.shuffle_buffer
        if cap > 0 and self._pool is None:
            self._pool = (torch.empty((cap,) + row, dtype=torch.float32, device=dev),
                          torch.empty((cap,) + row, dtype=torch.float32, device=dev))
        cur = None        # [lr batch, hr batch, rows filled]

        def emit(sources, tmp):
            """Copy the listed sources into batches; yields every batch that fills up."""
            nonlocal cur
            pos = 0
            while pos < len(sources):
                if cur is None:
                    cur = [torch.empty((B,) + row, dtype=torch.float32, device=dev),
                           torch.empty((B,) + row, dtype=torch.float32, device=dev), 0]
                take = sources[pos:pos + (B - cur[2])]
                for kind, bufs in (("pool", self._pool), ("tmp", tmp)):
                    sel = [(j, s[1]) for j, s in enumerate(take) if s[0] == kind]
                    if not sel:
                        continue
                    ids = torch.tensor([[s for _, s in sel], [cur[2] + j for j, _ in sel]], dtype=torch.int32, device=dev)
                    for t in (0, 1):
                        ops.copy_rows(bufs[t], ids[0], cur[t], ids[1], len(sel))
                cur[2] += len(take)
                pos += len(take)
                if cur[2] == B:
                    out, cur = (cur[0], cur[1]), None
                    yield out

        while True:
            planner = ShufflePlanner(cap, self.seed + self._epoch)
            for img, origins in self._image_stream():
                tmp = self.patches_of(img, origins)
                emits, stores = planner.feed(origins.shape[0])
                yield from emit(emits, tmp)
                if stores:
                    ids = torch.tensor([[r for r, _ in stores], [s for _, s in stores]], dtype=torch.int32, device=dev)
                    for t in (0, 1):
                        ops.copy_rows(tmp[t], ids[0], self._pool[t], ids[1], len(stores))
            if cap > 0:
                yield from emit(planner.flush(), None)
            self._epoch += 1
            if not self.infinite:
                break
        if cur is not None and cur[2] > 0:
            yield cur[0][:cur[2]], cur[1][:cur[2]]
