"""Patch pipeline, CPU side: the cv2 restatement (oracle/cv_resize_np.py) against cv2 itself and against the
fixtures the REFERENCE'S OWN degrade_image / random_patches / grid_patches produced
(tests/golden/make_pipeline_golden.py), the library's host tap tables against the oracle's (bit-exact), and the
shuffle-buffer bookkeeping of the device dataset against the host mirror's order."""
import os

import numpy as np
import pytest

from oracle import cv_resize_np as R

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pipeline_ref.npz")
TOL = 1e-5   # abs; cv2's SIMD paths fuse multiply-adds, the restatement does not (observed <= 2e-6)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_oracle_degrade_matches_reference_fixtures(gold):
    for key in gold.files:
        if not key.startswith("deg_x_"):
            continue
        _, _, p, s = key.split("_")
        P, scale = int(p[1:]), float(s[1:])
        y = R.degrade_image(gold[key], scale, P)
        ref = gold[key.replace("deg_x_", "deg_y_")]
        assert y.shape == ref.shape and y.dtype == np.float32
        assert np.abs(y - ref).max() <= TOL, key


def test_oracle_random_and_grid_patches_match_reference_fixtures(gold):
    img_u8 = gold["image_u8"]
    g = np.random.default_rng(1234)
    org = R.patch_origins(img_u8.shape[0], img_u8.shape[1], 64, 3, g)
    hr = R.crop(img_u8, org, 64)
    assert np.array_equal(hr, gold["rand_hr_p64"])              # same rng draws, same /255 scaling: bit-exact
    for scale in (0.25, 0.5, 0.7):
        lr = R.degrade_image(hr, scale, 64)
        assert np.abs(lr - gold[f"rand_lr_p64_s{scale}"]).max() <= TOL
    org = R.grid_origins(img_u8.shape[0], img_u8.shape[1], 48, 60)
    hr = R.crop(img_u8, org, 48)
    assert np.array_equal(hr, gold["grid_hr_p48_s60"])
    assert np.abs(R.degrade_image(hr, 0.3, 48) - gold["grid_lr_p48_s60_s0.3"]).max() <= TOL
    sub = img_u8[:40, :95]
    assert np.array_equal(R.crop(sub, R.grid_origins(40, 95, 40), 40), gold["grid_hr_p40"])


def test_oracle_matches_cv2_directly():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for P, s in [(128, 0.25), (128, 0.5), (128, 0.7), (128, 0.2), (64, 0.7), (256, 0.6), (96, 0.33), (37, 0.4), (17, 0.1)]:
        x = np.clip(rng.random((P, P, 3), dtype=np.float32) * 1.2 - 0.1, 0, 1).astype(np.float32)
        small = R.degraded_extent(P, s)
        a = cv2.resize(x, (small, small), interpolation=cv2.INTER_AREA)
        assert np.abs(a - R.resize(x, small, small, R.INTER_AREA)).max() <= TOL
        c = cv2.resize(a, (P, P), interpolation=cv2.INTER_CUBIC)
        assert np.abs(c - R.resize(a, P, P, R.INTER_CUBIC)).max() <= TOL
    x = rng.random((70, 45, 3), dtype=np.float32)     # non-square, both directions
    assert np.abs(cv2.resize(x, (20, 31), interpolation=cv2.INTER_AREA) - R.resize(x, 31, 20, R.INTER_AREA)).max() <= TOL
    assert np.abs(cv2.resize(x, (120, 131), interpolation=cv2.INTER_CUBIC) - R.resize(x, 131, 120, R.INTER_CUBIC)).max() <= TOL


def test_library_tap_tables_equal_oracle_tables():
    from b200unet import _ffi
    L = _ffi.load()
    area = [(128, 32), (128, 64), (128, 90), (128, 26), (128, 38), (64, 45), (256, 154), (96, 32), (37, 15), (50, 45),
            (17, 2), (128, 128), (5, 1), (1, 1)]
    cubic = area + [(32, 128), (26, 128), (45, 64), (2, 17), (1, 5), (90, 128)]
    for interp, cases in ((R.INTER_AREA, area), (R.INTER_CUBIC, cubic)):
        for i, o in cases:
            taps = L.b200_cv_resize_taps(i, o, interp)
            ri, rw = R.table(i, o, interp)
            assert taps == ri.shape[1], (interp, i, o)
            idx, w = np.zeros((o, taps), np.int32), np.zeros((o, taps), np.float32)
            assert L.b200_cv_resize_plan(i, o, interp, idx.ctypes.data, w.ctypes.data, taps) == 0
            assert np.array_equal(idx, ri) and np.array_equal(w, rw), (interp, i, o)
    assert L.b200_cv_resize_taps(32, 128, R.INTER_AREA) == -2          # enlarging with INTER_AREA: unsupported, loudly
    assert b"shrinking" in L.b200_last_error()
    assert L.b200_cv_resize_taps(0, 4, R.INTER_CUBIC) == -1


def test_origin_helpers_match_host_mirror():
    from b200unet.shared import pipeline as PL
    img = np.random.default_rng(3).random((90, 130, 3), dtype=np.float32)
    a = PL.random_patches(img, 32, 7, rng=np.random.default_rng(11))
    org = PL.random_patch_origins(90, 130, 32, 7, np.random.default_rng(11))
    assert np.array_equal(org, R.patch_origins(90, 130, 32, 7, np.random.default_rng(11)))
    assert np.array_equal(a, np.stack([img[t:t + 32, l:l + 32] for t, l in org]))
    # an image exactly one patch high draws only the column
    org = PL.random_patch_origins(32, 130, 32, 4, np.random.default_rng(5))
    ref = PL.random_patches(img[:32], 32, 4, rng=np.random.default_rng(5))
    assert np.array_equal(ref, np.stack([img[t:t + 32, l:l + 32] for t, l in org])) and (org[:, 0] == 0).all()
    g = PL.grid_patch_origins(90, 130, 32, 40)
    assert np.array_equal(PL.grid_patches(img, 32, stride=40), np.stack([img[t:t + 32, l:l + 32] for t, l in g]))
    assert np.array_equal(g, R.grid_origins(90, 130, 32, 40))
    with pytest.raises(ValueError):
        PL.random_patch_origins(20, 130, 32, 1, np.random.default_rng(0))
    with pytest.raises(ValueError):
        PL.grid_patch_origins(90, 130, 32, -1)


@pytest.mark.parametrize("cap,per_image,n_images", [(0, 3, 4), (5, 3, 6), (4, 9, 3), (16, 4, 3), (1, 2, 5)])
def test_shuffle_planner_reproduces_host_order(cap, per_image, n_images):
    """Drive the planner with pair ids and replay its emits/stores on an id pool: the order must equal
    PatchDataset._pairs' (same Generator, same draws), including the end-of-stream flush."""
    from b200unet.shared import pipeline as PL
    ids = [(i * per_image + j) for i in range(n_images) for j in range(per_image)]
    host = PL.PatchDataset(lambda: iter([(np.float32(k), np.float32(k)) for k in ids]), batch_size=4, shuffle_buffer=cap,
                           seed=21)
    want = [int(lr) for lr, _ in host._pairs()]
    planner, pool, got = PL.ShufflePlanner(cap, 21), {}, []
    for i in range(n_images):
        tmp = ids[i * per_image:(i + 1) * per_image]
        emits, stores = planner.feed(per_image)
        got += [pool[s] if kind == "pool" else tmp[s] for kind, s in emits]
        for row, slot in stores:
            pool[slot] = tmp[row]
    got += [pool[s] for _, s in planner.flush()]
    assert got == want and sorted(got) == ids


def test_device_dataset_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from b200unet import _ffi
    from b200unet.shared import pipeline as PL
    with pytest.raises(_ffi.B200Error):
        PL.make_training_patch_dataset(["x.png"], 32, 2, 0.5, 4, 0, device="cuda")


REF_PIPELINE = "/root/reference/shared/pipeline.py"


@pytest.mark.skipif(not os.path.exists(REF_PIPELINE), reason="reference checkout not present (build container only)")
def test_host_mirror_equals_the_reference_functions(tmp_path):
    """With the reference checkout at hand, run ITS pipeline functions (behind an empty tensorflow stub: the module only
    needs TF for the tf.data wrappers) next to the host mirror on the same files: identical arrays, identical order."""
    cv2 = pytest.importorskip("cv2")
    import importlib.util
    import sys
    import types
    from b200unet.shared import pipeline as PL
    stub = types.ModuleType("tensorflow")
    stub.data = types.SimpleNamespace(Dataset=object, AUTOTUNE=-1)
    had = sys.modules.get("tensorflow")
    sys.modules["tensorflow"] = stub
    try:
        spec = importlib.util.spec_from_file_location("ref_pipeline", REF_PIPELINE)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        if had is None:
            sys.modules.pop("tensorflow", None)
        else:
            sys.modules["tensorflow"] = had
    rng = np.random.default_rng(0)
    files = []
    for i in (1, 2, 10):
        path = str(tmp_path / f"im{i}.png")
        cv2.imwrite(path, rng.integers(0, 256, (50 + i, 70, 3), dtype=np.uint8))
        files.append(path)
    assert np.array_equal(ref.load_image_stack(tmp_path, 32), PL.load_image_stack(tmp_path, 32))
    assert np.array_equal(ref.load_rgb_image(files[1], 24), PL.load_rgb_image(files[1], 24))
    full = ref.load_rgb_image_full(files[2])
    assert np.array_equal(full, PL.load_rgb_image_full(files[2]))
    assert np.array_equal(PL.load_rgb_image_u8(files[2]).astype(np.float32) / 255.0, full)
    names = ["a10", "a2", "B1", "a1", "img_007x", "img_7y"]
    assert ref.sorted_alphanumeric(names) == PL.sorted_alphanumeric(names)
    assert np.array_equal(ref.degrade_image(full[:48, :48], 0.5, 48), PL.degrade_image(full[:48, :48], 0.5, 48))
    assert np.array_equal(ref.random_patches(full, 32, 5, rng=np.random.default_rng(4)),
                          PL.random_patches(full, 32, 5, rng=np.random.default_rng(4)))
    assert np.array_equal(ref.grid_patches(full, 32, stride=20), PL.grid_patches(full, 32, stride=20))
    # the generators behind the tf.data wrappers: same (lr, hr) pairs in the same order
    want = ref._iter_random_patch_pairs(files, 32, 2, 0.5, seed=3)
    got = PL._random_pairs(files, 32, 2, 0.5, 3)
    for _ in range(8):
        (wl, wh), (gl, gh) = next(want), next(got)
        assert np.array_equal(wl, gl) and np.array_equal(wh, gh)
    for (wl, wh), (gl, gh) in zip(ref._iter_grid_patch_pairs(files, 32, 20, 0.5), PL._grid_pairs(files, 32, 20, 0.5)):
        assert np.array_equal(wl, gl) and np.array_equal(wh, gh)


class _CpuDeviceDataset:
    """DevicePatchDataset with its three device calls replaced by CPU equivalents (oracle crop + degrade, index copies), so
    that the batching / shuffle-pool / epoch logic of __iter__ runs here and can be compared with the host stream."""

    @staticmethod
    def make(files, patch, scale, batch, monkeypatch, **kw):
        import torch
        from b200unet import ops
        from b200unet.shared import pipeline as PL
        ds = PL.DevicePatchDataset.__new__(PL.DevicePatchDataset)
        ds.files, ds.patch, ds.scale, ds.batch_size = list(files), patch, scale, batch
        ds.per_image, ds.stride, ds.seed = kw.get("per_image", 0), kw.get("stride"), kw.get("seed", 0)
        ds.shuffle_buffer, ds.infinite = kw.get("shuffle_buffer", 0), kw.get("infinite", False)
        ds.device, ds.small, ds._epoch, ds._pool = torch.device("cpu"), max(1, int(round(patch * scale))), 0, None

        def patches_of(image_u8, origins):
            hr = R.crop(image_u8, origins, patch)
            return torch.from_numpy(PL_degrade(hr)), torch.from_numpy(hr)

        def PL_degrade(hr):
            return np.stack([PL.degrade_image(p, scale, patch) for p in hr])      # the host stream's own cv2 degrade

        def copy_rows(src, src_rows, dst, dst_rows, n):
            s = src_rows.long() if src_rows is not None else torch.arange(n)
            d = dst_rows.long() if dst_rows is not None else torch.arange(n)
            dst[d] = src[s]
            return dst
        ds.patches_of = patches_of
        monkeypatch.setattr(ops, "copy_rows", copy_rows)
        return ds


@pytest.mark.parametrize("cap", [0, 5, 64])
def test_device_dataset_batching_logic_on_cpu(tmp_path, monkeypatch, cap):
    """Finite streams with a shuffle pool (fill, replace, end-of-stream flush), partial last batches, a second pass over the
    dataset (new shuffle epoch) and the infinite training stream: batch for batch equal to the host PatchDataset."""
    cv2 = pytest.importorskip("cv2")
    from b200unet.shared import pipeline as PL
    rng = np.random.default_rng(0)
    files = []
    for i in range(4):
        path = str(tmp_path / f"im{i}.png")
        cv2.imwrite(path, rng.integers(0, 256, (int(rng.integers(40, 60)), int(rng.integers(40, 70)), 3), dtype=np.uint8))
        files.append(path)
    P, scale, B = 16, 0.5, 5
    # finite grid stream through a shuffle pool: two passes (the second reshuffles with seed + 1)
    host = PL.PatchDataset(lambda: PL._grid_pairs(files, P, 24, scale), B, cap, seed=3)
    dev = _CpuDeviceDataset.make(files, P, scale, B, monkeypatch, stride=24, shuffle_buffer=cap, seed=3)
    for _ in range(2):
        hb, db = list(host), list(dev)
        assert len(hb) == len(db) > 2 and hb[-1][0].shape[0] == db[-1][0].shape[0] < B      # a ragged tail batch
        for (hl, hh), (dl, dh) in zip(hb, db):
            assert np.array_equal(hh, dh.numpy()) and np.array_equal(hl, dl.numpy())
    # infinite random stream (the training dataset)
    host, _ = PL.make_training_patch_dataset(files, P, 3, scale, B, seed=11, shuffle_buffer=cap)
    dev = _CpuDeviceDataset.make(files, P, scale, B, monkeypatch, per_image=3, shuffle_buffer=cap, seed=11, infinite=True)
    for i, ((hl, hh), (dl, dh)) in enumerate(zip(host, dev)):
        assert np.array_equal(hh, dh.numpy()) and np.array_equal(hl, dl.numpy()), i
        if i == 30:
            break
