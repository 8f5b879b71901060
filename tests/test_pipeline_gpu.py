"""Device patch pipeline (csrc/pipeline.cu) through the C ABI: crop + 1/255, INTER_AREA shrink, INTER_CUBIC
enlargement and the HBM shuffle buffer, against (a) fixtures produced by the reference's own
random_patches / grid_patches / degrade_image, (b) the numpy restatement of OpenCV, (c) the host OpenCV
stream of shared/pipeline.py (same pairs, same order), and at full size through size-independent properties."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "pipeline_ref.npz")
TOL = 1e-5     # abs, fp32; hr crops are bit-exact


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _ds(P, scale):
    from b200unet.shared.pipeline import DevicePatchDataset
    return DevicePatchDataset([], P, scale, 4)


def test_crop_and_degrade_match_reference_fixtures(gold):
    from oracle import cv_resize_np as R
    img = gold["image_u8"]
    org = R.patch_origins(img.shape[0], img.shape[1], 64, 3, np.random.default_rng(1234))
    for scale in (0.25, 0.5, 0.7):
        lr, hr = _ds(64, scale).patches_of(img, org)
        assert np.array_equal(hr.cpu().numpy(), gold["rand_hr_p64"])
        assert np.abs(lr.cpu().numpy() - gold[f"rand_lr_p64_s{scale}"]).max() <= TOL
    lr, hr = _ds(48, 0.3).patches_of(img, R.grid_origins(img.shape[0], img.shape[1], 48, 60))
    assert np.array_equal(hr.cpu().numpy(), gold["grid_hr_p48_s60"])
    assert np.abs(lr.cpu().numpy() - gold["grid_lr_p48_s60_s0.3"]).max() <= TOL
    # float32 image source (already-normalised images) takes the other instantiation of the crop kernel
    lr2, hr2 = _ds(48, 0.3).patches_of(img.astype(np.float32) / np.float32(255.0),
                                        R.grid_origins(img.shape[0], img.shape[1], 48, 60))
    assert torch.equal(hr2, hr) and torch.equal(lr2, lr)


def test_degrade_matches_reference_on_out_of_range_patches(gold):
    for key in gold.files:
        if not key.startswith("deg_x_"):
            continue
        _, _, p, s = key.split("_")
        P, scale = int(p[1:]), float(s[1:])
        x = torch.from_numpy(gold[key]).cuda()
        lr = _ds(P, scale).degrade(x).cpu().numpy()
        ref = gold[key.replace("deg_x_", "deg_y_")]
        assert np.abs(lr - ref).max() <= TOL, key      # input clipped before the shrink, output NOT clipped
    assert any(gold[k].min() < 0 or gold[k].max() > 1 for k in gold.files if k.startswith("deg_y_"))


def test_origins_outside_the_image_are_clamped(gold):
    from b200unet import ops
    img = torch.from_numpy(gold["image_u8"]).cuda()
    org = torch.tensor([[-5, 1000], [149, -3]], dtype=torch.int32, device="cuda")
    hr = torch.empty((2, 32, 32, 3), dtype=torch.float32, device="cuda")
    ops.patch_extract(img, org, hr)
    ref = gold["image_u8"].astype(np.float32) / np.float32(255.0)
    got = hr.cpu().numpy()
    assert np.array_equal(got[0], ref[0:32, 201 - 32:201]) and np.array_equal(got[1], ref[150 - 32:150, 0:32])
    with pytest.raises(Exception):
        ops.patch_extract(img, org, torch.empty((2, 160, 160, 3), dtype=torch.float32, device="cuda"))


def test_full_size_batch_properties():
    """BASELINE config C2's input shape: 64 patches of 128x128, scale 0.25 (and the trainer's fixed 0.5)."""
    from oracle import cv_resize_np as R
    g = torch.Generator().manual_seed(3)
    a = torch.rand((64, 128, 128, 3), generator=g).cuda()
    b = torch.rand((64, 128, 128, 3), generator=g).cuda()
    for scale in (0.25, 0.5):
        ds = _ds(128, scale)
        la, lb = ds.degrade(a), ds.degrade(b)
        # linear on in-range inputs: D(a/2 + b/2) = D(a)/2 + D(b)/2
        lm = ds.degrade(0.5 * a + 0.5 * b)
        assert (lm - (0.5 * la + 0.5 * lb)).abs().max().item() <= 5e-6
        # both resizes have rows of weights summing to 1: a constant patch is a fixed point
        c = torch.full((2, 128, 128, 3), 0.37, device="cuda")
        assert (ds.degrade(c) - 0.37).abs().max().item() <= 2e-6
        # out-of-range values are clipped before the shrink
        assert torch.equal(ds.degrade(a * 3.0 - 1.0), ds.degrade((a * 3.0 - 1.0).clamp(0, 1)))
        # spot check against the OpenCV restatement
        ref = R.degrade_image(a[:2].cpu().numpy(), scale, 128)
        assert np.abs(la[:2].cpu().numpy() - ref).max() <= TOL


def test_copy_rows_gather_scatter():
    from b200unet import ops
    for row in (12, 7 * 7 * 3):                           # vectorised (16-byte) rows and the scalar path
        src = torch.arange(10 * row, dtype=torch.float32, device="cuda").reshape(10, row)
        dst = torch.zeros((6, row), dtype=torch.float32, device="cuda")
        si = torch.tensor([9, 0, 3], dtype=torch.int32, device="cuda")
        di = torch.tensor([5, 1, 2], dtype=torch.int32, device="cuda")
        ops.copy_rows(src, si, dst, di, 3)
        want = torch.zeros_like(dst)
        want[[5, 1, 2]] = src[[9, 0, 3]]
        assert torch.equal(dst, want)
        ops.copy_rows(src, None, dst, None, 6)
        assert torch.equal(dst, src[:6])


def _write_images(tmp_path, n, seed):
    import cv2
    rng = np.random.default_rng(seed)
    files = []
    for i in range(n):
        h, w = int(rng.integers(40, 70)), int(rng.integers(40, 90))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        path = str(tmp_path / f"img{i}.png")
        cv2.imwrite(path, img)
        files.append(path)
    return files


def test_device_dataset_equals_host_stream(tmp_path):
    """make_training_patch_dataset / make_eval_patch_dataset with device='cuda' yield the host stream's pairs, in
    its order (file shuffle, origin draws and shuffle-buffer draws share the numpy Generator discipline)."""
    pytest.importorskip("cv2")
    from b200unet.shared import pipeline as PL
    files = _write_images(tmp_path, 5, 0)
    for cap in (0, 7):
        host, n0 = PL.make_training_patch_dataset(files, 32, 3, 0.5, 4, seed=9, shuffle_buffer=cap)
        dev, n1 = PL.make_training_patch_dataset(files, 32, 3, 0.5, 4, seed=9, shuffle_buffer=cap, device="cuda")
        assert n0 == n1 == 15
        for i, ((hl, hh), (dl, dh)) in enumerate(zip(host, dev)):
            assert dl.is_cuda and dl.dtype == torch.float32 and tuple(dl.shape) == hl.shape
            assert np.array_equal(dh.cpu().numpy(), hh), (cap, i)
            assert np.abs(dl.cpu().numpy() - hl).max() <= TOL, (cap, i)
            if i == 11:
                break
    host, n0, lab0 = PL.make_eval_patch_dataset(files, 32, 0.5, 4, stride=20)
    dev, n1, lab1 = PL.make_eval_patch_dataset(files, 32, 0.5, 4, stride=20, device="cuda")
    hb, db = list(host), list(dev)
    assert n0 == n1 and lab0 == lab1 and len(hb) == len(db) and sum(b[0].shape[0] for b in db) == n0
    for (hl, hh), (dl, dh) in zip(hb, db):
        assert np.array_equal(dh.cpu().numpy(), hh) and np.abs(dl.cpu().numpy() - hl).max() <= TOL


def test_fit_on_device_dataset(tmp_path):
    pytest.importorskip("cv2")
    from b200unet import builders as B
    from b200unet.keras import clear_session, mixed_precision
    from b200unet.keras.optimizers import Adam
    from b200unet.shared import pipeline as PL
    clear_session()
    files = _write_images(tmp_path, 4, 1)
    for policy in ("float32", "mixed_bfloat16"):
        mixed_precision.set_global_policy(policy)
        model, _ = B.build_super_resolution_unet(0.5, depth_override=2, input_size=32)
        loss, metrics = B.build_losses_and_metrics("charbonnier")
        model.compile(optimizer=Adam(learning_rate=1e-3), loss=loss, metrics=metrics)
        dev, n = PL.make_training_patch_dataset(files, 32, 4, 0.5, 4, seed=2, shuffle_buffer=8, device="cuda")
        val, nv, _ = PL.make_eval_patch_dataset(files[:2], 32, 0.5, 4, device="cuda")
        hist = model.fit(dev, epochs=2, steps_per_epoch=n // 4, validation_data=val, verbose=0)
        assert len(hist.history["loss"]) == 2 and np.isfinite(hist.history["val_loss"]).all()
    mixed_precision.set_global_policy("float32")
    clear_session()
