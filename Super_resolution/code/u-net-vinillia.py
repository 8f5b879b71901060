"""Legacy fixed-depth U-Net baseline for single-image super-resolution on the B200 kernels.

Mirror of /root/reference/Super_resolution/code/u-net-vinillia.py: whole images of ``--hr_size`` from a low-res and a
high-res directory (``load_image_stack`` :56-75), shuffled index split (:78-106), ``build_super_resolution_unet``
(:128-167: BatchNorm conv blocks, MaxPool2D, bilinear UpSampling2D + Conv3x3/ReLU, 3-channel sigmoid head), Adam,
EarlyStopping / ModelCheckpoint(``unet_vanilla_best.keras``) / BackupAndRestore (:268-272), and the RGB
PSNR / SSIM / MS-SSIM evaluation (:222-243) with the same flags (:292-305).

One documented deviation: the reference trains on ``MSE + 0.1 (1 - SSIM) + 0.01 VGG19-block4_conv4`` (:173-214), whose
VGG19 ImageNet weights are downloaded at run time; without network access this trainer optimises the MSE term alone
(``--loss mse``; ``combined`` fails loudly).  Extra flags: ``--precision``, ``--synthetic N``, ``--seed``, ``--limit``
(the reference reads ``args.seed`` / ``args.limit`` without declaring them).
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))

import numpy as np  # noqa: E402

PROJECT_ROOT = Path(__file__).resolve().parents[1]
DEFAULT_MODEL_DIR = PROJECT_ROOT / "models"


class ArrayDataset:
    """``tf.data.Dataset.from_tensor_slices(...).shuffle(len, seed, reshuffle_each_iteration=True).batch(B)`` (:109-120)."""

    def __init__(self, lr, hr, indices, batch_size, shuffle, seed):
        self.lr, self.hr, self.idx, self.bs = lr, hr, np.asarray(indices), int(batch_size)
        self.shuffle, self.seed, self._epoch = shuffle, seed, 0

    def __len__(self):
        return (len(self.idx) + self.bs - 1) // self.bs

    def __iter__(self):
        order = self.idx.copy()
        if self.shuffle:
            np.random.default_rng(self.seed + self._epoch).shuffle(order)
            self._epoch += 1
        for i in range(0, len(order), self.bs):
            sel = order[i:i + self.bs]
            yield self.lr[sel], self.hr[sel]


def _synthetic_stacks(n, size, seed):
    import cv2
    rng = np.random.default_rng(seed)
    hr = np.stack([np.clip(cv2.resize(rng.random((size // 8, size // 8, 3)).astype(np.float32), (size, size),
                                      interpolation=cv2.INTER_CUBIC), 0, 1) for _ in range(n)])
    small = max(1, size // 2)
    lr = np.stack([cv2.resize(cv2.resize(im, (small, small), interpolation=cv2.INTER_AREA), (size, size),
                              interpolation=cv2.INTER_CUBIC) for im in hr])
    return lr.astype(np.float32), hr.astype(np.float32)


def evaluate(model, dataset):
    """RGB metrics of one split: {"psnr": (mean, std), "ssim": ..., "ms_ssim": ...} (:222-243)."""
    from b200unet import metrics as MT
    vals = {"psnr": [], "ssim": [], "ms_ssim": []}
    for lr_b, hr_b in dataset:
        pred = model(lr_b, training=False).float().clamp(0.0, 1.0)
        vals["psnr"].append(MT.psnr_rgb(hr_b, pred))
        vals["ssim"].append(MT.ssim(hr_b, pred))
        vals["ms_ssim"].append(MT.ssim_multiscale(hr_b, pred))
    if not vals["psnr"]:
        return {}
    out = {}
    for k, v in vals.items():
        arr = np.concatenate(v, axis=0).astype(np.float64)
        out[k] = (float(np.mean(arr)), float(np.std(arr)))
    return out


def main(args: argparse.Namespace):
    from b200unet import builders as B
    from b200unet.keras import losses as LS, mixed_precision, set_random_seed
    from b200unet.keras.callbacks import BackupAndRestore, EarlyStopping, ModelCheckpoint
    from b200unet.keras.optimizers import Adam
    from b200unet.shared.pipeline import load_image_stack, split_indices

    if args.loss == "combined":
        raise NotImplementedError("loss 'combined' needs downloaded VGG19 ImageNet weights (no network here); use --loss mse")
    mixed_precision.set_global_policy("mixed_bfloat16" if args.precision == "bf16" else "float32")
    set_random_seed(args.seed)
    if args.synthetic:
        lr_images, hr_images = _synthetic_stacks(args.synthetic, args.hr_size, args.seed)
    else:
        hr_images = load_image_stack(Path(args.high_res_dir).expanduser(), args.hr_size, limit=args.limit)
        lr_images = load_image_stack(Path(args.low_res_dir).expanduser(), args.hr_size, limit=args.limit)
    if hr_images.shape != lr_images.shape:
        raise ValueError("High-resolution and low-resolution stacks must align one-to-one.")
    tr, va, te = split_indices(hr_images.shape[0], args.train_split, args.val_split, args.test_split, args.seed)
    train_ds = ArrayDataset(lr_images, hr_images, tr, args.batch_size, True, args.seed)
    val_ds = ArrayDataset(lr_images, hr_images, va, args.batch_size, False, args.seed)
    test_ds = ArrayDataset(lr_images, hr_images, te, args.batch_size, False, args.seed) if len(te) else None

    model = B.build_vanilla_super_resolution_unet((args.hr_size, args.hr_size, 3))
    model.compile(optimizer=Adam(learning_rate=args.learning_rate), loss=LS.SRLoss("mse"), metrics=[LS.PSNRMetric()])
    model_dir = Path(args.model_dir).expanduser()
    model_dir.mkdir(parents=True, exist_ok=True)
    callbacks = [
        EarlyStopping(monitor="val_loss", mode="min", patience=args.patience, restore_best_weights=True, verbose=1),
        ModelCheckpoint(filepath=str(model_dir / "unet_vanilla_best.keras"), monitor="val_loss", mode="min",
                        save_best_only=True, verbose=1),
        BackupAndRestore(str(model_dir / "train_backup")),
    ]
    history = model.fit(train_ds, epochs=args.epochs, validation_data=val_ds, callbacks=callbacks, verbose=2)
    results = {"validation": evaluate(model, val_ds)}
    print("Validation metrics:", results["validation"])
    if test_ds is not None:
        results["test"] = evaluate(model, test_ds)
        print("Test metrics:", results["test"])
    return history, results


def parse_args(argv=None) -> argparse.Namespace:
    p = argparse.ArgumentParser(description="Train the vanilla super-resolution U-Net baseline.")
    p.add_argument("--high_res_dir", type=str, default=None, help="Directory containing high-resolution images.")
    p.add_argument("--low_res_dir", type=str, default=None, help="Directory containing low-resolution images.")
    p.add_argument("--hr_size", type=int, default=256, help="Input/output spatial size.")
    p.add_argument("--batch_size", type=int, default=4)
    p.add_argument("--epochs", type=int, default=100)
    p.add_argument("--learning_rate", type=float, default=1e-4)
    p.add_argument("--patience", type=int, default=10)
    p.add_argument("--train_split", type=float, default=0.8, help="Relative portion of samples for training.")
    p.add_argument("--val_split", type=float, default=0.1, help="Relative portion of samples for validation.")
    p.add_argument("--test_split", type=float, default=0.1, help="Relative portion of samples for testing.")
    p.add_argument("--model_dir", type=str, default=str(DEFAULT_MODEL_DIR), help="Directory to store checkpoints.")
    p.add_argument("--seed", type=int, default=1234)
    p.add_argument("--limit", type=int, default=None)
    p.add_argument("--loss", choices=["mse", "combined"], default="mse")
    p.add_argument("--precision", choices=["fp32", "bf16"], default="fp32", help="Compute/storage precision.")
    p.add_argument("--synthetic", type=int, default=0, help="Train on this many random images instead of the directories.")
    args = p.parse_args(argv)
    if not args.synthetic and (args.high_res_dir is None or args.low_res_dir is None):
        p.error("--high_res_dir and --low_res_dir are required (or use --synthetic N)")
    return args


if __name__ == "__main__":
    main(parse_args())
