"""Baseline (LayerNorm + Conv2DTranspose) segmentation U-Net trainer on the B200 kernels.

Mirror of /root/reference/Segmenation/code/unet_vinillia.py: ``build_unet(image_size, num_classes, base_channels,
depth)`` (:72-91: LayerNorm conv blocks, MaxPooling2D, Conv2DTranspose k2 s2, sigmoid / softmax head), the global
Dice metric (:94-99), BinaryCrossentropy training with ModelCheckpoint / EarlyStopping / ReduceLROnPlateau
(:263-290) and the same flags (:208-233).  ``--num_classes > 1`` trains the softmax head with categorical
cross-entropy (the BASELINE config-4 extrapolation).  Extra flags: ``--precision``,
``--synthetic N``.
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))

from dataset_paths import MODEL_ROOT, TRAIN_IMAGE_DIR, TRAIN_MASK_DIR, VALID_IMAGE_DIR, VALID_MASK_DIR  # noqa: E402

DEFAULT_IMAGE_SUFFIX = ".jpg"
DEFAULT_MASK_SUFFIX = "_segmentation.png"


def parse_args(argv=None) -> argparse.Namespace:
    p = argparse.ArgumentParser(description="Train a baseline U-Net on the ISIC-2017 dataset.")
    p.add_argument("--train_image_dir", type=Path, default=TRAIN_IMAGE_DIR, help="Directory of training images.")
    p.add_argument("--train_mask_dir", type=Path, default=TRAIN_MASK_DIR, help="Directory of training segmentation masks.")
    p.add_argument("--val_image_dir", type=Path, default=VALID_IMAGE_DIR, help="Directory of validation images.")
    p.add_argument("--val_mask_dir", type=Path, default=VALID_MASK_DIR, help="Directory of validation masks.")
    p.add_argument("--image_suffix", type=str, default=DEFAULT_IMAGE_SUFFIX, help="Suffix/pattern for image files.")
    p.add_argument("--mask_suffix", type=str, default=DEFAULT_MASK_SUFFIX, help="Suffix/pattern for mask files.")
    p.add_argument("--image_size", type=int, default=256, help="Square input resolution.")
    p.add_argument("--batch_size", type=int, default=8, help="Batch size.")
    p.add_argument("--epochs", type=int, default=60, help="Number of training epochs.")
    p.add_argument("--learning_rate", type=float, default=1e-4, help="Adam learning rate.")
    p.add_argument("--base_channels", type=int, default=32, help="Number of filters in the first encoder block.")
    p.add_argument("--depth", type=int, default=4, help="Depth of the encoder/decoder.")
    p.add_argument("--model_dir", type=Path, default=MODEL_ROOT, help="Directory to save checkpoints.")
    p.add_argument("--run_name", type=str, default="unet_isic", help="Prefix for saved checkpoints.")
    p.add_argument("--seed", type=int, default=13, help="Random seed for shuffling.")
    p.add_argument("--limit_train", type=int, default=None, help="Optional limit on number of training samples.")
    p.add_argument("--limit_val", type=int, default=None, help="Optional limit on number of validation samples.")
    p.add_argument("--augment", action="store_true", help="Enable simple geometric augmentations.")
    p.add_argument("--mixed_precision", action="store_true", help="Enable the 16-bit policy (bf16 on B200).")
    p.add_argument("--fit_verbose", type=int, choices=[0, 1, 2], default=2, help="Keras verbosity mode.")
    p.add_argument("--num_classes", type=int, default=1, help="1 = sigmoid/BCE (reference); >1 = softmax + categorical CE.")
    p.add_argument("--precision", choices=["fp32", "bf16"], default="fp32", help="Compute/storage precision.")
    p.add_argument("--synthetic", type=int, default=0, help="Train on this many random blob images instead of ISIC.")
    return p.parse_args(argv)


def _datasets(args):
    import numpy as np
    from b200unet.shared import seg_data as SD
    aug = "flip" if args.augment else None
    if args.synthetic:
        tr = SD.synthetic_arrays(args.synthetic, args.image_size, args.seed)
        va = SD.synthetic_arrays(max(args.synthetic // 4, args.batch_size), args.image_size, args.seed + 1)
        if args.num_classes > 1:   # class ids: background 0, lesion 1 + a pseudo-class from the image intensity
            def to_ids(imgs, msks):
                ids = (msks[..., 0] * (1 + (imgs.mean(-1) * (args.num_classes - 1)).astype(np.int64) % (args.num_classes - 1)))
                return imgs, ids.astype(np.int32)        # [N,S,S] class ids (sparse targets)
            tr, va = to_ids(*tr), to_ids(*va)
        return (SD.SegDataset([], args.image_size, args.batch_size, True, aug, args.seed, False, arrays=tr),
                SD.SegDataset([], args.image_size, args.batch_size, False, None, args.seed, False, arrays=va))
    for d, label in ((args.train_image_dir, "training images"), (args.train_mask_dir, "training masks"),
                     (args.val_image_dir, "validation images"), (args.val_mask_dir, "validation masks")):
        if not Path(d).expanduser().exists():
            raise FileNotFoundError(f"Missing {label} directory: {d}")
    sfx = lambda s: (Path(s).suffix.lower() or s.lower(),)
    tr_pairs = SD.collect_pairs(args.train_image_dir.expanduser(), args.train_mask_dir.expanduser(), sfx(args.image_suffix),
                                sfx(args.mask_suffix), args.limit_train)
    va_pairs = SD.collect_pairs(args.val_image_dir.expanduser(), args.val_mask_dir.expanduser(), sfx(args.image_suffix),
                                sfx(args.mask_suffix), args.limit_val)
    return (SD.SegDataset(tr_pairs, args.image_size, args.batch_size, True, aug, args.seed, area_resize=False),
            SD.SegDataset(va_pairs, args.image_size, args.batch_size, False, None, args.seed, area_resize=False))


def train(args: argparse.Namespace):
    from b200unet import builders as B
    from b200unet.keras import mixed_precision, set_random_seed
    from b200unet.keras.callbacks import EarlyStopping, ModelCheckpoint, ReduceLROnPlateau
    from b200unet.keras.losses import BinaryCrossentropy, CategoricalCrossentropy, global_dice_metric
    from b200unet.keras.optimizers import Adam

    precision = "bf16" if (args.mixed_precision or args.precision == "bf16") else "fp32"
    mixed_precision.set_global_policy("mixed_bfloat16" if precision == "bf16" else "float32")
    set_random_seed(args.seed)
    train_ds, val_ds = _datasets(args)
    print(f"Loaded {train_ds.samples} training samples and {val_ds.samples} validation samples.")
    model = B.build_unet(args.image_size, num_classes=args.num_classes, base_channels=args.base_channels, depth=args.depth)
    binary = args.num_classes == 1
    model.compile(optimizer=Adam(learning_rate=args.learning_rate),
                  loss=BinaryCrossentropy(global_dice=True, keras_metrics=True) if binary else CategoricalCrossentropy(),
                  metrics=[global_dice_metric] if binary else [])      # :94-99: Dice as one ratio over the batch
    args.model_dir.mkdir(parents=True, exist_ok=True)
    checkpoint_path = args.model_dir / f"{args.run_name}_best.keras"
    print(f"Checkpoints will be written to {checkpoint_path}")
    monitor, mode = ("val_dice_coefficient", "max") if binary else ("val_loss", "min")
    callbacks = [
        ModelCheckpoint(filepath=str(checkpoint_path), monitor=monitor, mode=mode, save_best_only=True, verbose=1),
        EarlyStopping(monitor=monitor, patience=10, mode=mode, restore_best_weights=True, verbose=1),
        ReduceLROnPlateau(monitor="val_loss", factor=0.5, patience=5, min_lr=1e-6, verbose=1),
    ]
    history = model.fit(train_ds, validation_data=val_ds, epochs=args.epochs, callbacks=callbacks, verbose=args.fit_verbose)
    final_path = args.model_dir / f"{args.run_name}_final.keras"
    model.save(str(final_path))
    return history


if __name__ == "__main__":
    train(parse_args())
