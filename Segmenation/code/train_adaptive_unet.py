"""Adaptive-depth segmentation U-Net trainer on the B200 kernels -- entry point with the reference's CLI.

Mirror of /root/reference/Segmenation/code/train_adaptive_unet.py: ``build_adaptive_depth_unet`` (:335-362,
BatchNorm conv blocks, MaxPooling2D, UpSampling2D bilinear, sigmoid head), the hybrid BCE+Dice losses
(:283-304), Dice / IoU metrics (:258-318), training protocols A / B (:382-403, CosineDecay for A), the
callbacks (:411-448), and the run artefacts ``config.json`` / ``model_summary.txt`` (:548-571).  Same flags
and defaults (:583-607).  Extra flags: ``--precision`` and ``--synthetic N`` (N random blob images, no dataset).
"""
import argparse
import json
import math
import sys
from dataclasses import dataclass
from datetime import datetime
from pathlib import Path
from typing import Callable, Dict, Optional

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))

from dataset_paths import (LOG_ROOT, MODEL_ROOT, TRAIN_IMAGE_DIR, TRAIN_MASK_DIR, VALID_IMAGE_DIR,  # noqa: E402
                           VALID_MASK_DIR)

DEFAULT_IMAGE_SIZE = 256
DEFAULT_BASE_CHANNELS = 64
DEFAULT_DEPTH = 4
DEFAULT_SEED = 42
DEFAULT_THRESHOLD = 0.5


@dataclass
class ProtocolConfig:
    key: str
    description: str
    loss_builder: Callable
    initial_lr: float
    epochs: int
    batch_size: int
    cosine_schedule: bool
    early_stopping_patience: Optional[int]


def _protocols() -> Dict[str, ProtocolConfig]:
    from b200unet.keras.losses import make_bce_dice_loss, make_hybrid_ce_dice_loss
    return {
        "A": ProtocolConfig("A", "MSCA-UNet hybrid loss (0.4·CE + 0.6·Dice) with cosine annealing",
                            lambda: make_hybrid_ce_dice_loss(alpha=0.4, beta=0.6), 1e-3, 100, 8, True, 15),
        "B": ProtocolConfig("B", "D2HU-Net BCE+Dice loss (0.5·BCE + 1.0·Dice)",
                            lambda: make_bce_dice_loss(bce_weight=0.5, dice_weight=1.0), 3e-4, 200, 16, False, None),
    }


PROTOCOL_KEYS = ("A", "B")


def prepare_callbacks(run_dir: Path, ckpt_path: Path, patience):
    from b200unet.keras.callbacks import BackupAndRestore, EarlyStopping, ModelCheckpoint, TensorBoard
    cbs = [ModelCheckpoint(filepath=str(ckpt_path), monitor="val_dice", mode="max", save_best_only=True, verbose=1),
           BackupAndRestore(str(run_dir / "train_backup")),
           TensorBoard(log_dir=str(run_dir), histogram_freq=0, write_graph=True, write_images=False, profile_batch=0)]
    if patience is not None and patience > 0:
        cbs.append(EarlyStopping(monitor="val_dice", mode="max", patience=patience, restore_best_weights=True, verbose=1))
    return cbs


def build_optimizer(protocol: ProtocolConfig, steps_per_epoch: int, epochs: int):
    from b200unet.keras.optimizers import Adam, CosineDecay
    if protocol.cosine_schedule:
        return Adam(learning_rate=CosineDecay(initial_learning_rate=protocol.initial_lr,
                                              decay_steps=epochs * max(steps_per_epoch, 1), alpha=0.0))
    return Adam(learning_rate=protocol.initial_lr)


def prepare_datasets(args, image_size, batch_size):
    from b200unet.shared import seg_data as SD
    if args.synthetic:
        tr = SD.synthetic_arrays(args.synthetic, image_size, args.seed)
        va = SD.synthetic_arrays(max(args.synthetic // 4, batch_size), image_size, args.seed + 1)
        return (SD.SegDataset([], image_size, batch_size, True, "rich", args.seed, arrays=tr),
                SD.SegDataset([], image_size, batch_size, False, None, args.seed, arrays=va), len(tr[0]), len(va[0]))
    dirs = [Path(p).expanduser() for p in (args.train_images or TRAIN_IMAGE_DIR, args.train_masks or TRAIN_MASK_DIR,
                                            args.val_images or VALID_IMAGE_DIR, args.val_masks or VALID_MASK_DIR)]
    tr_pairs = SD.collect_pairs(dirs[0], dirs[1], require_segmentation_token=True)
    va_pairs = SD.collect_pairs(dirs[2], dirs[3], require_segmentation_token=True)
    return (SD.SegDataset(tr_pairs, image_size, batch_size, True, "rich", args.seed),
            SD.SegDataset(va_pairs, image_size, batch_size, False, None, args.seed), len(tr_pairs), len(va_pairs))


def train(args: argparse.Namespace):
    from b200unet import builders as B
    from b200unet.keras import mixed_precision, set_random_seed
    from b200unet.keras.losses import dice_metric, iou_metric

    set_random_seed(args.seed)
    protocol = _protocols()[args.protocol]
    epochs = args.epochs or protocol.epochs
    batch_size = args.batch_size or protocol.batch_size
    image_size = args.image_size
    precision = "bf16" if (args.mixed_precision or args.precision == "bf16") else "fp32"
    mixed_precision.set_global_policy("mixed_bfloat16" if precision == "bf16" else "float32")

    train_ds, val_ds, train_count, val_count = prepare_datasets(args, image_size, batch_size)
    steps_per_epoch = math.ceil(train_count / batch_size)
    val_steps = math.ceil(val_count / batch_size)

    model = B.build_adaptive_depth_unet(input_size=image_size, base_channels=args.base_channels, depth=args.depth)
    model.compile(optimizer=build_optimizer(protocol, steps_per_epoch, epochs), loss=protocol.loss_builder(),
                  metrics=[dice_metric, iou_metric], jit_compile=False)
    summary_lines = []
    model.summary(print_fn=summary_lines.append)
    print("\n".join(summary_lines))

    model_dir = Path(args.model_dir or MODEL_ROOT).expanduser(); model_dir.mkdir(parents=True, exist_ok=True)
    log_root = Path(args.log_dir or LOG_ROOT).expanduser(); log_root.mkdir(parents=True, exist_ok=True)
    timestamp = datetime.now().strftime("%Y%m%d-%H%M%S")
    run_name = args.run_name or f"protocol{protocol.key}_seed{args.seed}_{timestamp}"
    run_dir = log_root / run_name; run_dir.mkdir(parents=True, exist_ok=True)
    ckpt_path = model_dir / f"{run_name}.keras"
    callbacks = prepare_callbacks(run_dir, ckpt_path,
                                  args.patience if args.patience is not None else protocol.early_stopping_patience)
    history = model.fit(train_ds, epochs=epochs, validation_data=val_ds, callbacks=callbacks, verbose=args.fit_verbose)
    eval_metrics = model.evaluate(val_ds, return_dict=True, verbose=1)
    config_payload = {
        "protocol": protocol.key, "description": protocol.description, "epochs_requested": epochs,
        "epochs_ran": len(history.history.get("loss", [])), "initial_lr": protocol.initial_lr, "batch_size": batch_size,
        "image_size": image_size, "train_samples": train_count, "val_samples": val_count,
        "train_steps_per_epoch": steps_per_epoch, "val_steps": val_steps, "seed": args.seed,
        "mixed_precision": precision == "bf16", "threshold": DEFAULT_THRESHOLD, "model_checkpoint": str(ckpt_path),
        "train_images": str(args.train_images or TRAIN_IMAGE_DIR), "train_masks": str(args.train_masks or TRAIN_MASK_DIR),
        "val_images": str(args.val_images or VALID_IMAGE_DIR), "val_masks": str(args.val_masks or VALID_MASK_DIR),
        "metrics": {k: float(v) for k, v in eval_metrics.items()},
    }
    (run_dir / "config.json").write_text(json.dumps(config_payload, indent=2))
    (run_dir / "model_summary.txt").write_text("\n".join(summary_lines))
    print("Validation metrics:")
    for key, value in eval_metrics.items():
        print(f"  {key}: {value:.4f}")
    return history, eval_metrics


def parse_args(argv=None) -> argparse.Namespace:
    p = argparse.ArgumentParser(description="Train Adaptive-Depth U-Net on ISIC-2017 segmentation.")
    p.add_argument("--protocol", type=str, choices=sorted(PROTOCOL_KEYS), default="A", help="Training protocol to follow.")
    p.add_argument("--epochs", type=int, default=0, help="Override epochs (0 keeps protocol default).")
    p.add_argument("--batch_size", type=int, default=0, help="Override batch size (0 keeps protocol default).")
    p.add_argument("--base_channels", type=int, default=DEFAULT_BASE_CHANNELS)
    p.add_argument("--depth", type=int, default=DEFAULT_DEPTH)
    p.add_argument("--image_size", type=int, default=DEFAULT_IMAGE_SIZE)
    p.add_argument("--seed", type=int, default=DEFAULT_SEED)
    p.add_argument("--patience", type=int, default=None, help="Override patience (None uses protocol default).")
    p.add_argument("--mixed_precision", action="store_true", help="Enable the 16-bit policy (bf16 on B200).")
    p.add_argument("--model_dir", type=str, default=str(MODEL_ROOT))
    p.add_argument("--log_dir", type=str, default=str(LOG_ROOT))
    p.add_argument("--run_name", type=str, default=None)
    p.add_argument("--train_images", type=str, default=None, help="Override training image directory.")
    p.add_argument("--train_masks", type=str, default=None, help="Override training mask directory.")
    p.add_argument("--val_images", type=str, default=None, help="Override validation image directory.")
    p.add_argument("--val_masks", type=str, default=None, help="Override validation mask directory.")
    p.add_argument("--precision", choices=["fp32", "bf16"], default="fp32", help="Compute/storage precision.")
    p.add_argument("--synthetic", type=int, default=0, help="Train on this many random blob images instead of ISIC.")
    p.add_argument("--fit_verbose", type=int, choices=[0, 1, 2], default=1)
    return p.parse_args(argv)


if __name__ == "__main__":
    train(parse_args())
