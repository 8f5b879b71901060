"""Adaptive-depth SR U-Net trainer on the B200 kernels -- entry point with the reference's CLI.

Mirror of /root/reference/Super_resolution/code/train_adaptive_unet.py (train :380-722, parse_args
:725-804): same flags and defaults, same config.json / model_summary.txt / checkpoint file names, same
Keras-style epoch log lines and the same post-training evaluation (Y-channel MSE / PSNR / SSIM / MS-SSIM
with a `2*round(1/scale)` border shave).  The model, losses and train step run through ``b200unet``.
Extra flags: ``--precision {fp32,bf16}`` and ``--synthetic N`` (train on N random images, no dataset).
Data parallel: launch under ``python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 ...``; every rank
reads the same patch stream and trains on ITS contiguous slice of each global batch (``--batch_size`` is the GLOBAL
batch and must divide by N), gradients are exchanged inside ``Model.distribute()``; rank 0 writes the artefacts.
"""
import argparse
import glob
import json
import math
import os
import sys
from datetime import datetime
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))

import numpy as np  # noqa: E402

from dataset_paths import HR_TRAIN_DIR, LOG_ROOT, LR_TRAIN_DIR, MODEL_ROOT  # noqa: E402

DEFAULT_HR_SIZE = 256
DEFAULT_BASE_CHANNELS = 64
DEFAULT_RESIDUAL_HEAD_CHANNELS = 64
DATA_LR_SHRINK = 0.5   # the LR input is always a x0.5 degradation, whatever --scale is (reference :60, :438)
IMAGE_SUFFIX = ".png"


def _require(cond, msg, exc=ValueError):
    if not cond:
        raise exc(msg)


def _synthetic_dir(n, size, seed):
    import tempfile
    import cv2
    d = Path(tempfile.mkdtemp(prefix="b200unet_synth_"))
    rng = np.random.default_rng(seed)
    for i in range(n):
        low = cv2.resize(rng.random((size // 8, size // 8, 3)).astype(np.float32), (size, size), interpolation=cv2.INTER_CUBIC)
        cv2.imwrite(str(d / f"{i:04d}.png"), (np.clip(low, 0, 1) * 255).astype(np.uint8))
    return d


class _ShardedBatches:
    """Re-iterable view of a (lr, hr) batch stream that yields this rank's slice of every full global batch."""

    def __init__(self, ds, rank, world):
        self.ds, self.rank, self.world = ds, rank, world

    def __iter__(self):
        from b200unet.parallel import shard_range
        for lr, hr in self.ds:
            if lr.shape[0] % self.world:
                continue                      # ragged tail batch: equal shards are required (mean of shard means)
            lo, hi = shard_range(lr.shape[0], self.rank, self.world)
            yield lr[lo:hi], hr[lo:hi]


def _init_distributed():
    """(rank, world) from the torchrun environment; initialises NCCL when world > 1."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist.get_rank(), world


def train(args: argparse.Namespace) -> None:
    import torch
    rank, world = _init_distributed()
    from b200unet import builders as B, metrics as MT
    from b200unet.keras import mixed_precision, set_random_seed
    from b200unet.keras.callbacks import BackupAndRestore, EarlyStopping, ModelCheckpoint, TensorBoard
    from b200unet.keras.optimizers import Adam
    from b200unet.shared.pipeline import (make_eval_patch_dataset, make_training_patch_dataset, sorted_alphanumeric,
                                          split_indices)

    P = args.patch_size
    _require(P > 0, "patch_size must be a positive integer.")
    _require(args.patches_per_image > 0, "patches_per_image must be positive.")
    _require(args.eval_stride is None or args.eval_stride > 0, "eval_stride must be positive when provided.")
    _require(args.shuffle_buffer >= 0, "shuffle_buffer must be non-negative.")
    _require(args.max_depth >= 1, "max_depth must be at least 1.")
    _require(args.initial_epoch >= 0, "initial_epoch must be non-negative.")
    _require(args.initial_epoch < args.epochs, "initial_epoch must be smaller than --epochs to resume training.")

    if args.synthetic:
        hr_dir = _synthetic_dir(args.synthetic, max(2 * P, 128), args.seed)
    else:
        hr_dir = Path(args.high_res_dir or HR_TRAIN_DIR).expanduser()
    _require(hr_dir.exists(), f"High-resolution directory not found: {hr_dir}", FileNotFoundError)
    hr_paths = sorted_alphanumeric(glob.glob(str(hr_dir / f"*{IMAGE_SUFFIX}")))
    if args.limit and args.limit > 0:
        hr_paths = hr_paths[:args.limit]
    _require(hr_paths, "No high-resolution images found with the given suffix.")
    if args.low_res_dir:
        print("[info] --low_res_dir is ignored in patch mode; LR patches are generated on the fly.")
    train_frac = 1.0 - (args.val_split + args.test_split)
    _require(train_frac > 0, "Validation and test splits leave no room for training data.")
    tr, va, te = split_indices(len(hr_paths), train_frac, args.val_split, args.test_split, args.seed)
    train_paths, val_paths, test_paths = ([hr_paths[i] for i in idx] for idx in (tr, va, te))

    # crop / degrade / shuffle buffer on the GPU (same pairs, same order as the host OpenCV stream); --host_pipeline
    # or B200_HOST_PIPELINE=1 keeps the reference's host-side OpenCV stream
    pipe_dev = None if (args.host_pipeline or os.environ.get("B200_HOST_PIPELINE") == "1") else "cuda"
    eval_ds = lambda paths: make_eval_patch_dataset(paths, patch_size=P, scale=DATA_LR_SHRINK,
                                                    batch_size=args.batch_size, stride=args.eval_stride, device=pipe_dev)
    train_ds, n_train = make_training_patch_dataset(train_paths, patch_size=P, patches_per_image=args.patches_per_image,
                                                    scale=DATA_LR_SHRINK, batch_size=args.batch_size, seed=args.seed,
                                                    shuffle_buffer=args.shuffle_buffer, device=pipe_dev)
    val_fit_ds, n_val = None, 0
    if val_paths:
        val_fit_ds, n_val, _ = eval_ds(val_paths)
    n_test = eval_ds(test_paths)[1] if test_paths else 0
    steps_per_epoch = math.ceil(n_train / args.batch_size)
    _require(steps_per_epoch > 0, "Training dataset produced zero patches. Check patches_per_image or dataset splits.")
    val_steps = math.ceil(n_val / args.batch_size) if n_val else None

    precision = "bf16" if (args.mixed_precision or args.precision == "bf16") else "fp32"
    mixed_precision.set_global_policy("mixed_bfloat16" if precision == "bf16" else "float32")
    set_random_seed(args.seed)
    model, info = B.build_super_resolution_unet(scale=args.scale, base_channels=DEFAULT_BASE_CHANNELS,
                                                residual_head_channels=DEFAULT_RESIDUAL_HEAD_CHANNELS,
                                                depth_override=args.depth_override, input_size=P,
                                                max_depth=args.max_depth)
    loss_fn, metrics = B.build_losses_and_metrics(args.loss)
    model.compile(optimizer=Adam(learning_rate=args.learning_rate), loss=loss_fn, metrics=metrics, jit_compile=False)
    if world > 1:
        _require(args.batch_size % world == 0, f"--batch_size {args.batch_size} must divide by the {world} ranks")
        model.distribute()
        train_ds = _ShardedBatches(train_ds, rank, world)
        if val_fit_ds is not None:
            val_fit_ds = _ShardedBatches(val_fit_ds, rank, world)

    if args.resume_from:
        cand = Path(args.resume_from).expanduser()
        if cand.is_dir():
            ckpts = sorted(cand.glob("*.keras"), key=lambda p: p.stat().st_mtime, reverse=True)
            _require(ckpts, f"--resume_from directory {cand} contains no '.keras' checkpoints.", FileNotFoundError)
            cand = ckpts[0]
        _require(cand.exists(), f"Checkpoint not found: {cand}", FileNotFoundError)
        print(f"[info] Loading weights from {cand}")
        model.load_weights(str(cand))
        if args.initial_epoch == 0:
            print("[warn] --resume_from supplied without --initial_epoch; training will restart from epoch 0.")
    elif args.initial_epoch > 0:
        print("[warn] --initial_epoch was set without --resume_from; training will skip the initial epochs but "
              "start from random weights.")

    summary = []
    model.summary(print_fn=summary.append)
    print("\n".join(summary))

    model_dir = Path(args.model_dir).expanduser()
    if rank > 0:      # weight reads are collective under the sharded optimizer: every rank saves, rank 0's files are canonical
        model_dir = model_dir / f"rank{rank}"
        args.log_dir = str(Path(args.log_dir).expanduser() / f"rank{rank}")
    model_dir.mkdir(parents=True, exist_ok=True)
    ckpt_path = model_dir / f"unet_adaptive_scale_new_loss{args.scale:.2f}_depth{info['depth']}.keras"
    stamp = datetime.now().strftime("%Y%m%d-%H%M%S")
    run_dir = Path(args.log_dir).expanduser() / (args.run_name or
                                                 f"scale{args.scale:.2f}_bs{args.batch_size}_lr{args.learning_rate:.0e}_{stamp}")
    run_dir.mkdir(parents=True, exist_ok=True)
    config = {
        "scale": args.scale, "depth": info["depth"], "max_depth": args.max_depth, "patch_size": P,
        "patches_per_image": args.patches_per_image, "eval_stride": args.eval_stride or P,
        "base_channels": DEFAULT_BASE_CHANNELS, "residual_head_channels": DEFAULT_RESIDUAL_HEAD_CHANNELS,
        "learning_rate": args.learning_rate, "batch_size": args.batch_size, "epochs": args.epochs,
        "patience": args.patience, "train_images": len(train_paths), "val_images": len(val_paths),
        "test_images": len(test_paths), "train_patches_per_epoch": int(n_train), "val_patches": int(n_val),
        "test_patches": int(n_test), "steps_per_epoch": int(steps_per_epoch),
        "validation_steps": int(val_steps) if val_steps is not None else None,
        "mixed_precision": precision == "bf16", "high_res_dir": str(hr_dir), "low_res_mode": "synthetic_patches",
        "model_dir": str(model_dir), "log_dir": str(run_dir), "created_at": stamp,
    }
    (run_dir / "config.json").write_text(json.dumps(config, indent=2))
    (run_dir / "model_summary.txt").write_text("\n".join(summary))

    callbacks = [
        EarlyStopping(monitor="val_loss", patience=args.patience, restore_best_weights=True, verbose=1),
        ModelCheckpoint(filepath=str(ckpt_path), monitor="val_loss", save_best_only=True, verbose=1),
        BackupAndRestore(str(run_dir / "train_backup")),
        TensorBoard(log_dir=str(run_dir), histogram_freq=0, write_graph=False, write_images=False, profile_batch=0,
                    update_freq="epoch"),
    ]
    history = model.fit(train_ds, epochs=args.epochs, initial_epoch=args.initial_epoch, steps_per_epoch=steps_per_epoch,
                        validation_data=val_fit_ds, validation_steps=val_steps if val_fit_ds is not None else None,
                        validation_freq=1, callbacks=callbacks, verbose=2)
    print("Training complete.")
    print(f"Model info: {info}")
    print(f"Checkpoint saved to: {ckpt_path}")

    if args.eval_shave is not None:
        shave = max(0, int(args.eval_shave))
    else:
        shave = 2 * int(round(1.0 / args.scale)) if args.scale > 0 else 0
    if shave * 2 >= P > 0:
        adjusted = max(0, P // 2 - 1)
        print(f"[warn] eval_shave={shave} removes the full frame for hr_size={P}; reducing to {adjusted} pixels.")
        shave = adjusted
    for name, paths in (("Validation", val_paths), ("Test", test_paths)):
        if not paths:
            continue
        ds, _, _ = eval_ds(paths)
        acc = {"mse": [], "psnr": [], "ssim": [], "msssim": []}
        seen = 0
        for lr_b, hr_b in ds:
            b = MT.eval_luma_metrics(model(lr_b, training=False), hr_b, shave)
            for k in ("psnr", "ssim", "msssim", "mse"):
                acc[k].append(b[k])
            seen += hr_b.shape[0]
        if not seen:
            print(f"{name}: no samples, skipping metric aggregation.")
            continue
        ms = {k: (float(np.mean(np.concatenate(v).astype(np.float64))), float(np.std(np.concatenate(v).astype(np.float64))))
              for k, v in acc.items()}
        print(f"{name} patches evaluated: {seen}")
        print(f"  MSE(Y)     : {ms['mse'][0]:.6f} ± {ms['mse'][1]:.6f}")
        print(f"  PSNR(Y)    : {ms['psnr'][0]:.4f} ± {ms['psnr'][1]:.4f} dB")
        print(f"  SSIM(Y)    : {ms['ssim'][0]:.4f} ± {ms['ssim'][1]:.4f}")
        print(f"  MS-SSIM(Y) : {ms['msssim'][0]:.4f} ± {ms['msssim'][1]:.4f}")
    return history


def parse_args(argv=None) -> argparse.Namespace:
    p = argparse.ArgumentParser(description="Train adaptive-depth U-Net for super-resolution.")
    p.add_argument("--scale", type=float, required=True, help="Downscale factor (0 < scale < 1).")
    p.add_argument("--batch_size", type=int, default=4)
    p.add_argument("--epochs", type=int, default=100)
    p.add_argument("--learning_rate", type=float, default=1e-4)
    p.add_argument("--loss", type=str, default="charbonnier", choices=["charbonnier", "l1", "combined"],
                   help="Training loss to optimise. Use 'charbonnier' or 'l1' for PSNR-focused benchmarks.")
    p.add_argument("--patience", type=int, default=10)
    p.add_argument("--val_split", type=float, default=0.1)
    p.add_argument("--test_split", type=float, default=0.1)
    p.add_argument("--limit", type=int, default=None, help="Optionally limit the number of training samples.")
    p.add_argument("--seed", type=int, default=1234, help="Seed for dataset shuffling and splitting.")
    p.add_argument("--patch_size", type=int, default=DEFAULT_HR_SIZE, help="Side length for HR/LR training patches.")
    p.add_argument("--patches_per_image", type=int, default=4,
                   help="Random patches sampled per image for each training epoch.")
    p.add_argument("--eval_stride", type=int, default=None,
                   help="Stride to use when tiling evaluation patches (defaults to patch_size).")
    p.add_argument("--shuffle_buffer", type=int, default=1024, help="Shuffle buffer size for the training patch dataset.")
    p.add_argument("--preview_patches", type=int, default=3, help="Accepted for compatibility (TensorBoard previews).")
    p.add_argument("--eval_shave", type=int, default=None,
                   help="Pixels to trim from each border before computing PSNR/SSIM (default: 2 * round(1 / scale)).")
    p.add_argument("--depth_override", type=int, default=None, help="Force a specific encoder depth.")
    p.add_argument("--max_depth", type=int, default=7,
                   help="Maximum encoder depth when inferring from scale (ignored if --depth_override is set).")
    p.add_argument("--mixed_precision", action="store_true", help="Enable the 16-bit policy (bf16 on B200).")
    p.add_argument("--precision", choices=["fp32", "bf16"], default="fp32", help="Compute/storage precision.")
    p.add_argument("--model_dir", type=str, default=str(MODEL_ROOT), help="Directory to store checkpoints.")
    p.add_argument("--log_dir", type=str, default=str(LOG_ROOT), help="Directory to store TensorBoard logs.")
    p.add_argument("--run_name", type=str, default=None, help="Optional explicit run name for TensorBoard.")
    p.add_argument("--high_res_dir", type=str, default=None, help="Override the high-resolution dataset directory.")
    p.add_argument("--low_res_dir", type=str, default=None, help="Ignored in patch mode.")
    p.add_argument("--resume_from", type=str, default=None,
                   help="Optional path to a .keras checkpoint (or directory containing checkpoints) to resume from.")
    p.add_argument("--initial_epoch", type=int, default=0,
                   help="Epoch index to begin training from when resuming (must be < --epochs).")
    p.add_argument("--synthetic", type=int, default=0, help="Train on this many random images instead of a dataset.")
    p.add_argument("--host_pipeline", action="store_true",
                   help="Crop / degrade / shuffle patches on the host with OpenCV as the reference does "
                        "(default: on the GPU, same pairs in the same order; also B200_HOST_PIPELINE=1).")
    return p.parse_args(argv)


if __name__ == "__main__":
    train(parse_args())
