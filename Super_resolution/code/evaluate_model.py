"""Offline evaluation of a trained adaptive-depth SR U-Net on the B200 kernels.

Mirror of /root/reference/Super_resolution/code/evaluate_model.py (load_checkpoint_model :57-91, evaluate
:94-163, write_outputs :173-190, CLI :193-213, main :216-286): same flags, same Y-channel metrics (BT.601 luma,
border shave ``2*round(1/scale)``, PSNR / SSIM / MS-SSIM / MSE, mean and std over patches) and the same report
files (``config.json``, ``metrics.json``, ``per_image_metrics.csv`` with the reference's column names).
The forward pass runs through ``b200unet`` (inference plan, CUDA graph).  Extra flags: ``--precision`` and
``--synthetic N`` (evaluate on N random images; with ``--random-init`` no checkpoint is needed -- smoke runs).
"""
import argparse
import csv
import glob
import json
import sys
from dataclasses import asdict, dataclass
from datetime import datetime
from pathlib import Path
from typing import Dict, List, Sequence, Tuple

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))

import numpy as np  # noqa: E402

from dataset_paths import HR_TRAIN_DIR, LOG_ROOT  # noqa: E402

try:
    from dataset_paths import HR_VALID_DIR  # noqa: E402
except ImportError:   # older dataset_paths: fall back to the training directory
    HR_VALID_DIR = HR_TRAIN_DIR

DEFAULT_OUTPUT_ROOT = Path(LOG_ROOT) / "evaluations"


@dataclass
class EvalResults:
    mse_mean: float
    mse_std: float
    psnr_mean: float
    psnr_std: float
    ssim_mean: float
    ssim_std: float
    msssim_mean: float
    msssim_std: float
    samples: int


def infer_eval_shave(scale: float, override) -> int:
    """Reference :42-54: explicit override (clamped at 0), else 2*round(1/scale)."""
    if override is not None:
        return max(0, int(override))
    if scale <= 0:
        return 0
    return 2 * int(round(1.0 / scale))


def load_checkpoint_model(model_path, scale: float, patch_size: int, depth_override=None, precision="fp32",
                          random_init=False):
    """Rebuild the architecture for this scale / patch size and load the checkpoint's weights (the reference
    first tries keras.models.load_model and falls back to exactly this rebuild + load_weights, :79-91)."""
    from b200unet import builders as B
    from b200unet.keras import mixed_precision
    mixed_precision.set_global_policy("mixed_bfloat16" if precision == "bf16" else "float32")
    model, info = B.build_super_resolution_unet(scale=scale, depth_override=depth_override, input_size=patch_size)
    if not random_init:
        if not Path(model_path).exists():
            raise FileNotFoundError(f"Model checkpoint not found: {model_path}")
        model.load_weights(str(model_path))
    return model


def evaluate(model, dataset, eval_shave: int) -> Tuple[EvalResults, List[Dict[str, float]]]:
    from b200unet import metrics as MT
    vals = {"psnr": [], "ssim": [], "msssim": [], "mse": []}
    per_image: List[Dict[str, float]] = []
    offset = 0
    for lr_batch, hr_batch in dataset:
        pred = model(lr_batch, training=False)        # clipped to [0,1] inside the fused luma kernel
        b = MT.eval_luma_metrics(pred, hr_batch, eval_shave)
        for k in vals:
            vals[k].append(b[k])
        for i in range(len(b["psnr"])):
            per_image.append({"index": offset + i, "psnr_y": float(b["psnr"][i]), "ssim_y": float(b["ssim"][i]),
                              "msssim_y": float(b["msssim"][i]), "mse_y": float(b["mse"][i])})
        offset += len(b["psnr"])
    if not per_image:
        raise RuntimeError("Evaluation dataset yielded no samples.")

    return summarise(vals, len(per_image)), per_image


def summarise(vals: Dict[str, List[np.ndarray]], samples: int) -> EvalResults:
    """Run-level statistics of the per-patch values (reference :146-163): float64 mean and POPULATION standard deviation."""
    def stats(v):
        arr = np.concatenate(v, axis=0).astype(np.float64)
        return float(np.mean(arr)), float(np.std(arr))

    (mse_m, mse_s), (ps_m, ps_s), (ss_m, ss_s), (ms_m, ms_s) = (stats(vals[k]) for k in ("mse", "psnr", "ssim", "msssim"))
    return EvalResults(mse_m, mse_s, ps_m, ps_s, ss_m, ss_s, ms_m, ms_s, samples)


def attach_filenames(per_image: List[Dict[str, float]], filenames: Sequence[str]) -> None:
    if len(per_image) != len(filenames):
        raise ValueError("Per-image metric count does not match filename list.")
    for item, name in zip(per_image, filenames):
        item["filename"] = name


def write_outputs(run_dir: Path, summary: EvalResults, per_image, config: Dict[str, object], write_per_image: bool) -> None:
    run_dir.mkdir(parents=True, exist_ok=True)
    (run_dir / "config.json").write_text(json.dumps(config, indent=2))
    (run_dir / "metrics.json").write_text(json.dumps(asdict(summary), indent=2))
    if write_per_image:
        with (run_dir / "per_image_metrics.csv").open("w", newline="") as handle:
            writer = csv.DictWriter(handle, fieldnames=["index", "filename", "psnr_y", "ssim_y", "msssim_y", "mse_y"])
            writer.writeheader()
            for row in per_image:
                writer.writerow(row)


def parse_args(argv=None) -> argparse.Namespace:
    p = argparse.ArgumentParser(description="Evaluate a trained adaptive-depth U-Net checkpoint.")
    p.add_argument("--model-path", type=Path, default=None, help="Path to the saved .keras checkpoint.")
    p.add_argument("--scale", type=float, required=True, help="Downscale factor used during training (0 < scale < 1).")
    p.add_argument("--hr-dir", type=Path, default=Path(HR_VALID_DIR), help="Directory of high-resolution images to evaluate.")
    p.add_argument("--patch-size", type=int, default=256, help="Patch size (matches training crops).")
    p.add_argument("--eval-stride", type=int, default=None, help="Stride used when tiling evaluation patches (default: patch size).")
    p.add_argument("--batch-size", type=int, default=8)
    p.add_argument("--limit", type=int, default=None, help="Optionally limit the number of evaluation samples.")
    p.add_argument("--eval-shave", type=int, default=None, help="Crop border pixels before metrics (mirrors training logic).")
    p.add_argument("--depth-override", type=int, default=None, help="Force a specific encoder depth when rebuilding the model.")
    p.add_argument("--output-dir", type=Path, default=DEFAULT_OUTPUT_ROOT, help="Directory to store evaluation reports.")
    p.add_argument("--run-name", type=str, default=None, help="Optional folder name inside --output-dir.")
    p.add_argument("--skip-per-image", action="store_true", help="Do not write per-patch CSV metrics.")
    p.add_argument("--use-train-split", action="store_true", help="Evaluate against the training split defaults instead of validation.")
    p.add_argument("--precision", choices=["fp32", "bf16"], default="fp32", help="Compute/storage precision.")
    p.add_argument("--synthetic", type=int, default=0, help="Evaluate on this many random images instead of --hr-dir.")
    p.add_argument("--host-pipeline", action="store_true",
                   help="Crop and degrade the evaluation patches on the host with OpenCV (default: on the GPU).")
    p.add_argument("--random-init", action="store_true", help="Skip the checkpoint (smoke runs of the forward path).")
    return p.parse_args(argv)


def main(argv=None) -> EvalResults:
    from b200unet.shared.pipeline import make_eval_patch_dataset, sorted_alphanumeric
    args = parse_args(argv)
    if args.model_path is None and not args.random_init:
        raise SystemExit("--model-path is required (or pass --random-init for a smoke run)")
    if args.synthetic:
        from train_adaptive_unet import _synthetic_dir
        hr_dir = _synthetic_dir(args.synthetic, max(2 * args.patch_size, 128), 1234)
    else:
        if args.use_train_split and args.hr_dir == Path(HR_VALID_DIR):
            args.hr_dir = Path(HR_TRAIN_DIR)
        hr_dir = Path(args.hr_dir).expanduser()
    if not hr_dir.exists():
        raise FileNotFoundError(f"High-resolution directory not found: {hr_dir}")
    hr_files = sorted_alphanumeric(glob.glob(str(hr_dir / "*.png")))
    if args.limit is not None and args.limit > 0:
        hr_files = hr_files[:args.limit]
    if not hr_files:
        raise ValueError(f"No high-resolution PNG files found in {hr_dir}")
    # the offline evaluator degrades by --scale (reference :233-239), unlike training's fixed 0.5
    eval_ds, total_patches, patch_labels = make_eval_patch_dataset(hr_files, patch_size=args.patch_size, scale=args.scale,
                                                                   batch_size=args.batch_size, stride=args.eval_stride,
                                                                   device=None if args.host_pipeline else "cuda")
    model = load_checkpoint_model(args.model_path.expanduser() if args.model_path else None, args.scale, args.patch_size,
                                  args.depth_override, args.precision, args.random_init)
    eval_shave = infer_eval_shave(args.scale, args.eval_shave)
    if eval_shave * 2 >= args.patch_size:
        eval_shave = max(0, args.patch_size // 2 - 1)
    summary, per_patch = evaluate(model, eval_ds, eval_shave=eval_shave)
    attach_filenames(per_patch, patch_labels)
    print(f"Evaluated {summary.samples} patches ({len(hr_files)} images).")
    print(f"  PSNR(Y):     {summary.psnr_mean:.4f} ± {summary.psnr_std:.4f} dB")
    print(f"  SSIM(Y):     {summary.ssim_mean:.4f} ± {summary.ssim_std:.4f}")
    print(f"  MS-SSIM(Y):  {summary.msssim_mean:.4f} ± {summary.msssim_std:.4f}")
    print(f"  MSE(Y):      {summary.mse_mean:.6f} ± {summary.mse_std:.6f}")
    timestamp = datetime.now().strftime("%Y%m%d-%H%M%S")
    run_dir = Path(args.output_dir).expanduser() / (args.run_name or f"scale{args.scale:.2f}_{timestamp}")
    config_payload = {
        "model_path": str(args.model_path.expanduser()) if args.model_path else None, "scale": args.scale,
        "hr_dir": str(hr_dir), "patch_size": args.patch_size, "eval_stride": args.eval_stride or args.patch_size,
        "batch_size": args.batch_size, "limit": args.limit, "eval_shave": eval_shave,
        "depth_override": args.depth_override, "samples": summary.samples, "images": len(hr_files),
        "created_at": timestamp,
    }
    write_outputs(run_dir, summary, per_patch, config_payload, write_per_image=not args.skip_per_image)
    print(f"[done] Report written to {run_dir}")
    return summary


if __name__ == "__main__":
    main()
