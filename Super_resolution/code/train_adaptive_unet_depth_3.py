"""Fixed-depth variant: the adaptive trainer with the encoder depth pinned to 3
(mirror of /root/reference/Super_resolution/code/train_adaptive_unet_depth_3.py)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))

from train_adaptive_unet import parse_args, train  # noqa: E402

FIXED_DEPTH = 3

if __name__ == "__main__":
    args = parse_args()
    if args.depth_override not in (None, FIXED_DEPTH):
        print(f"[warn] overriding --depth_override={args.depth_override} with the fixed depth {FIXED_DEPTH}.")
    args.depth_override = FIXED_DEPTH
    train(args)
