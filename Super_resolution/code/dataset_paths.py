"""Default dataset / output locations (override with environment variables or CLI flags)."""
import os
from pathlib import Path

_ROOT = Path(os.environ.get("B200UNET_DATA_ROOT", Path(__file__).resolve().parents[1] / "data"))
HR_TRAIN_DIR = Path(os.environ.get("HR_TRAIN_DIR", _ROOT / "DIV2K_train_HR"))
LR_TRAIN_DIR = Path(os.environ.get("LR_TRAIN_DIR", _ROOT / "DIV2K_train_LR"))
HR_VALID_DIR = Path(os.environ.get("HR_VALID_DIR", _ROOT / "DIV2K_valid_HR"))
MODEL_ROOT = Path(os.environ.get("MODEL_ROOT", _ROOT / "models"))
LOG_ROOT = Path(os.environ.get("LOG_ROOT", _ROOT / "logs"))
