"""Default ISIC-2017 dataset / output locations (override with environment variables or CLI flags).

Same constant names as /root/reference/Segmenation/code/dataset_paths.py; the reference hard-codes a site path."""
import os
from pathlib import Path

DATA_ROOT = Path(os.environ.get("B200UNET_ISIC_ROOT", Path(__file__).resolve().parents[1] / "data" / "ISIC-2017"))
TRAIN_IMAGE_DIR = DATA_ROOT / "ISIC-2017_Training_Data"
TRAIN_MASK_DIR = DATA_ROOT / "ISIC-2017_Training_Part1_GroundTruth"
VALID_IMAGE_DIR = DATA_ROOT / "ISIC-2017_Validation_Data"
VALID_MASK_DIR = DATA_ROOT / "ISIC-2017_Validation_Part1_GroundTruth"
TEST_IMAGE_DIR = DATA_ROOT / "ISIC-2017_Test_v2_Data"
TEST_MASK_DIR = DATA_ROOT / "ISIC-2017_Test_v2_Part1_GroundTruth"
MODEL_ROOT = Path(os.environ.get("MODEL_ROOT", Path(__file__).resolve().parents[1] / "models"))
LOG_ROOT = Path(os.environ.get("LOG_ROOT", Path(__file__).resolve().parents[1] / "logs" / "tensorboard"))
VISUAL_ROOT = Path(__file__).resolve().parents[1] / "scale_visualizations"
