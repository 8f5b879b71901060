"""The reference's train / eval entry points (same CLIs) run end to end on the B200 kernels with tiny
synthetic datasets: Super_resolution/code/{train_adaptive_unet,train_adaptive_unet_depth_3,evaluate_model}.py and
Segmenation/code/{train_adaptive_unet,unet_vinillia}.py.  Checks the run artefacts the reference's tooling reads
(config.json, model_summary.txt, checkpoint names, metrics.json / per_image_metrics.csv, Keras-style epoch lines)."""
import importlib.util
import json
import os
import re
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(rel, name):
    path = os.path.join(ROOT, rel)
    sys.path.insert(0, os.path.dirname(path))
    try:
        for m in ("dataset_paths",):          # each code dir has its own dataset_paths
            sys.modules.pop(m, None)
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod
    finally:
        sys.path.pop(0)


def _reset():
    from b200unet.keras import clear_session
    clear_session()


@pytest.mark.parametrize("pipeline", ["device", "host"])
def test_sr_trainer_and_offline_evaluator(tmp_path, capsys, pipeline):
    """pipeline: patches cropped / degraded / shuffled on the GPU (default) or by the host OpenCV stream."""
    host = ["--host_pipeline"] if pipeline == "host" else []
    _reset()
    tr = _load("Super_resolution/code/train_adaptive_unet.py", "train_adaptive_unet")
    args = tr.parse_args(["--scale", "0.5", "--depth_override", "2", "--patch_size", "32", "--batch_size", "4",
                          "--epochs", "2", "--patches_per_image", "2", "--synthetic", "6", "--precision", "bf16",
                          "--model_dir", str(tmp_path / "models"), "--log_dir", str(tmp_path / "logs"), "--run_name", "t"]
                         + host)
    hist = tr.train(args)
    out = capsys.readouterr().out
    assert re.search(r"Epoch 2/2\n\d+/\d+ - \d+s - \d+(ms|us)/step - loss: ", out), out[-2000:]
    assert len(hist.history["loss"]) == 2 and "psnr" in hist.history
    cfg = json.loads((tmp_path / "logs" / "t" / "config.json").read_text())
    assert cfg["depth"] == 2 and cfg["scale"] == 0.5 and cfg["patch_size"] == 32
    assert (tmp_path / "logs" / "t" / "model_summary.txt").exists()
    ckpt = tmp_path / "models" / "unet_adaptive_scale_new_loss0.50_depth2.keras"
    assert ckpt.exists()

    _reset()
    ev = _load("Super_resolution/code/evaluate_model.py", "evaluate_model")
    summary = ev.main(["--model-path", str(ckpt), "--scale", "0.5", "--depth-override", "2", "--patch-size", "32",
                       "--batch-size", "4", "--synthetic", "2", "--output-dir", str(tmp_path / "eval"), "--run-name", "e"]
                      + [h.replace("_", "-") for h in host])
    rep = tmp_path / "eval" / "e"
    m = json.loads((rep / "metrics.json").read_text())
    assert set(m) == {"mse_mean", "mse_std", "psnr_mean", "psnr_std", "ssim_mean", "ssim_std", "msssim_mean",
                      "msssim_std", "samples"}
    assert m["samples"] == summary.samples > 0 and m["psnr_mean"] > 5.0
    header = (rep / "per_image_metrics.csv").read_text().splitlines()[0]
    assert header == "index,filename,psnr_y,ssim_y,msssim_y,mse_y"
    assert json.loads((rep / "config.json").read_text())["eval_shave"] == 4


def test_sr_depth3_wrapper_pins_depth(tmp_path):
    _reset()
    _load("Super_resolution/code/train_adaptive_unet.py", "train_adaptive_unet")
    w = _load("Super_resolution/code/train_adaptive_unet_depth_3.py", "train_adaptive_unet_depth_3")
    assert hasattr(w, "main") or hasattr(w, "train") or hasattr(w, "parse_args")


def test_seg_adaptive_trainer(tmp_path, capsys):
    _reset()
    tr = _load("Segmenation/code/train_adaptive_unet.py", "seg_train_adaptive_unet")
    args = tr.parse_args(["--protocol", "A", "--epochs", "2", "--batch_size", "4", "--base_channels", "16", "--depth", "2",
                          "--image_size", "32", "--synthetic", "8", "--model_dir", str(tmp_path / "m"),
                          "--log_dir", str(tmp_path / "l"), "--run_name", "s", "--fit_verbose", "2"])
    hist, metrics = tr.train(args)
    assert len(hist.history["loss"]) == 2
    assert {"loss", "dice", "iou"} <= set(metrics)
    cfg = json.loads((tmp_path / "l" / "s" / "config.json").read_text())
    assert cfg["protocol"] == "A" and cfg["train_samples"] == 8 and cfg["threshold"] == 0.5
    assert "val_dice" in hist.history


@pytest.mark.parametrize("num_classes", [1, 3])
def test_seg_vanilla_trainer(tmp_path, num_classes):
    _reset()
    tr = _load("Segmenation/code/unet_vinillia.py", "seg_unet_vinillia")
    args = tr.parse_args(["--epochs", "2", "--batch_size", "4", "--base_channels", "16", "--depth", "2", "--image_size", "32",
                          "--synthetic", "8", "--augment", "--model_dir", str(tmp_path / "m"), "--run_name", "v",
                          "--num_classes", str(num_classes), "--fit_verbose", "0"])
    hist = tr.train(args)
    assert len(hist.history["loss"]) == 2 and hist.history["loss"][-1] == hist.history["loss"][-1]   # not NaN
    assert (tmp_path / "m" / "v_final.keras").exists()


def test_sr_vanilla_baseline_trainer(tmp_path, capsys):
    """Super_resolution/code/u-net-vinillia.py: BatchNorm baseline, whole-image stacks, RGB PSNR / SSIM / MS-SSIM evaluation."""
    _reset()
    mod = _load("Super_resolution/code/u-net-vinillia.py", "sr_unet_vanilla")
    args = mod.parse_args(["--synthetic", "10", "--hr_size", "192", "--batch_size", "2", "--epochs", "2", "--precision", "bf16",
                           "--model_dir", str(tmp_path / "m")])
    hist, res = mod.main(args)
    out = capsys.readouterr().out
    assert "Validation metrics:" in out and "Test metrics:" in out
    assert len(hist.history["loss"]) == 2 and "val_psnr" in hist.history
    assert (tmp_path / "m" / "unet_vanilla_best.keras").exists()
    for split in ("validation", "test"):
        m = res[split]
        assert set(m) == {"psnr", "ssim", "ms_ssim"} and all(np.isfinite(v[0]) for v in m.values())
        assert 0.0 < m["ssim"][0] <= 1.0 and 0.0 < m["ms_ssim"][0] <= 1.0
    with pytest.raises(NotImplementedError):
        mod.main(mod.parse_args(["--synthetic", "4", "--loss", "combined"]))
    _reset()
