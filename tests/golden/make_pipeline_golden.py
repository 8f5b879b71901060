"""Golden fixtures for the patch pipeline, produced by the REFERENCE'S OWN CODE.

``/root/reference/shared/pipeline.py`` is plain numpy + OpenCV except for the ``import tensorflow`` at its top
(used only by the tf.data wrappers).  OpenCV is importable in the build container, so this script loads the
unmodified reference module behind an empty ``tensorflow`` stub and records what ``random_patches``,
``grid_patches`` and ``degrade_image`` return on seeded inputs.  The vectors therefore pin this row against the
reference itself (cv2 4.13 here; the reference pins 4.9.0.80 -- same resize algorithm), unlike the TF ops.

Run in the build container (the GPU box has no /root/reference):
    python tests/golden/make_pipeline_golden.py        -> tests/golden/pipeline_ref.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/shared/pipeline.py"
SEED = 1234   # the reference's default seed (Super_resolution/code/train_adaptive_unet.py:742)


def load_reference():
    stub = types.ModuleType("tensorflow")
    stub.data = types.SimpleNamespace(Dataset=object, AUTOTUNE=-1)
    stub.TensorSpec = lambda **kw: None
    stub.float32 = "float32"
    sys.modules.setdefault("tensorflow", stub)
    spec = importlib.util.spec_from_file_location("ref_pipeline", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    rng = np.random.default_rng(SEED)
    d = {}
    # one synthetic "photo" (uint8 RGB, smooth + texture so that the cubic overshoots) and its float form
    yy, xx = np.mgrid[0:150, 0:201].astype(np.float32)
    base = np.stack([np.sin(xx / 9.0) * np.cos(yy / 7.0), np.sin((xx + yy) / 13.0), np.cos(xx / 5.0 - yy / 11.0)], -1)
    img_u8 = np.clip(127.5 + 90.0 * base + 40.0 * rng.standard_normal(base.shape), 0, 255).astype(np.uint8)
    img = img_u8.astype(np.float32) / 255.0          # = load_rgb_image_full after decode (pipeline.py:70-76)
    d["image_u8"] = img_u8
    # random_patches with the reference's generator discipline (pipeline.py:97-136)
    g = np.random.default_rng(SEED)
    hr = ref.random_patches(img, 64, count=3, rng=g)
    d["rand_hr_p64"] = hr
    for scale in (0.25, 0.5, 0.7):
        d[f"rand_lr_p64_s{scale}"] = np.stack([ref.degrade_image(p, scale, 64) for p in hr])
    # grid_patches (pipeline.py:139-175): a strided grid, and an image exactly one patch high
    gp = ref.grid_patches(img, 48, stride=60)
    d["grid_hr_p48_s60"] = gp
    d["grid_lr_p48_s60_s0.3"] = np.stack([ref.degrade_image(p, 0.3, 48) for p in gp])
    d["grid_hr_p40"] = ref.grid_patches(img[:40, :95], 40)
    # degrade_image on float patches that leave [0,1] (clipped before the shrink, not after the enlargement)
    for P, scale in ((128, 0.25), (64, 0.5), (64, 0.2), (96, 0.33), (37, 0.4)):
        x = (rng.random((1, P, P, 3), dtype=np.float32) * 1.2 - 0.1).astype(np.float32)
        d[f"deg_x_p{P}_s{scale}"] = x
        d[f"deg_y_p{P}_s{scale}"] = np.stack([ref.degrade_image(p, scale, P) for p in x])
    np.savez_compressed(os.path.join(OUT, "pipeline_ref.npz"), **d)
    print("wrote pipeline_ref.npz:", {k: v.shape for k, v in d.items()})


if __name__ == "__main__":
    main()
