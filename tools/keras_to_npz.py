"""Convert a Keras-3 ``.keras`` archive (zip: config.json + metadata.json + model.weights.h5) written by the reference
(/root/reference/Super_resolution/code/train_adaptive_unet.py:531,617) into the ``.keras`` flavour this repo reads
(zip with a ``model.weights.npz`` member: arrays w0000.. in layer / weight order, + config.json with "weight_names").

Needs h5py, which the build image lacks -- run it wherever the reference checkpoints live:
    python tools/keras_to_npz.py unet_adaptive_scale_new_loss0.50_depth3.keras out.keras

Weight order: Keras stores every layer's variables under ``layers/<layer>/vars/<i>`` in build order; conv kernels are
HWIO and Conv2DTranspose kernels [kh,kw,Cout,Cin] in both formats, so arrays are copied as they are.  The layer order
of ``build_super_resolution_unet`` (names conv2d, layer_normalization, ... residual_rgb) is the same in both
implementations (checked against the reference's model_summary dumps by tests/test_oracle_pins.py)."""
import io
import json
import sys
import zipfile

import numpy as np


def _natural(name):
    import re
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", name)]


def convert(src, dst):
    try:
        import h5py
    except ImportError as e:  # pragma: no cover
        raise SystemExit("keras_to_npz needs h5py (pip install h5py) -- not available in the build image") from e
    with zipfile.ZipFile(src) as z:
        cfg = json.loads(z.read("config.json"))
        blob = z.read("model.weights.h5")
    layer_names = [ly["config"]["name"] for ly in cfg["config"]["layers"]]
    arrays, names = [], []
    with h5py.File(io.BytesIO(blob), "r") as f:
        root = f["layers"] if "layers" in f else f["_layer_checkpoint_dependencies"]
        for ln in layer_names:
            if ln not in root or "vars" not in root[ln]:
                continue
            grp = root[ln]["vars"]
            for k in sorted(grp.keys(), key=_natural):
                arrays.append(np.asarray(grp[k], np.float32))
                names.append(f"{ln}/var{k}")
    buf = io.BytesIO()
    np.savez(buf, **{f"w{i:04d}": a for i, a in enumerate(arrays)})
    with zipfile.ZipFile(dst, "w") as z:
        z.writestr("config.json", json.dumps({"name": cfg["config"].get("name"), "weight_names": names, "source": str(src)}))
        z.writestr("model.weights.npz", buf.getvalue())
    print(f"{dst}: {len(arrays)} arrays, {sum(a.size for a in arrays):,} parameters")


if __name__ == "__main__":
    if len(sys.argv) != 3:
        raise SystemExit(__doc__)
    convert(sys.argv[1], sys.argv[2])
