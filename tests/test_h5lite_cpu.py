"""The in-tree HDF5 reader / writer behind `.keras` checkpoints (h5py is not installable in this image).

  * reader pinned on a file the REAL HDF5 library wrote: scipy ships a MATLAB v7.3 (= HDF5 behind a 512-byte user
    block) test file holding `testdouble = linspace(0, 2*pi, 9)`;
  * writer pinned on the reader (round trip, groups of 0 .. 300 children = empty / one-leaf / two-level B-trees) and
    on the structural rules of the HDF5 specification (`validate`), which the real file passes too;
  * the keras-3 group layout (`layers/<class_snake[_k]>/vars/<i>`) of Model.save without a GPU."""
import importlib.util
import os
import struct

import numpy as np
import pytest

import b200unet  # noqa: F401
from b200unet import h5lite


def _real_file():
    spec = importlib.util.find_spec("scipy")
    if spec is None:
        return None
    p = os.path.join(os.path.dirname(spec.origin), "io", "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    return p if os.path.exists(p) else None


def _equal(a, b):
    if isinstance(a, dict):
        return isinstance(b, dict) and set(a) == set(b) and all(_equal(a[k], b[k]) for k in a)
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.skipif(_real_file() is None, reason="scipy's HDF5 test file is not installed")
def test_reader_on_a_file_written_by_the_hdf5_library():
    data = open(_real_file(), "rb").read()
    r = h5lite.H5Reader(data)
    assert r.base == 512 and (r.leaf_k, r.internal_k) == (4, 16)
    tree = r.tree()
    assert list(tree) == ["testdouble"]
    a = tree["testdouble"]
    assert a.dtype == np.float64 and a.shape == (9, 1)
    assert np.allclose(a[:, 0], np.linspace(0, 2 * np.pi, 9), rtol=0, atol=1e-15)
    assert h5lite.validate(data) == 2          # root group + the dataset


def _sample_tree(n_layers, seed=0):
    rng = np.random.default_rng(seed)
    layers = {}
    for i in range(n_layers):
        layers[f"conv2d_{i}" if i else "conv2d"] = {"vars": {"0": rng.standard_normal((3, 3, 4, 8)).astype(np.float32),
                                                             "1": rng.standard_normal(8).astype(np.float32)}}
    return {"vars": {}, "layers": layers,
            "optimizer": {"vars": {"0": np.array(7, dtype=np.int64), "1": np.array(1e-4, dtype=np.float32)}},
            "activation": {"vars": {}}, "f64": np.arange(6, dtype=np.float64).reshape(2, 3),
            "empty": np.zeros((0, 3), np.float32), "half": np.arange(4, dtype=np.float16),
            "u8": np.arange(5, dtype=np.uint8), "i32": -np.arange(3, dtype=np.int32)}


@pytest.mark.parametrize("n_layers", [0, 1, 8, 9, 70, 257, 300])
def test_write_read_round_trip_and_structure(n_layers):
    tree = _sample_tree(n_layers)
    blob = h5lite.write_h5(tree)
    assert blob[:8] == h5lite.SIGNATURE and len(blob) % 8 == 0
    assert struct.unpack_from("<Q", blob, 40)[0] == len(blob)            # end-of-file address
    n_objects = h5lite.validate(blob)
    # root, vars, layers | per layer: group, vars, 2 arrays | optimizer: group, vars, 2 scalars | activation: group, vars | 5 arrays
    assert n_objects == 3 + 4 * n_layers + 4 + 2 + 5
    assert _equal(tree, h5lite.read_h5(blob))
    paths = dict(h5lite.H5Reader(blob).datasets())
    assert ("layers/conv2d/vars/0" in paths) == (n_layers > 0)


def test_names_are_utf8_sorted_and_dtype_messages_match_the_library():
    tree = {"b": np.zeros(2, np.float32), "a_long_name_over_eight_bytes": np.ones(1, np.float64), "Z": np.zeros(1, np.float32)}
    blob = h5lite.write_h5(tree)
    assert list(h5lite.read_h5(blob)) == ["Z", "a_long_name_over_eight_bytes", "b"]      # strcmp order
    # the float64 datatype message equals, byte for byte, the one the HDF5 library wrote into the real file
    want = bytes.fromhex("11203f0008000000000040003 40b0034ff030000".replace(" ", ""))
    assert h5lite._datatype_message(np.float64) == want


def test_reader_rejects_what_it_does_not_support():
    with pytest.raises(h5lite.H5Error, match="not an HDF5 file"):
        h5lite.read_h5(b"PK\x03\x04" + b"\0" * 600)
    blob = bytearray(h5lite.write_h5({"x": np.zeros(3, np.float32)}))
    blob[8] = 2                                                          # superblock version 2 (libver='latest')
    with pytest.raises(h5lite.H5Error, match="superblock version 2"):
        h5lite.read_h5(bytes(blob))
    with pytest.raises(h5lite.H5Error):
        h5lite.write_h5({"bad/name": np.zeros(1, np.float32)})
    with pytest.raises(h5lite.H5Error):
        h5lite.write_h5({"c": np.zeros(1, np.complex64)})


def test_validator_catches_corruption():
    blob = bytearray(h5lite.write_h5(_sample_tree(12)))
    i = blob.index(b"SNOD")
    # swap the first two symbol-table entries of a leaf: names no longer sorted
    a, b = bytes(blob[i + 8:i + 48]), bytes(blob[i + 48:i + 88])
    blob[i + 8:i + 48], blob[i + 48:i + 88] = b, a
    with pytest.raises((AssertionError, h5lite.H5Error)):
        h5lite.validate(bytes(blob))


def test_keras3_group_names_of_the_sr_model():
    """keras names the groups of model.weights.h5 after the layer CLASS, numbered per class in model.layers order."""
    from b200unet import builders as B
    from b200unet.keras import clear_session
    clear_session()
    model, _ = B.build_super_resolution_unet(0.5, depth_override=1, input_size=16)
    names = model._h5_layer_names()
    assert names[0] == "input_layer" and names[1] == "conv2d" and "conv2d_1" in names
    assert names.count("resize_by_scale") == 1 and names.count("resize_to_match") == 1 and names[-1] == "clipped_residual_add"
    assert len(set(names)) == len(names) == len(model.layers)
    convs = [n for n, ly in zip(names, model.layers) if type(ly).__name__ == "Conv2D"]
    assert convs == ["conv2d"] + [f"conv2d_{k}" for k in range(1, len(convs))]
