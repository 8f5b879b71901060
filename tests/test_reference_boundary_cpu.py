"""Drop-in boundary against the REFERENCE'S OWN entry points (build container only: needs /root/reference).

The reference's scripts are Python on TensorFlow/Keras, which cannot be installed here -- but their command lines and
their pure-Python helpers do not need TF to run.  An import hook hands out inert stand-ins for every ``tensorflow`` /
``keras`` / ``optuna`` module, the unmodified reference files are imported next to the mirrors of this repo, and the
test compares (a) every command-line flag (option strings, type, choices, default, store_true-ness) of the five entry
points, (b) the helper functions both sides share (depth rules, index splits, eval shave, natural sort, mask-file
matching) on grids of inputs, (c) the training protocols' hyper-parameters."""
import argparse
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types
from pathlib import Path
from unittest.mock import MagicMock

import numpy as np
import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (build container only)")


SERIALIZABLE_KEYS = {}


class _LayerBase:
    """Stand-in for keras.layers.Layer: enough for the reference's custom layers to be REAL classes whose code runs."""

    def __init__(self, name=None, **kwargs):
        self.name = name

    def get_config(self):
        return {"name": self.name}

    def __call__(self, inputs):
        return self.call(inputs)


def _register_keras_serializable(package="Custom", name=None):
    def deco(cls):
        SERIALIZABLE_KEYS[f"{package}>{name or cls.__name__}"] = cls
        return cls
    return deco


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name in ("keras", "layers", "saving"):        # ``from tensorflow.keras import layers as L``: a stand-in MODULE
            import importlib
            m = importlib.import_module(f"{self.__name__}.{name}")
        elif name == "Layer":
            m = _LayerBase
        elif name == "register_keras_serializable":
            m = _register_keras_serializable
        else:
            m = MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ROOTS = ("tensorflow", "keras", "optuna")

    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def _load(path, name, extra_paths):
    for p in extra_paths:
        sys.path.insert(0, p)
    sys.modules.pop("dataset_paths", None)          # every code directory has its own
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod
    finally:
        for p in extra_paths:
            sys.path.remove(p)


ENTRY_POINTS = {
    # relative path: (code dir, minimal argv that satisfies the reference's required flags)
    "Super_resolution/code/train_adaptive_unet.py": ("Super_resolution/code", ["--scale", "0.5"]),
    "Super_resolution/code/evaluate_model.py": ("Super_resolution/code", ["--scale", "0.5", "--model-path", "m.keras"]),
    "Super_resolution/code/u-net-vinillia.py": ("Super_resolution/code", ["--high_res_dir", "a", "--low_res_dir", "b"]),
    "Segmenation/code/train_adaptive_unet.py": ("Segmenation/code", []),
    "Segmenation/code/unet_vinillia.py": ("Segmenation/code", []),
}


@pytest.fixture(scope="module")
def sides():
    """{rel: (reference module, mirror module)} with the TF stand-ins installed only while this module's tests run."""
    finder = _StubFinder()
    before = set(sys.modules)
    sys.meta_path.insert(0, finder)
    try:
        out = {}
        for rel, (code_dir, _) in ENTRY_POINTS.items():
            tag = os.path.basename(rel)[:-3].replace("-", "_")
            ref = _load(os.path.join(REF, rel), f"_ref_{code_dir[:3]}_{tag}", [REF, os.path.join(REF, code_dir)])
            mine = _load(os.path.join(ROOT, rel), f"_mine_{code_dir[:3]}_{tag}", [os.path.join(ROOT, code_dir)])
            out[rel] = (ref, mine)
        out["custom_layers"] = (_load(os.path.join(REF, "shared/custom_layers.py"), "_ref_custom_layers", [REF]), None)
        yield out
    finally:
        sys.meta_path.remove(finder)
        for name in set(sys.modules) - before:
            if name.split(".")[0] in _StubFinder.ROOTS or name.startswith(("_ref_", "_mine_")) or name == "dataset_paths":
                sys.modules.pop(name, None)


def _parser_of(mod, argv, pass_argv):
    """The ArgumentParser a module's parse_args() builds (captured at its parse_args call)."""
    cap = {}
    orig = argparse.ArgumentParser.parse_args

    def spy(self, args=None, namespace=None):
        cap["parser"] = self
        return orig(self, argv, namespace)

    argparse.ArgumentParser.parse_args = spy
    try:
        mod.parse_args(argv) if pass_argv else mod.parse_args()
    finally:
        argparse.ArgumentParser.parse_args = orig
    return cap["parser"]


def _flags(parser):
    out = {}
    for a in parser._actions:
        if a.option_strings and a.dest != "help":
            out[a.dest] = {"options": tuple(a.option_strings), "action": type(a).__name__, "default": a.default,
                           "type": getattr(a.type, "__name__", a.type), "choices": tuple(a.choices) if a.choices else None,
                           "required": a.required}
    return out


# defaults that are site paths of the authors' cluster / of this checkout, and flags this repo makes optional because
# --synthetic / --random-init can stand in for them
SITE_PATHS = {"model_dir", "log_dir", "hr_dir", "output_dir", "high_res_dir", "low_res_dir", "train_image_dir",
              "train_mask_dir", "val_image_dir", "val_mask_dir", "image_dir", "mask_dir"}
RELAXED_REQUIRED = {"model_path", "high_res_dir", "low_res_dir"}


@pytest.mark.parametrize("rel", list(ENTRY_POINTS))
def test_command_line_flags_match_the_reference(sides, rel):
    ref_mod, mine_mod = sides[rel]
    argv = ENTRY_POINTS[rel][1]
    ref, mine = _flags(_parser_of(ref_mod, argv, False)), _flags(_parser_of(mine_mod, argv, True))
    assert ref, rel
    for dest, spec in ref.items():
        assert dest in mine, f"{rel}: flag {spec['options']} of the reference is missing"
        got = mine[dest]
        for key in ("options", "action", "type", "choices"):
            assert got[key] == spec[key], (rel, dest, key, spec[key], got[key])
        if dest not in SITE_PATHS:
            assert got["default"] == spec["default"], (rel, dest, spec["default"], got["default"])
        assert got["required"] == spec["required"] or (dest in RELAXED_REQUIRED and not got["required"]), (rel, dest)
    extra = set(mine) - set(ref)        # additions must be optional: a reference command line runs unchanged
    assert all(not mine[d]["required"] for d in extra), extra


def test_depth_and_shape_rules_match(sides):
    ref = sides["custom_layers"][0]
    from b200unet.shared import custom_layers as CL
    scales = [0.1, 0.2, 0.25, 0.3, 0.33, 0.4, 0.45, 0.5, 0.6, 0.7, 0.75, 0.8, 0.9]
    for s in scales:
        assert CL.infer_depth_from_scale(s) == ref.infer_depth_from_scale(s)
        for res in (64, 128, 256, 512):
            for md in (3, 7):
                assert CL.custom_depth_from_scale(s, max_depth=md, base_resolution=res) == \
                    ref.custom_depth_from_scale(s, max_depth=md, base_resolution=res), (s, res, md)
            for d in (1, 2, 3, 5):
                assert CL.estimate_bottleneck_size(res, s, d) == ref.estimate_bottleneck_size(res, s, d)
        assert CL.depth_and_sizes(s) == ref.depth_and_sizes(s)
    for bad in (0.0, 1.0, -0.5, 1.5):
        with pytest.raises(ValueError):
            ref.custom_depth_from_scale(bad)
        with pytest.raises(ValueError):
            CL.custom_depth_from_scale(bad)


def test_index_split_eval_shave_and_sort_match(sides):
    from b200unet.shared import pipeline as PL
    ref_train = sides["Super_resolution/code/train_adaptive_unet.py"][0]
    ref_van = sides["Super_resolution/code/u-net-vinillia.py"][0]
    ref_eval, mine_eval = sides["Super_resolution/code/evaluate_model.py"]
    for n in (1, 2, 3, 5, 10, 37, 800):
        for tr, va, te in ((0.8, 0.1, 0.1), (0.9, 0.1, 0.0), (0.5, 0.25, 0.25), (0.98, 0.01, 0.01)):
            for seed in (0, 1234):
                try:
                    want = ref_train.split_indices(n, tr, va, te, seed)
                except ValueError:
                    with pytest.raises(ValueError):
                        PL.split_indices(n, tr, va, te, seed)
                    continue
                got = PL.split_indices(n, tr, va, te, seed)
                assert all(np.array_equal(a, b) for a, b in zip(want, got)), (n, tr, va, te, seed)
                v = ref_van.split_indices(n_samples=n, train=tr, val=va, test=te, seed=seed)
                assert all(np.array_equal(a, b) for a, b in zip(v, got))
    for s in (0.2, 0.25, 0.3, 0.5, 0.7, 0.9):
        for override in (None, 0, 3, -2):
            assert mine_eval.infer_eval_shave(s, override) == ref_eval.infer_eval_shave(s, override)
    for names in (["10.png", "9.png", "0001.png", "1.png", "100.png"],
                  ["img_12a.png", "IMG_3.png", "img_2b10.png", "img_2b9.png", "a", "B", "img_", "IMG_03.png"]):
        assert PL.sorted_alphanumeric(names) == ref_train.sorted_alphanumeric(names) == ref_van.sorted_alphanumeric(names)
    mixed = ["10.png", "img_1.png"]       # a digit-led and a letter-led name compare int with str: both sides raise
    for fn in (PL.sorted_alphanumeric, ref_train.sorted_alphanumeric):
        with pytest.raises(TypeError):
            fn(mixed)


def test_segmentation_protocols_and_file_matching(sides, tmp_path):
    ref, mine = sides["Segmenation/code/train_adaptive_unet.py"]
    mp = mine._protocols()
    assert set(mp) == set(ref.PROTOCOLS)
    for key, rp in ref.PROTOCOLS.items():
        for field in ("key", "initial_lr", "epochs", "batch_size", "cosine_schedule", "early_stopping_patience"):
            assert getattr(mp[key], field) == getattr(rp, field), (key, field)
    from b200unet.shared import seg_data as SD
    ref_van = sides["Segmenation/code/unet_vinillia.py"][0]
    for name in ("ISIC_0000001.jpg", "ISIC_0000001_segmentation.png", "isic_12_Segmentation.PNG", "plain.jpeg"):
        assert SD.canonical_key(Path(name)) == ref.normalise_isic_key(Path(name)) == ref_van._canonical_key(Path(name)), name
    for name in ("x_mask.png", "city_000_leftImg8bit.png", "city_000_gtFine_labelIds.png", "a_gtCoarse_color.png"):
        assert SD.canonical_key(Path(name)) == ref_van._canonical_key(Path(name)), name     # the baseline's longer token list
    img, msk = tmp_path / "img", tmp_path / "msk"
    img.mkdir(); msk.mkdir()
    for stem in ("ISIC_0000010", "ISIC_0000002", "ISIC_0000033", "ISIC_0000002_superpixels"):
        (img / f"{stem}.jpg").write_bytes(b"x")
    for stem in ("ISIC_0000002", "ISIC_0000010", "ISIC_0000099", "ISIC_0000033"):
        (msk / f"{stem}_segmentation.png").write_bytes(b"x")
    want = ref.collect_isic_pairs(img, msk)
    got = SD.collect_pairs(img, msk)
    assert [(Path(a).name, Path(b).name) for a, b in got] == [(Path(a).name, Path(b).name) for a, b in want]
    assert len(got) == 3                                        # superpixel images skipped, the spare mask ignored
    (img / "ISIC_0000500.jpg").write_bytes(b"x")                # an image without a mask: both sides refuse
    for fn in (ref.collect_isic_pairs, SD.collect_pairs):
        with pytest.raises(ValueError):
            fn(img, msk)
    for fn in (ref.collect_isic_pairs, SD.collect_pairs):
        with pytest.raises(FileNotFoundError):
            fn(tmp_path / "nope", msk)


# ------------------------------------------------------------------------------------------------------------------
# Model builders: run the REFERENCE'S builder code and this repo's builder code against the same recording stand-ins for
# Input / layers / Model and compare what they construct -- every layer (kind, constructor arguments normalised through
# the Keras signatures), every call (which tensors go in, in which order: the concat order [up, skip] matters), the model
# name.  Equal event lists = the same graph.
# ------------------------------------------------------------------------------------------------------------------
class _Tok:
    def __init__(self, i):
        self.id = i


class _GraphRecorder:
    ALIASES = {"MaxPool2D": "MaxPooling2D", "ClipAdd": "ClippedResidualAdd"}

    def __init__(self):
        self.events, self.n_tok, self.n_layer = [], 0, 0

    def _new_tok(self):
        self.n_tok += 1
        return _Tok(self.n_tok)

    def _canon(self, kind, args, kwargs):
        import inspect
        from b200unet.keras import layers as ML
        from b200unet.shared import custom_layers as CL
        cls = getattr(ML, kind, None) or getattr(CL, kind)
        sig = inspect.signature(cls.__init__)
        known = {k: v for k, v in kwargs.items() if k in sig.parameters}
        bound = sig.bind(None, *args, **known)
        bound.apply_defaults()
        d = {k: v for k, v in bound.arguments.items() if k not in ("self", "kwargs")}
        d.update({k: v for k, v in kwargs.items() if k not in sig.parameters})
        for k in ("kernel_size", "pool_size", "size", "strides"):
            if k in d and d[k] is not None and not isinstance(d[k], tuple):
                d[k] = (d[k], d[k])
        return tuple(sorted((k, repr(v)) for k, v in d.items()))

    def layer(self, kind):
        kind = self.ALIASES.get(kind, kind)

        def ctor(*args, **kwargs):
            lid = self.n_layer
            self.n_layer += 1
            self.events.append(("layer", lid, kind, self._canon(kind, args, kwargs)))

            def call(inputs):
                ins = tuple(t.id for t in inputs) if isinstance(inputs, (list, tuple)) else (inputs.id,)
                out = self._new_tok()
                self.events.append(("call", lid, ins, out.id))
                return out
            return call
        return ctor

    def input(self, shape=None, name=None, **kw):
        t = self._new_tok()
        self.events.append(("input", tuple(shape), name, t.id))
        return t

    def model(self, *args, **kwargs):
        inputs = kwargs.get("inputs", args[0] if args else None)
        outputs = kwargs.get("outputs", args[1] if len(args) > 1 else None)
        self.events.append(("model", inputs.id, outputs.id, kwargs.get("name")))
        return ("model", kwargs.get("name"))

    def patch(self, monkeypatch, mod):
        rec = self

        class _Layers:
            def __getattr__(_, kind):
                return rec.layer(kind)
        monkeypatch.setattr(mod, "L", _Layers(), raising=False)
        monkeypatch.setattr(mod, "Input", self.input, raising=False)
        monkeypatch.setattr(mod, "Model", self.model, raising=False)
        for kind in ("ResizeByScale", "ResizeToMatch", "ClippedResidualAdd"):
            if hasattr(mod, kind):
                monkeypatch.setattr(mod, kind, self.layer(kind))


def _record(monkeypatch, mod, fn_name, *args, **kwargs):
    rec = _GraphRecorder()
    with monkeypatch.context() as mp:
        rec.patch(mp, mod)
        result = getattr(mod, fn_name)(*args, **kwargs)
    return rec.events, result


def test_builders_construct_the_reference_graphs(sides, monkeypatch):
    from b200unet import builders as B
    ref_sr = sides["Super_resolution/code/train_adaptive_unet.py"][0]
    ref_van = sides["Super_resolution/code/u-net-vinillia.py"][0]
    ref_seg = sides["Segmenation/code/train_adaptive_unet.py"][0]
    ref_base = sides["Segmenation/code/unet_vinillia.py"][0]
    n = 0
    # adaptive-depth SR U-Net: explicit depths and the depth rule (train_adaptive_unet.py:217-287)
    for kw in (dict(scale=0.5, depth_override=3, input_size=64), dict(scale=0.25, depth_override=4, input_size=128),
               dict(scale=0.25, depth_override=5, input_size=128), dict(scale=0.7, input_size=256),
               dict(scale=0.2, input_size=256, max_depth=3), dict(scale=0.5, base_channels=32, residual_head_channels=16,
                                                                 depth_override=1, input_size=32)):
        (want, (_, winfo)), (got, (_, ginfo)) = (_record(monkeypatch, ref_sr, "build_super_resolution_unet", **kw),
                                                 _record(monkeypatch, B, "build_super_resolution_unet", **kw))
        assert got == want, kw
        assert ginfo == winfo, kw
        n += len(want)
    # fixed-depth BatchNorm SR baseline (u-net-vinillia.py:128-167)
    want, _ = _record(monkeypatch, ref_van, "build_super_resolution_unet", (256, 256, 3))
    got, _ = _record(monkeypatch, B, "build_vanilla_super_resolution_unet", (256, 256, 3))
    assert got == want
    n += len(want)
    # segmentation: adaptive BatchNorm U-Net (:335-362) and the LayerNorm / Conv2DTranspose baseline (unet_vinillia.py:72-91)
    for size, base, depth in ((256, 64, 4), (128, 32, 2), (256, 48, 5)):
        want, _ = _record(monkeypatch, ref_seg, "build_adaptive_depth_unet", size, base, depth)
        got, _ = _record(monkeypatch, B, "build_adaptive_depth_unet", size, base, depth)
        assert got == want, (size, base, depth)
        n += len(want)
    for kw in (dict(input_size=256), dict(input_size=256, num_classes=21, base_channels=32, depth=4),
               dict(input_size=64, num_classes=1, base_channels=16, depth=2)):
        want, _ = _record(monkeypatch, ref_base, "build_unet", **kw)
        got, _ = _record(monkeypatch, B, "build_unet", **kw)
        assert got == want, kw
        n += len(want)
    assert n > 1000          # layers + calls compared


# ------------------------------------------------------------------------------------------------------------------
# The reference's loss / metric / custom-layer CODE executed on numpy: a stand-in ``tf`` that maps the elementary ops
# these functions use (cast, reduce_mean/sum, square, sqrt, abs, clip_by_value, ceil, maximum, shape ...) onto numpy
# float32.  What is pinned is the COMPOSITION the reference wrote (which epsilon, where the clip sits, per-sample versus
# whole-batch sums, smoothing constants, the float32 size rule) -- the elementary ops are unambiguous.  The two library
# calls that are not elementary are stated here: tf.image.psnr = 10 log10(max^2 / per-image MSE); keras
# BinaryCrossentropy() = mean over all elements of the cross-entropy of probabilities clipped to [1e-7, 1 - 1e-7];
# tf.image.resize is only RECORDED (target size, method, antialias), its arithmetic stays with oracle/resize_np.py.
# ------------------------------------------------------------------------------------------------------------------
class _NumpyTF:
    float32, int32, Tensor = np.float32, np.int32, np.ndarray

    def __init__(self):
        self.resize_calls = []
        tf = self

        class _Math:
            ceil = staticmethod(np.ceil)

        class _Image:
            @staticmethod
            def psnr(a, b, max_val):
                mse = np.mean(np.square(np.asarray(a, np.float32) - np.asarray(b, np.float32)), axis=(-3, -2, -1))
                return (10.0 * np.log10(np.float32(max_val) ** 2 / mse)).astype(np.float32)

            @staticmethod
            def resize(x, size, method=None, antialias=False):
                tf.resize_calls.append((tuple(int(v) for v in size), method, antialias))
                return np.zeros((x.shape[0], int(size[0]), int(size[1]), x.shape[3]), np.float32)

        class _BCE:
            def __call__(self, y_true, y_pred):
                p = np.clip(np.asarray(y_pred, np.float32), 1e-7, 1.0 - 1e-7)
                t = np.asarray(y_true, np.float32)
                return np.mean(-(t * np.log(p) + (1.0 - t) * np.log(1.0 - p)), dtype=np.float32)

        self.math, self.image = _Math, _Image
        self.keras = types.SimpleNamespace(losses=types.SimpleNamespace(BinaryCrossentropy=_BCE))

    cast = staticmethod(lambda x, dtype: np.asarray(x).astype(dtype))
    constant = staticmethod(lambda v, dtype=None: np.asarray(v, dtype=dtype))
    reshape = staticmethod(lambda x, shape: np.reshape(x, shape))
    square, sqrt, abs = staticmethod(np.square), staticmethod(np.sqrt), staticmethod(np.abs)
    maximum = staticmethod(np.maximum)
    clip_by_value = staticmethod(lambda x, lo, hi: np.clip(x, lo, hi))
    shape = staticmethod(lambda x: np.asarray(x.shape, dtype=np.int32))

    @staticmethod
    def reduce_mean(x, axis=None, keepdims=False):
        return np.mean(x, axis=tuple(axis) if isinstance(axis, list) else axis, keepdims=keepdims)

    @staticmethod
    def reduce_sum(x, axis=None, keepdims=False):
        return np.sum(x, axis=tuple(axis) if isinstance(axis, list) else axis, keepdims=keepdims)


def test_reference_loss_metric_and_layer_code_on_numpy(sides, monkeypatch):
    import torch
    from oracle import keras_ops as K, metrics_ref as MR
    from b200unet.shared import custom_layers as CL
    ref_sr = sides["Super_resolution/code/train_adaptive_unet.py"][0]
    ref_seg = sides["Segmenation/code/train_adaptive_unet.py"][0]
    ref_base = sides["Segmenation/code/unet_vinillia.py"][0]
    ref_cl = sides["custom_layers"][0]
    tf = _NumpyTF()
    for mod in (ref_sr, ref_seg, ref_base, ref_cl):
        monkeypatch.setattr(mod, "tf", tf)
    rng = np.random.default_rng(0)
    t = lambda a: torch.from_numpy(np.asarray(a, np.float32))

    # SR losses and the PSNR metric (train_adaptive_unet.py:294-334) against the oracle restatements
    y = rng.random((3, 16, 16, 3), dtype=np.float32)
    p = (y + 0.1 * rng.standard_normal(y.shape)).astype(np.float32)          # leaves [0, 1]
    charb, (psnr,) = ref_sr.build_losses_and_metrics("charbonnier")
    l1, _ = ref_sr.build_losses_and_metrics("l1")
    assert abs(float(charb(y, p)) - K.charbonnier_loss(t(y), t(p)).item()) < 1e-6
    assert abs(float(l1(y, p)) - K.l1_loss(t(y), t(p)).item()) < 1e-6
    assert abs(float(psnr(y, p)) - K.psnr_metric(t(y), t(p)).item()) < 1e-4
    with pytest.raises(ValueError):
        ref_sr.build_losses_and_metrics("huber")
    # BT.601 luma of the eval loops (:144-157)
    assert np.abs(ref_sr.rgb_to_luma_bt601(p) - MR.rgb_to_luma_bt601(t(p)).numpy()).max() < 1e-6

    # segmentation: per-sample dice / iou, the two hybrid losses (seg :258-304), the baseline's whole-batch dice (:94-99)
    m = (rng.random((4, 12, 12, 1)) > 0.6).astype(np.float32)
    q = rng.random((4, 12, 12, 1), dtype=np.float32)
    q[0, :2] = 0.0; q[1, :2] = 1.0                                              # probabilities on the clip bounds
    assert abs(float(ref_seg.dice_coefficient(m, q)) - K.dice_coefficient(t(m), t(q)).item()) < 1e-6
    assert abs(float(ref_seg.iou_score(m, q)) - K.iou_score(t(m), t(q)).item()) < 1e-6
    for maker, (a, b) in ((ref_seg.make_hybrid_ce_dice_loss, (0.4, 0.6)), (ref_seg.make_bce_dice_loss, (0.5, 1.0))):
        assert abs(float(maker(a, b)(m, q)) - K.bce_dice_loss(t(m), t(q), a, b).item()) < 2e-6
    assert abs(float(ref_base.dice_coefficient(m, q)) - K.dice_coefficient_global(t(m), t(q)).item()) < 1e-6

    # custom layers: the reference classes are real here (stand-in base class), their call() code runs on numpy
    assert set(SERIALIZABLE_KEYS) >= {"resize>ResizeByScale", "resize>ResizeToMatch", "utils>ClippedResidualAdd"}
    assert set(CL.get_custom_objects()) >= {"resize>ResizeByScale", "resize>ResizeToMatch", "utils>ClippedResidualAdd"}
    assert ref_cl.ClipAdd is ref_cl.ClippedResidualAdd and CL.ClipAdd is CL.ClippedResidualAdd
    inp = rng.random((2, 5, 5, 3), dtype=np.float32)
    res = (0.8 * rng.standard_normal(inp.shape)).astype(np.float32)
    out = ref_cl.ClippedResidualAdd(name="enhanced_rgb").call([inp, res])
    assert np.abs(out - K.clipped_residual_add(t(inp), t(res)).numpy()).max() < 1e-7 and out.dtype == np.float32
    for scale in (0.2, 0.25, 0.3, 0.5, 0.7, 0.9):
        for h, w in ((256, 256), (180, 126), (89, 63), (45, 45), (7, 3), (1, 1)):
            tf.resize_calls.clear()
            ref_cl.ResizeByScale(scale, name="enc_down").call(np.zeros((1, h, w, 4), np.float32))
            (size, method, aa), = tf.resize_calls
            mine = CL.ResizeByScale(scale).compute_output_shape([(1, h, w, 4)])
            assert size == (mine[1], mine[2]) == (CL.resized_extent(h, scale), CL.resized_extent(w, scale)), (scale, h, w)
            assert method == "bilinear" and aa is True
    tf.resize_calls.clear()
    ref_cl.ResizeToMatch(name="dec_up").call((np.zeros((1, 8, 9, 4), np.float32), np.zeros((1, 31, 17, 6), np.float32)))
    assert tf.resize_calls == [((31, 17), "bilinear", True)]
    assert CL.ResizeToMatch().compute_output_shape([(1, 8, 9, 4), (1, 31, 17, 6)]) == (1, 31, 17, 4)
    rc, mc = ref_cl.ResizeByScale(0.3, name="d").get_config(), CL.ResizeByScale(0.3, name="d").get_config()
    assert {k: mc[k] for k in rc} == rc
    rc, mc = ref_cl.ResizeToMatch(name="u").get_config(), CL.ResizeToMatch(name="u").get_config()
    assert {k: mc[k] for k in rc} == rc


def test_epoch_lines_parse_with_the_reference_log_exporter(tmp_path):
    """Model.fit(verbose=2) prints what Super_resolution/code/export_log_metrics.py turns into CSV rows: run the reference's
    parser on lines produced by this repo's formatter."""
    from b200unet.keras.model import format_epoch_line
    spec = importlib.util.spec_from_file_location("_ref_export_log_metrics",
                                                  os.path.join(REF, "Super_resolution/code/export_log_metrics.py"))
    exp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(exp)
    logs = {"loss": 0.012345, "psnr": 31.2468, "val_loss": 0.00045678, "val_psnr": 30.5, "learning_rate": 1e-4}
    row = exp.parse_metrics_line(format_epoch_line(250, 103.4, logs))
    assert row["steps_completed"] == row["steps_total"] == 250 and row["duration_s"] == 103 and row["ms_per_step"] == 414
    for k, v in logs.items():
        assert abs(row[k] - v) <= 5e-5 * max(1.0, abs(v)) + 5e-8, (k, row[k], v)
    log = tmp_path / "run-simple-1.log"
    text = ["some banner", "Epoch 1/2", format_epoch_line(250, 103.4, logs), "Epoch 1: val_loss improved from inf to 0.00046",
            "Epoch 2/2", format_epoch_line(250, 99.0, {**logs, "loss": 0.011}), "Training complete."]
    log.write_text("\n".join(text) + "\n")
    rows = exp.extract_epoch_rows(log)
    assert [int(r["epoch"]) for r in rows] == [1, 2] and abs(rows[1]["loss"] - 0.011) < 1e-6


def test_offline_evaluator_artefacts_are_byte_identical(sides, tmp_path):
    """evaluate_model.py's result files (config.json, metrics.json, per_image_metrics.csv :166-190): the reference's writer and
    this repo's, on the same results, produce the same bytes."""
    import dataclasses
    ref, mine = sides["Super_resolution/code/evaluate_model.py"]
    assert [f.name for f in dataclasses.fields(ref.EvalResults)] == [f.name for f in dataclasses.fields(mine.EvalResults)]
    vals = (1.25e-3, 4e-4, 31.25, 1.5, 0.9123, 0.02, float("nan"), float("nan"), 3)
    per = lambda: [{"index": i, "psnr_y": 30.0 + i, "ssim_y": 0.9 + 0.01 * i, "msssim_y": float("nan"), "mse_y": 1e-3 / (i + 1)}
                   for i in range(3)]
    names = ["0801.png#patch0000", "0801.png#patch0001", "0802.png#patch0000"]
    config = {"model_path": "m.keras", "scale": 0.5, "patch_size": 256, "eval_shave": 4, "samples": 3}
    out = {}
    for tag, mod in (("ref", ref), ("mine", mine)):
        rows = per()
        mod.attach_filenames(rows, names)
        mod.write_outputs(tmp_path / tag, mod.EvalResults(*vals), rows, config, True)
        out[tag] = {f.name: f.read_bytes() for f in sorted((tmp_path / tag).iterdir())}
        with pytest.raises(ValueError):
            mod.attach_filenames(per(), names[:2])
    assert set(out["ref"]) == {"config.json", "metrics.json", "per_image_metrics.csv"}
    assert out["mine"] == out["ref"]
    ref.write_outputs(tmp_path / "r2", ref.EvalResults(*vals), per(), config, False)
    mine.write_outputs(tmp_path / "m2", mine.EvalResults(*vals), per(), config, False)
    assert sorted(p.name for p in (tmp_path / "m2").iterdir()) == sorted(p.name for p in (tmp_path / "r2").iterdir())


def test_sr_training_protocol_matches_the_reference_trainer(sides, tmp_path, monkeypatch, capsys):
    """Run the reference's ``train(args)`` (its Keras calls land on inert stand-ins that record them) and this repo's
    ``train(args)`` (``Model.fit`` / ``Model.__call__`` / the metric kernels replaced by recorders: no GPU here) on the same
    image directory and the same command line.  Compared: the run's ``config.json`` (splits, patch counts, steps per epoch
    ...), the ``compile`` and ``fit`` arguments, the callbacks and their settings, the checkpoint name, the info line."""
    cv2 = pytest.importorskip("cv2")
    import json
    import torch
    ref, mine = sides["Super_resolution/code/train_adaptive_unet.py"]
    from b200unet import metrics as MT
    from b200unet.keras import clear_session, model as MM
    data = tmp_path / "hr"
    data.mkdir()
    rng = np.random.default_rng(0)
    for i in range(9):
        cv2.imwrite(str(data / f"{i:04d}.png"), rng.integers(0, 256, (80, 90, 3), dtype=np.uint8))

    def argv(tag):
        return ["--scale", "0.5", "--high_res_dir", str(data), "--patch_size", "32", "--batch_size", "4", "--epochs", "3",
                "--patches_per_image", "3", "--learning_rate", "2e-4", "--patience", "7", "--eval_stride", "24",
                "--model_dir", str(tmp_path / tag / "m"), "--log_dir", str(tmp_path / tag / "l"), "--run_name", "t"]

    # ---- the reference trainer on stand-ins
    for name in ("Model", "EarlyStopping", "ModelCheckpoint", "BackupAndRestore"):
        monkeypatch.setattr(ref, name, MagicMock(name=name))
    monkeypatch.setattr(sys, "argv", ["train_adaptive_unet.py"] + argv("ref"))
    ref.train(ref.parse_args())
    ref_out = capsys.readouterr().out
    rmodel = ref.Model.return_value
    r_fit, r_compile = rmodel.fit.call_args, rmodel.compile.call_args
    r_cfg = json.loads((tmp_path / "ref" / "l" / "t" / "config.json").read_text())

    # ---- this repo's trainer with the GPU work replaced by recorders
    rec = {}

    def fake_fit(self, x=None, **kw):
        rec["fit"], rec["model"] = kw, self
        h = MM.History()
        h.epoch, h.history = [0, 1, 2], {"loss": [1.0, 0.5, 0.4]}
        return h

    clear_session()
    monkeypatch.setattr(MM.Model, "fit", fake_fit)
    monkeypatch.setattr(MM.Model, "__call__", lambda self, x, training=False: torch.zeros(tuple(x.shape)))
    monkeypatch.setattr(MT, "eval_luma_metrics",
                        lambda pred, hr, shave=0: {k: np.ones(len(hr), np.float32) for k in ("psnr", "ssim", "msssim", "mse")})
    mine.train(mine.parse_args(argv("mine") + ["--host_pipeline"]))
    mine_out = capsys.readouterr().out
    m_cfg = json.loads((tmp_path / "mine" / "l" / "t" / "config.json").read_text())
    clear_session()

    # config.json: same keys in the same order, same values except the output locations and the time stamp
    assert list(m_cfg) == list(r_cfg)
    for k in r_cfg:
        if k not in ("model_dir", "log_dir", "created_at"):
            assert m_cfg[k] == r_cfg[k], (k, r_cfg[k], m_cfg[k])
    assert (tmp_path / "mine" / "l" / "t" / "model_summary.txt").exists() and (tmp_path / "ref" / "l" / "t" / "model_summary.txt").exists()
    # fit(...)
    for k in ("epochs", "initial_epoch", "steps_per_epoch", "validation_steps", "validation_freq", "verbose"):
        assert rec["fit"][k] == r_fit.kwargs[k], (k, r_fit.kwargs[k], rec["fit"][k])
    assert (rec["fit"]["validation_data"] is None) == (r_fit.kwargs["validation_data"] is None)
    # callbacks, in order, with their settings
    r_cbs = r_fit.kwargs["callbacks"]
    m_cbs = rec["fit"]["callbacks"]
    assert [type(c).__name__ for c in m_cbs] == ["EarlyStopping", "ModelCheckpoint", "BackupAndRestore", "TensorBoard"]
    assert r_cbs[0] is ref.EarlyStopping.return_value and r_cbs[1] is ref.ModelCheckpoint.return_value
    assert r_cbs[2] is ref.BackupAndRestore.return_value
    es, ck = ref.EarlyStopping.call_args.kwargs, ref.ModelCheckpoint.call_args.kwargs
    assert (m_cbs[0].monitor, m_cbs[0].patience, m_cbs[0].restore_best_weights) == (es["monitor"], es["patience"], es["restore_best_weights"])
    assert (m_cbs[1].monitor, m_cbs[1].save_best_only) == (ck["monitor"], ck["save_best_only"])
    assert os.path.basename(m_cbs[1].filepath) == os.path.basename(ck["filepath"]) == \
        f"unet_adaptive_scale_new_loss0.50_depth{r_cfg['depth']}.keras"
    assert os.path.basename(str(m_cbs[2].dir)) == os.path.basename(ref.BackupAndRestore.call_args.args[0]) == "train_backup"
    tb = ref.tf.keras.callbacks.TensorBoard.call_args.kwargs
    assert os.path.basename(m_cbs[3].log_dir) == os.path.basename(tb["log_dir"]) == "t"
    # compile(...): Adam with the command-line learning rate, the named loss, the psnr metric, no XLA
    assert ref.tf.keras.optimizers.Adam.call_args.kwargs == {"learning_rate": 2e-4}
    opt = rec["model"].optimizer
    assert type(opt).__name__ == "Adam" and opt.learning_rate == 2e-4 and (opt.beta_1, opt.beta_2, opt.epsilon) == (0.9, 0.999, 1e-7)
    assert r_compile.kwargs["jit_compile"] is False and r_compile.kwargs["loss"].__name__ == "charbonnier_loss"
    assert [m.__name__ for m in r_compile.kwargs["metrics"]] == ["psnr"]
    assert type(rec["model"].loss).__name__ == "SRLoss" and list(rec["model"].loss.metric_names) == ["psnr"]
    # what the run prints at the end
    pick = lambda text, head: [ln for ln in text.splitlines() if ln.startswith(head)]
    assert pick(mine_out, "Model info:") == pick(ref_out, "Model info:")
    assert pick(mine_out, "Training complete.") == pick(ref_out, "Training complete.") == ["Training complete."]
    assert [os.path.basename(ln) for ln in pick(mine_out, "Checkpoint saved to:")] == \
        [os.path.basename(ln) for ln in pick(ref_out, "Checkpoint saved to:")]


@pytest.mark.parametrize("protocol", ["A", "B"])
def test_seg_training_protocol_matches_the_reference_trainer(sides, tmp_path, monkeypatch, capsys, protocol):
    """The adaptive segmentation trainer (Segmenation/code/train_adaptive_unet.py:463-560), protocols A and B: the reference's
    train(args) on stand-ins next to this repo's train(args) with fit / evaluate replaced by recorders, same ISIC-style
    directories and command line.  Compared: config.json, the fit arguments, the callbacks, the optimizer (cosine schedule of
    protocol A), the loss the protocol builds, the printed validation metrics."""
    cv2 = pytest.importorskip("cv2")
    import json
    ref, mine = sides["Segmenation/code/train_adaptive_unet.py"]
    from b200unet.keras import clear_session, model as MM
    rng = np.random.default_rng(1)
    dirs = {}
    for split, n in (("train", 11), ("val", 5)):
        for kind in ("img", "msk"):
            d = dirs[split, kind] = tmp_path / f"{split}_{kind}"
            d.mkdir()
        for i in range(n):
            cv2.imwrite(str(dirs[split, "img"] / f"ISIC_{i:07d}.jpg"), rng.integers(0, 256, (40, 50, 3), dtype=np.uint8))
            cv2.imwrite(str(dirs[split, "msk"] / f"ISIC_{i:07d}_segmentation.png"),
                        (rng.random((40, 50)) > 0.5).astype(np.uint8) * 255)

    def argv(tag):
        return ["--protocol", protocol, "--train_images", str(dirs["train", "img"]), "--train_masks", str(dirs["train", "msk"]),
                "--val_images", str(dirs["val", "img"]), "--val_masks", str(dirs["val", "msk"]), "--image_size", "32",
                "--base_channels", "64", "--depth", "2", "--seed", "5", "--run_name", "r",
                "--model_dir", str(tmp_path / tag / "m"), "--log_dir", str(tmp_path / tag / "l")]

    metrics = {"loss": 0.4321, "dice_metric": 0.75, "iou_metric": 0.6}
    # ---- reference on stand-ins
    for name in ("Model", "EarlyStopping", "ModelCheckpoint", "BackupAndRestore", "TensorBoard", "CosineDecay"):
        monkeypatch.setattr(ref, name, MagicMock(name=name))
    rmodel = ref.Model.return_value
    rmodel.fit.return_value.history = {"loss": [1.0, 0.5]}
    rmodel.evaluate.return_value = dict(metrics)
    monkeypatch.setattr(sys, "argv", ["train_adaptive_unet.py"] + argv("ref"))
    ref.train(ref.parse_args())
    ref_out = capsys.readouterr().out
    r_cfg = json.loads((tmp_path / "ref" / "l" / "r" / "config.json").read_text())
    r_fit = rmodel.fit.call_args

    # ---- this repo's trainer, GPU work replaced
    rec = {}

    def fake_fit(self, x=None, **kw):
        rec["fit"], rec["model"] = kw, self
        h = MM.History()
        h.epoch, h.history = [0, 1], {"loss": [1.0, 0.5]}
        return h

    clear_session()
    monkeypatch.setattr(MM.Model, "fit", fake_fit)
    monkeypatch.setattr(MM.Model, "evaluate", lambda self, ds, **kw: dict(metrics))
    mine.train(mine.parse_args(argv("mine")))
    mine_out = capsys.readouterr().out
    m_cfg = json.loads((tmp_path / "mine" / "l" / "r" / "config.json").read_text())
    clear_session()

    assert list(m_cfg) == list(r_cfg)
    for k in r_cfg:
        if k != "model_checkpoint":
            assert m_cfg[k] == r_cfg[k], (k, r_cfg[k], m_cfg[k])
    assert os.path.basename(m_cfg["model_checkpoint"]) == os.path.basename(r_cfg["model_checkpoint"]) == "r.keras"
    for k in ("epochs", "verbose"):
        assert rec["fit"][k] == r_fit.kwargs[k], (k, r_fit.kwargs[k], rec["fit"][k])
    # callbacks
    m_cbs, r_cbs = rec["fit"]["callbacks"], r_fit.kwargs["callbacks"]
    want = ["ModelCheckpoint", "BackupAndRestore", "TensorBoard"] + (["EarlyStopping"] if ref.PROTOCOLS[protocol].early_stopping_patience else [])
    assert [type(c).__name__ for c in m_cbs] == want and len(r_cbs) == len(want)
    ck = ref.ModelCheckpoint.call_args.kwargs
    assert (m_cbs[0].monitor, m_cbs[0].save_best_only) == (ck["monitor"], ck["save_best_only"]) == ("val_dice", True)
    if "EarlyStopping" in want:
        es = ref.EarlyStopping.call_args.kwargs
        assert (m_cbs[3].monitor, m_cbs[3].patience, m_cbs[3].restore_best_weights) == (es["monitor"], es["patience"], True)
    else:
        assert not ref.EarlyStopping.called
    # optimizer: Adam on a cosine schedule over epochs * steps (A) or a constant rate (B)
    opt = rec["model"].optimizer
    if ref.PROTOCOLS[protocol].cosine_schedule:
        cd = ref.CosineDecay.call_args.kwargs
        sched = opt.learning_rate
        assert (sched.initial_learning_rate, sched.decay_steps, sched.alpha) == (cd["initial_learning_rate"], cd["decay_steps"], cd["alpha"])
        assert ref.tf.keras.optimizers.Adam.call_args.kwargs["learning_rate"] is ref.CosineDecay.return_value
    else:
        assert not ref.CosineDecay.called
        assert ref.tf.keras.optimizers.Adam.call_args.kwargs == {"learning_rate": opt.learning_rate} == {"learning_rate": 3e-4}
    # loss of the protocol and the metrics
    r_compile = rmodel.compile.call_args.kwargs
    assert r_compile["loss"].__name__ == rec["model"].loss.__name__ == ("hybrid_ce_dice" if protocol == "A" else "bce_dice")
    assert (rec["model"].loss.bw, rec["model"].loss.dw) == ((0.4, 0.6) if protocol == "A" else (0.5, 1.0))
    assert [m.__name__ for m in r_compile["metrics"]] == list(rec["model"].loss.metric_names) == ["dice", "iou"]
    assert r_compile["jit_compile"] is False
    tail = lambda text: text[text.index("Validation metrics:"):]
    assert tail(mine_out) == tail(ref_out)


def test_baseline_trainers_match_the_reference(sides, tmp_path, monkeypatch, capsys):
    """The two fixed-depth baselines: Segmenation/code/unet_vinillia.py:236-296 and Super_resolution/code/u-net-vinillia.py:
    246-290, reference train()/main() on stand-ins next to this repo's with fit / save / evaluation replaced by recorders."""
    cv2 = pytest.importorskip("cv2")
    import torch
    from b200unet.keras import clear_session, model as MM
    rng = np.random.default_rng(2)

    def recorder(rec):
        def fake_fit(self, x=None, **kw):
            rec["fit"], rec["model"], rec["x"] = kw, self, x
            h = MM.History()
            h.epoch, h.history = [0], {"loss": [1.0]}
            return h
        return fake_fit

    # ---------------------------------------------------------------- segmentation baseline
    ref, mine = sides["Segmenation/code/unet_vinillia.py"]
    dirs = {}
    for split, n in (("train", 7), ("val", 3)):
        for kind in ("img", "msk"):
            d = dirs[split, kind] = tmp_path / f"{split}_{kind}"
            d.mkdir()
        for i in range(n):
            cv2.imwrite(str(dirs[split, "img"] / f"ISIC_{i:07d}.jpg"), rng.integers(0, 256, (40, 50, 3), dtype=np.uint8))
            cv2.imwrite(str(dirs[split, "msk"] / f"ISIC_{i:07d}_segmentation.png"), (rng.random((40, 50)) > 0.5).astype(np.uint8) * 255)
    argv = lambda tag: ["--train_image_dir", str(dirs["train", "img"]), "--train_mask_dir", str(dirs["train", "msk"]),
                        "--val_image_dir", str(dirs["val", "img"]), "--val_mask_dir", str(dirs["val", "msk"]),
                        "--image_size", "32", "--batch_size", "2", "--epochs", "4", "--learning_rate", "3e-4", "--depth", "2",
                        "--model_dir", str(tmp_path / tag), "--run_name", "v", "--limit_train", "5"]
    for name in ("Model", "EarlyStopping", "ModelCheckpoint", "ReduceLROnPlateau", "Adam"):
        monkeypatch.setattr(ref, name, MagicMock(name=name))
    monkeypatch.setattr(sys, "argv", ["unet_vinillia.py"] + argv("ref"))
    ref.train(ref.parse_args())
    ref_out = capsys.readouterr().out
    rmodel = ref.Model.return_value
    rec = {}
    clear_session()
    monkeypatch.setattr(MM.Model, "fit", recorder(rec))
    monkeypatch.setattr(MM.Model, "save", lambda self, path: rec.setdefault("saved", str(path)))
    mine.train(mine.parse_args(argv("mine")))
    mine_out = capsys.readouterr().out
    clear_session()
    first = lambda text, head: [ln for ln in text.splitlines() if ln.startswith(head)][0]
    assert first(mine_out, "Loaded ") == first(ref_out, "Loaded ") == "Loaded 5 training samples and 3 validation samples."
    assert os.path.basename(first(mine_out, "Checkpoints will be written to")) == os.path.basename(first(ref_out, "Checkpoints will be written to"))
    r_fit = rmodel.fit.call_args.kwargs
    assert (rec["fit"]["epochs"], rec["fit"]["verbose"]) == (r_fit["epochs"], r_fit["verbose"]) == (4, 2)
    m_cbs = rec["fit"]["callbacks"]
    assert [type(c).__name__ for c in m_cbs] == ["ModelCheckpoint", "EarlyStopping", "ReduceLROnPlateau"]
    ck, es, rl = ref.ModelCheckpoint.call_args.kwargs, ref.EarlyStopping.call_args.kwargs, ref.ReduceLROnPlateau.call_args.kwargs
    assert (m_cbs[0].monitor, m_cbs[0].save_best_only, os.path.basename(m_cbs[0].filepath)) == \
        (ck["monitor"], ck["save_best_only"], os.path.basename(ck["filepath"])) == ("val_dice_coefficient", True, "v_best.keras")
    assert (m_cbs[1].monitor, m_cbs[1].patience, m_cbs[1].restore_best_weights) == (es["monitor"], es["patience"], es["restore_best_weights"])
    assert (m_cbs[2].monitor, m_cbs[2].factor, m_cbs[2].patience, m_cbs[2].min_lr) == (rl["monitor"], rl["factor"], rl["patience"], rl["min_lr"])
    assert ref.Adam.call_args.kwargs == {"learning_rate": rec["model"].optimizer.learning_rate} == {"learning_rate": 3e-4}
    assert "dice_coefficient" in rec["model"].loss.metric_names            # keras logs the metric under the function's name
    assert os.path.basename(rec["saved"]) == os.path.basename(str(rmodel.save.call_args.args[0])) == "v_final.keras"

    # ---------------------------------------------------------------- SR baseline (whole images from an LR and an HR directory)
    ref, mine = sides["Super_resolution/code/u-net-vinillia.py"]
    hr_dir, lr_dir = tmp_path / "hr", tmp_path / "lr"
    hr_dir.mkdir(); lr_dir.mkdir()
    for i in range(10):
        img = rng.integers(0, 256, (48, 48, 3), dtype=np.uint8)
        cv2.imwrite(str(hr_dir / f"{i}.png"), img)
        cv2.imwrite(str(lr_dir / f"{i}.png"), cv2.resize(cv2.resize(img, (24, 24)), (48, 48)))
    argv = lambda tag: ["--high_res_dir", str(hr_dir), "--low_res_dir", str(lr_dir), "--hr_size", "32", "--batch_size", "3",
                        "--epochs", "2", "--learning_rate", "5e-4", "--patience", "4", "--model_dir", str(tmp_path / ("sr_" + tag))]
    for name in ("Model", "EarlyStopping", "ModelCheckpoint", "BackupAndRestore", "Adam", "build_losses", "evaluate", "make_tf_dataset"):
        monkeypatch.setattr(ref, name, MagicMock(name=name))
    ref.build_losses.return_value = ("combined_loss", ["psnr"])
    ref.evaluate.return_value = {"psnr": (1.0, 0.0)}
    monkeypatch.setattr(sys, "argv", ["u-net-vinillia.py"] + argv("ref"))
    ref_args = ref.parse_args()
    ref_args.seed, ref_args.limit = 1234, None          # read by main() but not declared by the reference's parser
    ref.main(ref_args)
    capsys.readouterr()
    rec = {}
    clear_session()
    monkeypatch.setattr(MM.Model, "fit", recorder(rec))
    monkeypatch.setattr(mine, "evaluate", lambda model, ds: {"psnr": (1.0, 0.0)})
    mine.main(mine.parse_args(argv("mine")))
    capsys.readouterr()
    clear_session()
    # the same images in the same splits reach fit / evaluation
    splits = [c.args[2] for c in ref.make_tf_dataset.call_args_list]          # (lr, hr, indices, batch, shuffle, seed)
    r_lr, r_hr = ref.make_tf_dataset.call_args_list[0].args[:2]
    assert np.array_equal(rec["x"].lr, r_lr) and np.array_equal(rec["x"].hr, r_hr) and r_hr.shape == (10, 32, 32, 3)
    assert np.array_equal(rec["x"].idx, splits[0]) and np.array_equal(rec["fit"]["validation_data"].idx, splits[1])
    assert [c.args[3] for c in ref.make_tf_dataset.call_args_list] == [3, 3, 3] and rec["x"].bs == 3
    r_fit = ref.Model.return_value.fit.call_args.kwargs
    assert (rec["fit"]["epochs"], rec["fit"]["verbose"]) == (r_fit["epochs"], r_fit["verbose"]) == (2, 2)
    m_cbs = rec["fit"]["callbacks"]
    assert [type(c).__name__ for c in m_cbs] == ["EarlyStopping", "ModelCheckpoint", "BackupAndRestore"]
    es, ck = ref.EarlyStopping.call_args.kwargs, ref.ModelCheckpoint.call_args.kwargs
    assert (m_cbs[0].monitor, m_cbs[0].patience, m_cbs[0].restore_best_weights) == (es["monitor"], es["patience"], es["restore_best_weights"])
    assert (m_cbs[1].monitor, m_cbs[1].save_best_only, os.path.basename(m_cbs[1].filepath)) == \
        (ck["monitor"], ck["save_best_only"], os.path.basename(ck["filepath"])) == ("val_loss", True, "unet_vanilla_best.keras")
    assert os.path.basename(str(m_cbs[2].dir)) == os.path.basename(ref.BackupAndRestore.call_args.args[0]) == "train_backup"
    assert ref.Adam.call_args.kwargs == {"learning_rate": rec["model"].optimizer.learning_rate} == {"learning_rate": 5e-4}
    assert rec["model"].name == "U-Net_SR_32x32"


def test_offline_evaluator_main_matches_and_real_artefacts_agree(sides, tmp_path, monkeypatch, capsys):
    """evaluate_model.py end to end (model loading and the metric loop replaced on both sides by the same recorder): the
    reference's main() and this repo's main(argv) over the same image directory write the same config.json (but the time
    stamp), metrics.json and per_image_metrics.csv, patch labels included, and print the same report.  Then the REAL
    artefacts the reference shipped (experiments/*/evaluation/*: produced by its evaluator on DIV2K) are read: same keys in
    the same order as this repo's files, the recorded eval_shave equals infer_eval_shave(scale), the label format matches."""
    cv2 = pytest.importorskip("cv2")
    import glob
    import json
    import re
    ref, mine = sides["Super_resolution/code/evaluate_model.py"]
    data = tmp_path / "hr"
    data.mkdir()
    rng = np.random.default_rng(4)
    for i in (801, 802, 810):
        cv2.imwrite(str(data / f"{i:04d}.png"), rng.integers(0, 256, (70 + i % 7, 100, 3), dtype=np.uint8))

    def fake_eval(mod):
        def evaluate(model, dataset, eval_shave):
            rows, k = [], 0
            for lr, hr in dataset:
                assert lr.shape == hr.shape and lr.shape[1:] == (32, 32, 3)
                for j in range(len(hr)):
                    rows.append({"index": k, "psnr_y": 30.0 + float(np.mean(hr[j])), "ssim_y": 0.9, "msssim_y": float("nan"),
                                 "mse_y": float(np.mean((lr[j] - hr[j]) ** 2))})
                    k += 1
            return mod.EvalResults(1e-3, 1e-4, 30.0, 1.0, 0.9, 0.01, float("nan"), float("nan"), len(rows)), rows
        return evaluate

    argv = lambda tag: ["--model-path", str(tmp_path / "m.keras"), "--scale", "0.3", "--hr-dir", str(data), "--patch-size", "32",
                        "--eval-stride", "24", "--batch-size", "4", "--output-dir", str(tmp_path / tag), "--run-name", "e"]
    outs = {}
    for tag, mod in (("ref", ref), ("mine", mine)):
        monkeypatch.setattr(mod, "load_checkpoint_model", lambda *a, **k: "model")
        monkeypatch.setattr(mod, "evaluate", fake_eval(mod))
        if tag == "ref":
            # the reference wraps its generator in tf.data; hand it the host mirror's dataset (same generator, same labels:
            # tests/test_pipeline_cpu.py) so that its main() has batches to iterate
            from b200unet.shared.pipeline import make_eval_patch_dataset
            monkeypatch.setattr(ref, "make_eval_patch_dataset", make_eval_patch_dataset)
            monkeypatch.setattr(sys, "argv", ["evaluate_model.py"] + argv(tag))
            ref.main()
        else:
            mine.main(argv(tag) + ["--host-pipeline"])
        outs[tag] = capsys.readouterr().out
    files = {tag: {p.name: p.read_text() for p in sorted((tmp_path / tag / "e").iterdir())} for tag in ("ref", "mine")}
    assert set(files["ref"]) == {"config.json", "metrics.json", "per_image_metrics.csv"}
    assert files["mine"]["metrics.json"] == files["ref"]["metrics.json"]
    assert files["mine"]["per_image_metrics.csv"] == files["ref"]["per_image_metrics.csv"]
    rc, mc = json.loads(files["ref"]["config.json"]), json.loads(files["mine"]["config.json"])
    assert list(rc) == list(mc) and {k: v for k, v in rc.items() if k != "created_at"} == {k: v for k, v in mc.items() if k != "created_at"}
    assert rc["eval_shave"] == 6 and rc["samples"] > 10 and rc["images"] == 3
    strip = lambda text: [ln for ln in text.splitlines() if not ln.startswith("[done]")]
    assert strip(outs["mine"]) == strip(outs["ref"]) and outs["ref"].startswith("Evaluated ")

    # the reference's own dataset factories (tf.data calls land on stand-ins): patch counts and labels equal the mirror's
    from b200unet.shared import pipeline as PL
    RP = sys.modules["shared.pipeline"]
    pngs = PL.sorted_alphanumeric(glob.glob(str(data / "*.png")))
    for stride in (None, 24, 40):
        _, r_total, r_labels = RP.make_eval_patch_dataset(pngs, patch_size=32, scale=0.3, batch_size=4, stride=stride)
        _, m_total, m_labels = PL.make_eval_patch_dataset(pngs, 32, 0.3, 4, stride=stride)
        assert (r_total, r_labels) == (m_total, m_labels) and r_total == len(r_labels) > 0
    assert RP.make_training_patch_dataset(pngs, 32, 5, 0.5, 4, seed=1)[1] == PL.make_training_patch_dataset(pngs, 32, 5, 0.5, 4, 1)[1] == 15
    for bad in (dict(hr_files=[], patch_size=32, patches_per_image=1), dict(hr_files=pngs, patch_size=32, patches_per_image=0)):
        for fn in (RP.make_training_patch_dataset, PL.make_training_patch_dataset):
            with pytest.raises(ValueError):
                fn(bad["hr_files"], bad["patch_size"], bad["patches_per_image"], 0.5, 4, 0)

    # the artefacts of the reference's own evaluation runs
    real = sorted(glob.glob(os.path.join(REF, "Super_resolution/experiments/*/evaluation/*/config.json")))
    assert len(real) >= 15
    for cfg_path in real:
        cfg = json.loads(open(cfg_path).read())
        assert list(cfg) == list(mc), cfg_path
        assert cfg["eval_shave"] == mine.infer_eval_shave(cfg["scale"], None), cfg_path
        met = json.loads(open(os.path.join(os.path.dirname(cfg_path), "metrics.json")).read())
        assert list(met) == list(json.loads(files["mine"]["metrics.json"])) and met["samples"] == cfg["samples"]
        with open(os.path.join(os.path.dirname(cfg_path), "per_image_metrics.csv")) as handle:
            head, first = handle.readline().strip(), handle.readline().strip()
        assert head == files["mine"]["per_image_metrics.csv"].splitlines()[0]
        assert re.match(r"^0,\d{4}\.png#patch0000,", first), first
        assert re.match(r"^0,\d{4}\.png#patch0000,", files["mine"]["per_image_metrics.csv"].splitlines()[1])
        # run-level statistics: this repo's aggregation of the reference's per-patch values reproduces its metrics.json
        import csv
        import dataclasses
        with open(os.path.join(os.path.dirname(cfg_path), "per_image_metrics.csv")) as handle:
            rows = list(csv.DictReader(handle))
        vals = {k: [np.array([float(r[k + "_y"]) for r in rows], dtype=np.float32)] for k in ("psnr", "ssim", "msssim", "mse")}
        got = dataclasses.asdict(mine.summarise(vals, len(rows)))
        for key, want in met.items():
            if isinstance(want, float) and not np.isfinite(want):
                assert not np.isfinite(got[key]), (cfg_path, key)            # a patch with zero error: PSNR = inf
            else:
                assert got[key] == pytest.approx(want, rel=1e-12, abs=0.0), (cfg_path, key, want, got[key])
