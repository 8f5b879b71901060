"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol, the
host-side resample planner matches the oracle, the Keras-shaped builders reproduce the reference's
parameter totals / layer names, and the product path refuses to run without a CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from b200unet import _ffi
    lib = _ffi.load()
    header = open(os.path.join(ROOT, "include", "b200_unet.h")).read()
    declared = set(re.findall(r"\b(b200_[a-zA-Z0-9_]+)\s*\(", header))
    declared -= {"b200_tensor", "b200_filter"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/b200_unet.h but not exported"
        assert name in _ffi.SIGNATURES, f"{name} has no ctypes prototype"
    assert b"sm_100a" in lib.b200_version()


def test_struct_layout_matches_header():
    from b200unet import _ffi
    assert ctypes.sizeof(_ffi.Tensor) == 8 + 4 * 4 + 3 * 8 + 8
    assert ctypes.sizeof(_ffi.Filter) == 16 + 4 * 4 + 8


def test_resample_plan_matches_oracle_tables():
    import b200unet.ops as ops
    from oracle import resize_np
    for (a, b) in [(256, 180), (180, 126), (126, 89), (89, 63), (63, 45), (128, 32), (32, 8), (8, 2), (2, 1), (1, 2),
                   (32, 128), (50, 16), (64, 64)]:
        for aa in (True, False):
            plan = ops.ResamplePlan(a, b, aa, "cpu")
            if a == b:
                assert plan.taps == 1 and np.array_equal(plan.host[0], np.arange(a))
                continue
            st, wt = resize_np.triangle_spans(a, b, aa)
            assert np.array_equal(plan.host[0], st) and np.array_equal(plan.host[1], wt), (a, b, aa)
            # transpose tables reproduce the dense matrix transposed
            mat = resize_np.resize_matrix(a, b, aa)
            t_st, t_wt = plan.host[2], plan.host[3]
            back = np.zeros((a, b), np.float32)
            for i in range(a):
                for k in range(t_wt.shape[1]):
                    if t_wt[i, k] != 0:
                        back[i, t_st[i] + k] += t_wt[i, k]
            assert np.array_equal(back, mat.T)
    assert ops.resize_extent(50, 0.3) == 16 and ops.resize_extent(256, 0.7) == 180 and ops.resize_extent(1, 0.25) == 1


def test_resample_compaction_keeps_the_operator_and_picks_a_marching_kernel():
    """The compacted device tables (leading zeros shifted out, re-packed) describe the same dense
    resize matrix, and b200_resample_mode classifies the row axis: up (taps <= 3) for up-sampling and
    transposed down-sampling, down (4 or 6 slots) for antialiased down-sampling and transposed up-sampling."""
    import b200unet.ops as ops
    from oracle import resize_np
    for (a, b) in [(128, 32), (32, 128), (63, 45), (45, 63), (256, 180), (180, 256), (8, 2), (2, 8), (50, 16), (1, 2),
                   (2, 1), (100, 90), (90, 100), (1024, 256)]:
        plan = ops.ResamplePlan(a, b, True, "cpu")
        mat = resize_np.resize_matrix(a, b, True)
        cs, cw, ts, tw = plan.compact_host
        assert cw.shape[1] == plan.taps and tw.shape[1] == plan.t_taps
        fwd = np.zeros((b, a), np.float32)
        for o in range(b):
            for k in range(cw.shape[1]):
                if cw[o, k] != 0:
                    fwd[o, cs[o] + k] += cw[o, k]
        assert np.array_equal(fwd, mat), (a, b)
        back = np.zeros((a, b), np.float32)
        for i in range(a):
            for k in range(tw.shape[1]):
                if tw[i, k] != 0:
                    back[i, ts[i] + k] += tw[i, k]
        assert np.array_equal(back, mat.T), (a, b)
        for st, w, mode in ((cs, cw, plan.mode), (ts, tw, plan.t_mode)):
            taps = w.shape[1]
            if taps <= 3:
                assert mode == 1
            elif mode in (2, 3):
                ns = 4 if mode == 2 else 6
                assert all(st[o + ns] - st[o] >= taps for o in range(len(st) - ns))
        if b > a:
            assert plan.mode == 1 and plan.taps <= 2          # bilinear up-sampling: at most two taps
        if (a, b) == (128, 32):
            assert plan.taps == 8 and plan.mode == 2 and plan.t_mode == 1
        if (a, b) == (32, 128):
            assert plan.t_taps == 8 and plan.t_mode == 2


def test_builders_match_reference_param_totals_and_names():
    from b200unet import builders as B
    from b200unet.keras import clear_session
    for depth, total in {1: 520_003, 2: 2_144_451, 3: 8_637_379, 4: 34_599_363, 5: 138_427_843}.items():
        clear_session()
        model, info = B.build_super_resolution_unet(0.5, depth_override=depth, input_size=256)
        assert model.count_params() == total
        assert model.name == f"U-Net_SR_scale0.50_depth{depth}" and info["depth"] == depth
    names = [ly.name for ly in model.layers]
    assert names[:5] == ["low_res_input", "conv2d", "layer_normalization", "activation", "conv2d_1"]
    assert {"enc_down", "dec_up", "residual_rgb", "enhanced_rgb"} <= set(names)
    clear_session()
    model, info = B.build_super_resolution_unet(0.7, input_size=256)     # adaptive rule: depth 7 at scale 0.7
    assert info["depth"] == 7
    lines = []
    model.summary(print_fn=lines.append)
    assert any("Total params" in ln for ln in lines)


def test_custom_layers_interface():
    from b200unet.shared import custom_layers as CL
    assert CL.ClipAdd is CL.ClippedResidualAdd
    assert {"resize>ResizeByScale", "resize>ResizeToMatch", "utils>ClippedResidualAdd"} <= set(CL.get_custom_objects())
    lay = CL.ResizeByScale(0.7, name="enc_down")
    assert lay.get_config()["scale"] == 0.7 and lay.get_config()["antialias"] is True
    with pytest.raises(ValueError):
        CL.custom_depth_from_scale(1.5)
    assert CL.custom_depth_from_scale(0.5, base_resolution=256) == 4
    assert CL.infer_depth_from_scale(0.3) == 2 and CL.estimate_bottleneck_size(256, 0.7, 3) == 88


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from b200unet import builders as B
    from b200unet._ffi import B200Error
    from b200unet.keras import clear_session
    clear_session()
    model, _ = B.build_super_resolution_unet(0.5, depth_override=1, input_size=16)
    with pytest.raises(B200Error):
        model(np.zeros((1, 16, 16, 3), np.float32))


@pytest.mark.parametrize("first", ["b200unet.shared.custom_layers", "b200unet.keras.engine", "b200unet.builders",
                                   "b200unet.shared.pipeline", "b200unet.metrics"])
def test_any_module_can_be_the_first_import(first):
    """The reference's scripts start with ``from shared.custom_layers import ...`` (train_adaptive_unet.py:27-34); the swapped
    import must work as the FIRST import of a fresh interpreter (custom_layers <-> keras.engine import each other)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (f"import sys; sys.path.insert(0, {root!r}); import importlib; importlib.import_module({first!r}); "
            "from b200unet.shared.custom_layers import ResizeByScale, ResizeToMatch, ClippedResidualAdd, ClipAdd; "
            "from b200unet.builders import build_super_resolution_unet; print('ok')")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_backward_stream_schedule(monkeypatch):
    """Model._run_bwd with recording stand-ins for the CUDA streams: every filter-gradient launch goes to the side stream
    behind a wait on the main stream, everything else stays on the main stream in order, the side stream is joined at the end
    of the range; with the per-layer Adam switch a layer's kernel range is updated on the side stream only after a wait
    issued AFTER that layer's dgrad (which reads the weights), and the final optimizer pass covers exactly the rest."""
    import contextlib
    import types
    import torch
    from b200unet.keras.model import Model
    from b200unet.parallel import complement_ranges

    log = []

    class FakeStream:
        def __init__(self, name):
            self.name = name

        def wait_stream(self, other):
            log.append(("wait", self.name, other.name))

    main, cur = FakeStream("main"), []
    cur.append(main)

    @contextlib.contextmanager
    def use(stream):
        cur.append(stream)
        try:
            yield
        finally:
            cur.pop()

    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: cur[-1])
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: FakeStream("side"))
    monkeypatch.setattr(torch.cuda, "stream", use)

    tags = ["bias_act", "wgrad:tc", "dgrad:tc", "ln", "wgrad:tc", "dgrad:tc", "resize", "ln", "wgrad:simt"]
    writes = [[(900, 10)], [(200, 100)], [], [(910, 8)], [(100, 100)], [], [], [(918, 4)], [(0, 100)]]
    plan = types.SimpleNamespace(bwd_tags=tags, bwd_writes=writes,
                                 bwd_steps=[(lambda i=i: log.append(("run", i, tags[i], cur[-1].name))) for i in range(len(tags))])

    class Opt:
        def advance(self):
            log.append(("advance", cur[-1].name))

        def apply_ranges(self, model, ranges):
            log.append(("adam", tuple(ranges), cur[-1].name))

        def apply(self, model, ranges=None):
            log.append(("adam_all", cur[-1].name))

    def run(overlap, adam):
        log.clear()
        me = types.SimpleNamespace(overlap_wgrad=overlap, overlap_adam=adam, _side_stream=None, _dist=None, optimizer=Opt(),
                                   G=types.SimpleNamespace(numel=lambda: 1000), _update_ranges=lambda plan: None)
        Model._run_bwd(me, plan, 0, len(tags))
        Model._apply_optimizer(me, plan)
        return list(log)

    # single stream: plain order, one optimizer pass
    ev = run(False, False)
    assert [e[1] for e in ev if e[0] == "run"] == list(range(len(tags))) and all(e[3] == "main" for e in ev if e[0] == "run")
    assert ev[-1] == ("adam_all", "main") and not any(e[0] == "wait" for e in ev)

    # wgrad on the side stream
    ev = run(True, False)
    runs = [e for e in ev if e[0] == "run"]
    assert [e[1] for e in runs] == list(range(len(tags)))
    for e in runs:
        assert e[3] == ("side" if e[2].startswith("wgrad") else "main"), e
    for k, e in enumerate(ev):
        if e[0] == "run" and e[2].startswith("wgrad"):
            assert ev[k - 1] == ("wait", "side", "main")               # its dz is final
    assert ev[-2] == ("wait", "main", "side") and ev[-1] == ("adam_all", "main")   # joined before the optimizer

    # + per-layer Adam behind the wgrad kernels
    ev = run(True, True)
    assert ev[0] == ("advance", "main")
    pos = {("run", i): k for k, e in enumerate(ev) if e[0] == "run" for i in [e[1]]}
    done = []
    for k, e in enumerate(ev):
        if e[0] == "adam" and e[2] == "side":
            for lo, hi in e[1]:
                w = next(i for i, wr in enumerate(writes) if (lo, hi - lo) in wr)     # the wgrad step that wrote this range
                assert tags[w].startswith("wgrad") and pos["run", w] < k                # same stream, after its wgrad
                dgrads = [i for i in range(w + 1, len(tags)) if tags[i].startswith("dgrad")]
                if dgrads:                                                              # ... and after a wait issued behind its dgrad
                    d = pos["run", dgrads[0]]
                    assert any(ev[j] == ("wait", "side", "main") for j in range(d + 1, k)), (e, d, k)
                done.append((lo, hi))
    assert sorted(done) == [(0, 100), (100, 200), (200, 300)]
    final = [e for e in ev if e[0] == "adam" and e[2] == "main"]
    assert len(final) == 1 and list(final[0][1]) == complement_ranges(done, 1000) == [(300, 1000)]
    assert ev.index(("wait", "main", "side")) < ev.index(final[0])


def test_lowering_structure_on_cpu():
    """keras/engine.lower (pure graph -> op list): ReLU folded into the producing conv / norm, conv bias gradients routed
    through the following norm, both concat inputs aliased into the concat buffer (written in place, no copy op), and -- under
    the bf16 policy -- every 3x3 convolution of the four reference nets selects the tcgen05 kernels (the RGB stems go through
    the im2col path the Plan adds), including the 32-channel level of the default segmentation width."""
    import torch
    from b200unet import builders as B
    from b200unet.keras import clear_session
    from b200unet.keras.engine import Plan, lower
    from b200unet._ffi import ACT_RELU
    clear_session()
    nets = {
        "sr": B.build_super_resolution_unet(0.25, depth_override=4, input_size=128)[0],
        "sr_vanilla": B.build_vanilla_super_resolution_unet((64, 64, 3), 64, 2),
        "seg_adaptive": B.build_adaptive_depth_unet(64, 64, 2),
        "seg_base32": B.build_unet(256, num_classes=21, base_channels=32, depth=4),
    }
    for name, model in nets.items():
        model._compute_dtype = torch.bfloat16
        ops_, _ = lower(model)
        kinds = [op.kind for op in ops_]
        assert "activation" not in kinds and set(kinds) <= {"cast_input", "conv", "ln", "bn", "resize", "maxpool", "convT",
                                                            "concat", "clipadd", "softmax"}, (name, set(kinds))
        norms = [op for op in ops_ if op.kind in ("ln", "bn")]
        assert norms and all(op.relu and op.conv_src is not None and op.conv_src.norm_bias for op in norms), name
        concats = [op for op in ops_ if op.kind == "concat"]
        assert concats and all(not op.copies for op in concats), name           # skip-concat written in place
        for op in concats:
            a, b = op.inputs
            assert a.parent is op.output and b.parent is op.output and (a.offset, b.offset) == (0, a.c)   # [up, skip]
        convs3 = [op for op in ops_ if op.kind == "conv" and op.layer.kernel_size == (3, 3)]
        stems = [op for op in convs3 if op.inputs[0].c == 3]
        assert len(stems) == 1 and all(Plan.is_tc(op) for op in convs3 if op not in stems), name
        plain = [op for op in convs3 if not op.norm_bias]                       # decoder post-upsample conv: bias + ReLU epilogue
        assert all(op.act == ACT_RELU for op in plain), name
    assert {op.output.c for op in lower(nets["seg_base32"])[0] if op.kind == "conv"} == {32, 64, 128, 256, 512, 21}
    clear_session()


def test_epoch_logs_are_sample_weighted_and_precision_recall_accumulate():
    """keras weights every batch by its sample count (a ragged last batch counts for less) and its Precision / Recall are
    stateful: ratios of the tp / fp / fn counters accumulated over the epoch, not means of per-batch ratios."""
    from b200unet import builders as B
    from b200unet.keras import clear_session
    from b200unet.keras.losses import BinaryCrossentropy
    clear_session()
    model = B.build_unet(16, num_classes=1, base_channels=32, depth=1)
    model.loss = BinaryCrossentropy(global_dice=True, keras_metrics=True)
    assert model._log_keys() == ["loss", "accuracy", "precision", "recall", "dice_coefficient", "iou", "_tp", "_fp", "_fn"]
    # two batches: 8 samples (loss 1.0, tp 10, fp 0, fn 10 per batch) and 2 samples (loss 3.0, tp 0, fp 6, fn 0)
    sums = {"loss": 1.0 * 8 + 3.0 * 2, "accuracy": 0.5 * 8 + 1.0 * 2, "precision": 1.0 * 8 + 0.0 * 2, "recall": 0.5 * 8 + 0.0 * 2,
            "dice_coefficient": 0.7 * 8 + 0.1 * 2, "iou": 0.5 * 10, "_tp": 10.0, "_fp": 6.0, "_fn": 10.0}
    logs = model._weighted_mean(sums, 10.0)
    assert abs(logs["loss"] - 1.4) < 1e-12 and abs(logs["accuracy"] - 0.6) < 1e-12
    assert abs(logs["precision"] - 10 / 16) < 1e-12 and abs(logs["recall"] - 0.5) < 1e-12      # NOT 0.8 / 0.4 (batch means)
    assert list(logs) == ["loss", "accuracy", "precision", "recall", "dice_coefficient", "iou"]
    assert model._weighted_mean({}, 0.0) == {}
