"""Oracle pinned against the only machine-checkable fixtures the reference holds for this path:
the 15 keras ``model.summary()`` dumps under Super_resolution/experiments/*/model_summary/*.txt
(parameter totals per depth and the ceil() size chains), plus self-consistency of the resize rule."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import keras_ops as K, models as M, resize_np as R

# totals: e.g. experiment_1_constant_depth_3/model_summary/exp1_depth3_scale0.50_model_summary.txt (last lines)
PARAM_TOTALS = {1: 520_003, 2: 2_144_451, 3: 8_637_379, 4: 34_599_363, 5: 138_427_843}


@pytest.mark.parametrize("depth", sorted(PARAM_TOTALS))
def test_param_totals_match_reference_summaries(depth):
    assert M.param_count(M.sr_unet_spec(depth)) == PARAM_TOTALS[depth]


def test_size_chain_matches_reference_summary():
    # exp2_adaptive_depth_scale0.70_model_summary.txt: 256 -> 180 -> 126 -> 89 -> 63 -> 45
    assert R.size_chain(256, 0.7, 5) == [256, 180, 126, 89, 63, 45]
    assert R.size_chain(128, 0.25, 4) == [128, 32, 8, 2, 1]
    # float32 product: ceil(f32(50) * f32(0.3)) = 16, not ceil(50 * 0.3) = 15
    assert R.resized_extent(50, 0.3) == 16 and math.ceil(50 * 0.3) == 15


def test_reference_summaries_if_present():
    """When /root/reference is mounted (build container), parse every summary file and check the
    total and the enc_down output extent against the oracle's builders."""
    root = "/root/reference/Super_resolution/experiments"
    if not os.path.isdir(root):
        pytest.skip("reference not mounted (GPU box)")
    import glob
    import re
    files = glob.glob(os.path.join(root, "*", "model_summary", "*.txt"))
    assert len(files) == 15
    for f in files:
        txt = open(f, encoding="utf-8").read()
        m = re.search(r'Model: "U-Net_SR_scale([0-9.]+)_depth(\d+)"', txt)
        scale, depth = float(m.group(1)), int(m.group(2))
        total = int(re.search(r"Total params: ([0-9,]+)", txt).group(1).replace(",", ""))
        assert M.param_count(M.sr_unet_spec(depth)) == total, f
        bottleneck = R.size_chain(256, scale, depth)[-1]
        assert re.search(rf"enc_down.*?\(None, {bottleneck}, {bottleneck},", txt.replace("\n", " ")) or depth == 0, f


def test_depth_rules():
    assert [M.custom_depth_from_scale(s, base_resolution=256) for s in (0.2, 0.3, 0.5, 0.7)] == [2, 3, 4, 7]
    assert [M.custom_depth_from_scale(s, base_resolution=128) for s in (0.25, 0.5)] == [2, 3]
    assert [M.infer_depth_from_scale(s) for s in (0.2, 0.3, 0.6)] == [1, 2, 3]
    assert M.estimate_bottleneck_size(256, 0.7, 3) == 88


def test_resize_rule_matches_torch_antialias():
    import torch.nn.functional as F
    for (a, b) in [(256, 180), (126, 89), (77, 24), (128, 32), (8, 2), (2, 1), (32, 128), (1, 2), (45, 63)]:
        x = torch.rand(1, a, a, 2)
        y = K.resize_bilinear(x, b, b)
        y2 = F.interpolate(x.permute(0, 3, 1, 2), size=(b, b), mode="bilinear", antialias=True,
                           align_corners=False).permute(0, 2, 3, 1)
        assert (y - y2).abs().max().item() < 2e-5
    st, wt = R.triangle_spans(8, 4)
    assert np.allclose(wt[1, :4], [0.125, 0.375, 0.375, 0.125])


def test_sr_flops():
    assert abs(M.sr_flops_per_sample(0.5, 3, 64) / 1e9 - 6.811) < 1e-3
    assert abs(M.sr_flops_per_sample(0.25, 4, 128) / 1e9 - 12.331) < 1e-3
    assert abs(M.sr_flops_per_sample(0.25, 5, 128) / 1e9 - 12.539) < 1e-3


def _keras_summary_rows(text):
    """Rows of a keras 3 ``model.summary()`` table: [(layer name, type, output shape, params, connected-to)], cells that
    keras wrapped over several lines re-joined, names / types / inputs it truncated kept with their '…'."""
    rows, cur = [], None
    for line in text.splitlines():
        if line.startswith(("├", "└", "┡")):
            if cur:
                rows.append(cur)
            cur = None
            continue
        if line.startswith("│"):
            cells = [c.strip() for c in line.strip("│").split("│")]
            if cur is None:
                cur = [[c] if c else [] for c in cells]
            else:
                for acc, c in zip(cur, cells):
                    if c:
                        acc.append(c)
    out = []
    for name_cell, shape, params, conn in rows:
        joined = " ".join(name_cell)
        name, _, typ = joined.partition(" (")
        out.append((name.strip(), typ.rstrip(")").strip(), " ".join(shape), int(params[0].replace(",", "")),
                    [c.rstrip(",") for c in conn]))
    return out


def test_reference_summaries_row_by_row():
    """Every row of the reference's 15 keras ``model.summary()`` dumps -- produced by the real Keras implementation of the
    model -- against this repo's builder: layer order, names, types, output shapes, parameter counts, and which layers feed
    which (the concat order [up, skip], the shared resize layers)."""
    root = "/root/reference/Super_resolution/experiments"
    if not os.path.isdir(root):
        pytest.skip("reference not mounted (GPU box)")
    import glob
    import re
    from b200unet import builders as B
    from b200unet.keras import clear_session
    files = sorted(glob.glob(os.path.join(root, "*", "model_summary", "*.txt")))
    assert len(files) == 15
    checked = 0

    def same(ref, mine):          # keras truncates long cells with an ellipsis
        return mine.startswith(ref[:-1]) if ref.endswith("…") else ref == mine

    for f in files:
        txt = open(f, encoding="utf-8").read()
        m = re.search(r'Model: "U-Net_SR_scale([0-9.]+)_depth(\d+)"', txt)
        scale, depth = float(m.group(1)), int(m.group(2))
        clear_session()
        model, _ = B.build_super_resolution_unet(scale, depth_override=depth, input_size=256)
        assert f'Model: "{model.name}"' in txt
        mine = []
        model.summary(print_fn=mine.append)
        mine_rows = []
        for line in mine:
            parts = [p.strip() for p in line.split("|")]
            if len(parts) == 4 and parts[0] not in ("Layer (type)",) and not set(parts[0]) <= set("-+"):
                name, _, typ = parts[0].partition(" (")
                mine_rows.append((name, typ.rstrip(")"), parts[1], int(parts[2].replace(",", "")), parts[3].split(", ")))
        ref_rows = _keras_summary_rows(txt)
        assert len(ref_rows) == len(mine_rows) > 20, f
        for (rn, rt, rs, rp, rc), (mn, mt, ms, mp, mc) in zip(ref_rows, mine_rows):
            if rt == "ClipAdd":      # some dumps predate the rename; the reference keeps ClipAdd as an alias (custom_layers.py:142)
                rt = "ClippedResidualAdd"
            assert same(rn, mn) and same(rt, mt), (f, rn, rt, mn, mt)
            assert rp == mp, (f, rn, rp, mp)
            assert rs == ms, (f, rn, rs, ms)          # a shared layer's row shows the shape of its LAST call, as keras does
            if rc != ["-"]:
                assert len(rc) == len(mc) and all(same(a, b) for a, b in zip(rc, mc)), (f, rn, rc, mc)
            checked += 1
        for key in ("Total params", "Trainable params", "Non-trainable params"):
            want = re.search(rf"{key}: ([0-9,]+)", txt).group(1)
            assert any(re.search(rf"{key}: {want}\b", line) for line in mine), (f, key, want)
    clear_session()
    assert checked == 885          # rows of the 15 dumps


def test_resize_rule_matches_pillow():
    """A third independent implementation of the antialiased triangle filter: Pillow's float ('F' mode) BILINEAR resize
    (support widened by the shrink factor, weights at pixel centres, renormalised at the borders -- the algorithm
    tf.image.resize(antialias=True) states it follows) against the ScaleAndTranslate restatement, shrinking and enlarging,
    non-square."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    for a, b in [(256, 180), (126, 89), (77, 24), (128, 32), (8, 2), (2, 1), (32, 128), (45, 63), (100, 100), (256, 52),
                 (63, 45), (1, 2)]:
        x = rng.random((a, a + 3), dtype=np.float32)
        oh, ow = b, max(1, int(round((a + 3) * b / a)))
        want = np.asarray(Image.fromarray(x, mode="F").resize((ow, oh), Image.BILINEAR), dtype=np.float32)
        got = K.resize_bilinear(torch.from_numpy(x)[None, :, :, None], oh, ow, True)[0, :, :, 0].numpy()
        assert np.abs(got - want).max() < 2e-5, (a, b)
