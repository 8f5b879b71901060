"""Whole-model parity: forward activations, loss, every weight gradient and one Adam update of the
SR U-Net (and the two segmentation U-Nets) against the torch-CPU oracle on identical inputs/weights.

Tolerances: fp32 policy rel-L2 <= 5e-4 per gradient tensor (fp32 accumulation-order noise through
~25 layers; measured 1.2e-4 at the stem); bf16 policy <= 3e-2 on weight gradients and on the
end-to-end output (per-layer kernels are at 1.7e-3; 1-ulp bf16 rounding flips accumulate over ~50
stored tensors), against the oracle run with bf16 storage rounding of the same tensors."""
import numpy as np
import pytest
import torch

from util import relerr

pytestmark = pytest.mark.gpu


def _setup(policy):
    from b200unet.keras import clear_session, mixed_precision
    clear_session()
    mixed_precision.set_global_policy(policy)


def _oracle_step(ws_np, x, t, fwd, loss_fn, rnd, dtype=torch.float32):
    from oracle import models as M
    ws = [torch.tensor(w, dtype=dtype, requires_grad=True) for w in ws_np]
    x, t = x.to(dtype), t.to(dtype)
    if rnd is not None:
        wsr = [rnd(w) if w.dim() == 4 else w for w in ws]   # kernels are stored in the compute dtype
    else:
        wsr = ws
    y = fwd(wsr, x)
    l = loss_fn(t, y)
    l.backward()
    return y.detach(), l.item(), [w.grad if w.grad is not None else torch.zeros_like(w) for w in ws]


@pytest.mark.parametrize("policy", ["float32", "mixed_bfloat16"])
def test_sr_unet_step(policy):
    from b200unet import builders as B
    from b200unet.keras.optimizers import Adam
    from oracle import keras_ops as K, models as M
    _setup(policy)
    scale, depth, P, batch = 0.5, 3, 64, 8
    model, info = B.build_super_resolution_unet(scale, depth_override=depth, input_size=P)
    spec = M.sr_unet_spec(depth)
    ws_np = M.init_weights(spec, seed=1234, randomize_zero_kernels=True, jitter=0.05)
    model.set_weights(ws_np)
    loss, metrics = B.build_losses_and_metrics("charbonnier")
    model.compile(optimizer=Adam(learning_rate=1e-3), loss=loss, metrics=metrics)
    rng = np.random.default_rng(1234)
    hr = rng.random((batch, P, P, 3), dtype=np.float32)
    lr = np.clip(hr + 0.05 * rng.standard_normal(hr.shape).astype(np.float32), 0, 1)
    bf = policy != "float32"
    rnd = M.bf16_round if bf else None
    xt, tt = torch.from_numpy(lr), torch.from_numpy(hr)
    fwd = lambda ws, x: M.sr_unet_forward(ws, x, scale, depth, rnd=(rnd or (lambda v: v)))
    y_ref, l_ref, g_ref = _oracle_step(ws_np, xt, tt, fwd, K.charbonnier_loss, rnd)

    y = model(lr)
    e_out = relerr(y, y_ref)
    logs = model.train_on_batch(lr, hr)
    torch.cuda.synchronize()
    print(f"[{policy}] out relerr {e_out:.3e}  loss {logs['loss']:.6f} vs {l_ref:.6f}  psnr {logs['psnr']:.3f}")
    assert e_out < (3e-2 if bf else 1e-5)
    assert abs(logs["loss"] - l_ref) < (1e-3 if bf else 1e-6) * max(1.0, abs(l_ref))
    assert abs(logs["psnr"] - K.psnr_metric(tt, y_ref).item()) < (0.05 if bf else 1e-3)
    worst = 0.0
    bad = []
    i = 0
    for ly in model.layers:
        for w in ly.weight_specs:
            g = model._grad(ly, w["name"].split("/", 1)[1])
            e = relerr(g, g_ref[i])
            worst = max(worst, e)
            tol = 3e-2 if bf else 5e-4
            small = g_ref[i].abs().max().item() < 1e-7
            print(f"   {w['name']:<40s} grad relerr {e:.3e}")
            if not (e < tol or small):
                bad.append((w["name"], e))
            i += 1
    assert not bad, bad
    # one Adam update: the oracle's update rule applied to the gradient the GPU produced
    i = 0
    for ly in model.layers:
        for w, p_new in zip(ly.weight_specs, ly.get_weights()):
            p0 = torch.tensor(ws_np[i])
            g_gpu = model._grad(ly, w["name"].split("/", 1)[1]).detach().cpu()
            p_ref, _, _ = K.adam_step(p0, g_gpu, torch.zeros_like(p0), torch.zeros_like(p0), 1, 1e-3)
            assert relerr(torch.from_numpy(p_new), p_ref) < 1e-6, w["name"]
            i += 1
    print(f"[{policy}] worst grad relerr {worst:.3e}")


def test_sr_unet_deep_levels_collapse_to_1x1():
    """scale 0.25, depth 4 at 20x20: 20 -> 5 -> 2 -> 1 -> 1: depth
    is forced past the point where the image has shrunk to 1x1 (BASELINE configs 2/3 do this), so the net contains a
    same-size ResizeByScale whose input also feeds a skip connection, stacked 2x2 tiles and flattened 1x1 batches."""
    from b200unet import builders as B
    from b200unet.keras.optimizers import Adam
    from oracle import keras_ops as K, models as M
    _setup("float32")
    scale, depth, P, batch = 0.25, 4, 20, 4
    from oracle import resize_np
    assert list(resize_np.size_chain(P, scale, depth)) == [20, 5, 2, 1, 1]
    model, _ = B.build_super_resolution_unet(scale, depth_override=depth, input_size=P)
    ws_np = M.init_weights(M.sr_unet_spec(depth), seed=4321, randomize_zero_kernels=True, jitter=0.05)
    model.set_weights(ws_np)
    loss, metrics = B.build_losses_and_metrics("l1")
    model.compile(optimizer=Adam(learning_rate=1e-3), loss=loss, metrics=metrics)
    rng = np.random.default_rng(99)
    hr = rng.random((batch, P, P, 3), dtype=np.float32)
    lr = np.clip(hr + 0.05 * rng.standard_normal(hr.shape).astype(np.float32), 0, 1)
    fwd = lambda ws, x: M.sr_unet_forward(ws, x, scale, depth)
    y_ref, l_ref, g_ref = _oracle_step(ws_np, torch.from_numpy(lr), torch.from_numpy(hr), fwd, K.l1_loss, None)
    logs = model.train_on_batch(lr, hr)
    torch.cuda.synchronize()
    assert abs(logs["loss"] - l_ref) < 1e-6 * max(1.0, abs(l_ref))
    i, bad = 0, []
    for ly in model.layers:
        for w in ly.weight_specs:
            e = relerr(model._grad(ly, w["name"].split("/", 1)[1]), g_ref[i])
            if not (e < 2e-3 or g_ref[i].abs().max().item() < 1e-7):
                bad.append((w["name"], e))
            i += 1
    assert not bad, bad
    # the same net under the bf16 policy (tcgen05 + split-K kernels at the 2x2 / 1x1 levels): loss within 1e-3
    _setup("mixed_bfloat16")
    model, _ = B.build_super_resolution_unet(scale, depth_override=depth, input_size=P)
    model.set_weights(ws_np)
    loss, metrics = B.build_losses_and_metrics("l1")
    model.compile(optimizer=Adam(learning_rate=1e-3), loss=loss, metrics=metrics)
    logs = model.train_on_batch(lr, hr)
    assert abs(logs["loss"] - l_ref) < 2e-3 * max(1.0, abs(l_ref))


def test_async_pipelined_steps_match_synchronous_steps():
    """train_on_batch_async (host->device copies staged on a copy stream, results through pinned host memory, one step
    in flight) must produce exactly the losses of the synchronous loop on the same batches."""
    from b200unet import builders as B
    from b200unet.keras.optimizers import Adam
    rng = np.random.default_rng(5)
    batches = []
    for _ in range(6):
        hr = rng.random((4, 32, 32, 3), dtype=np.float32)
        batches.append((np.clip(hr + 0.1 * rng.standard_normal(hr.shape).astype(np.float32), 0, 1), hr))
    runs = []
    for mode in ("sync", "async"):
        _setup("mixed_bfloat16")
        model, _ = B.build_super_resolution_unet(0.5, depth_override=2, input_size=32)
        loss, metrics = B.build_losses_and_metrics("charbonnier")
        model.compile(optimizer=Adam(learning_rate=1e-3), loss=loss, metrics=metrics)
        out = []
        if mode == "sync":
            for lr, hr in batches:
                out.append(model.train_on_batch(lr, hr))
        else:
            prev = None
            for lr, hr in batches:
                h = model.train_on_batch_async(torch.from_numpy(lr).pin_memory(), torch.from_numpy(hr).pin_memory())
                if prev is not None:
                    out.append(prev.result())
                prev = h
            out.append(prev.result())
        runs.append(out)
    print([(round(a["loss"], 7), round(b["loss"], 7)) for a, b in zip(*runs)])
    # step 1 sees identical weights: identical loss; later steps differ in the last bits only (gradient sums use atomics)
    assert abs(runs[0][0]["loss"] - runs[1][0]["loss"]) <= 1e-6
    for a, b in zip(*runs):
        assert abs(a["loss"] - b["loss"]) <= 2e-4 * max(1.0, abs(a["loss"])) and abs(a["psnr"] - b["psnr"]) <= 2e-2, (a, b)


def test_sr_unet_training_reduces_loss():
    """A few steps on a fixed batch: the loss must go down and stay finite (bf16 policy, graph replay)."""
    from b200unet import builders as B
    from b200unet.keras.optimizers import Adam
    _setup("mixed_bfloat16")
    model, _ = B.build_super_resolution_unet(0.5, depth_override=2, input_size=32)
    loss, metrics = B.build_losses_and_metrics("l1")
    model.compile(optimizer=Adam(learning_rate=1e-3), loss=loss, metrics=metrics)
    rng = np.random.default_rng(0)
    hr = rng.random((4, 32, 32, 3), dtype=np.float32)
    lr = np.clip(hr + 0.1 * rng.standard_normal(hr.shape).astype(np.float32), 0, 1)
    # zero head => identity start: loss == mean|hr-lr| exactly
    l0 = model.train_on_batch(lr, hr)["loss"]
    assert abs(l0 - np.abs(hr - lr).mean()) < 2e-3
    ls = [model.train_on_batch(lr, hr)["loss"] for _ in range(30)]
    print("losses", l0, ls[-1])
    assert np.isfinite(ls).all() and ls[-1] < l0


@pytest.mark.parametrize("policy", ["float32", "mixed_bfloat16"])
def test_seg_adaptive_step(policy):
    from b200unet import builders as B
    from b200unet.keras import losses as LS
    from b200unet.keras.optimizers import Adam
    from oracle import keras_ops as K, models as M
    _setup(policy)
    depth, base, P, batch = 2, 64, 32, 4
    model = B.build_adaptive_depth_unet(P, base, depth)
    spec = M.seg_adaptive_spec(depth, base)
    ws_np = M.init_weights(spec, seed=7, jitter=0.05)
    model.set_weights(ws_np)
    model.compile(optimizer=Adam(1e-3), loss=LS.make_hybrid_ce_dice_loss(0.4, 0.6), metrics=[LS.dice_metric, LS.iou_metric])
    rng = np.random.default_rng(3)
    x = rng.random((batch, P, P, 3), dtype=np.float32)
    t = (rng.random((batch, P, P, 1)) > 0.5).astype(np.float32)
    bf = policy != "float32"
    # bf16 policy: the oracle rounds every stored activation AND its gradient (M.bf16_storage), as the kernels do
    rnd = M.bf16_storage if bf else (lambda v: v)
    new_stats = []
    fwd = lambda ws, xx: M.seg_adaptive_forward(ws, rnd(xx), depth, True, rnd, new_stats)   # (bf16: the input is cast too)
    y_ref, l_ref, g_ref = _oracle_step(ws_np, torch.from_numpy(x), torch.from_numpy(t), fwd,
                                       lambda tt, yy: K.bce_dice_loss(tt, yy, 0.4, 0.6), M.bf16_round if bf else None,
                                       dtype=torch.float32 if bf else torch.float64)
    if not bf:   # how far the fp32 ORACLE itself is from the fp64 truth (conditioning of BatchNorm backward)
        _, _, g32 = _oracle_step(ws_np, torch.from_numpy(x), torch.from_numpy(t), fwd,
                                 lambda tt, yy: K.bce_dice_loss(tt, yy, 0.4, 0.6), None)
        print("   fp32-oracle vs fp64-oracle worst grad relerr:",
              max(relerr(a, b) for a, b in zip(g32, g_ref) if b.abs().max() > 1e-7))
    g_alt = None
    if bf:
        # the policy's own noise band: the SAME oracle, same bf16 storage points, evaluated in float64 instead of float32
        # arithmetic.  Pre-rounding values then differ by ~1e-7, a few stored elements round the other way, and
        # BatchNorm backward (d - mean(d) - xhat * mean(d * xhat): a small difference of nearly equal terms on this net;
        # tools/bn_debug.py shows dy within 4e-3 and dz 8e-2 off at the LAST block) amplifies it.  How far the two oracle
        # runs sit from each other is how far two faithful implementations of the policy sit from each other.
        fwd_alt = lambda ws, xx: M.seg_adaptive_forward(ws, rnd(xx), depth, True, rnd, [])
        _, _, g_alt = _oracle_step(ws_np, torch.from_numpy(x), torch.from_numpy(t), fwd_alt,
                                   lambda tt, yy: K.bce_dice_loss(tt, yy, 0.4, 0.6), M.bf16_round, dtype=torch.float64)
    logs = model.train_on_batch(x, t)
    print(f"[{policy}] seg loss {logs['loss']:.6f} vs {l_ref:.6f}")
    assert abs(logs["loss"] - l_ref) < (2e-2 if bf else 1e-5)
    i = 0
    bad = []
    for ly in model.layers:
        for w in ly.weight_specs:
            nm = w["name"].split("/", 1)[1]
            if w["trainable"]:
                e = relerr(model._grad(ly, nm), g_ref[i])
                is_pre_bn_bias = nm == "bias" and ly.name != "lesion_mask"
                print(f"   {w['name']:<40s} grad relerr {e:.3e}")
                # bf16 policy: BatchNorm backward is ill-conditioned on this net (see g_alt above): bf16 rounding flips in
                # the forward pass move the weight gradients by several percent in the ORACLE ITSELF (printed band).
                # The 3e-2 tolerance of SURVEY 8c holds for this net under the fp32 policy (4e-6 vs fp64) and per
                # kernel (test_batchnorm: 1e-2 on identical inputs); end to end the kernels must sit inside the band the
                # policy itself spans: within 3e-2, or no further from the oracle than 2 x the distance between the two
                # oracle evaluations and pointing the same way (cos > 0.95).
                if bf:
                    band = relerr(g_alt[i], g_ref[i])
                    a, b = model._grad(ly, nm).detach().double().cpu().flatten(), g_ref[i].double().flatten()
                    cos = float((a @ b) / (a.norm() * b.norm() + 1e-30))
                    print(f"      policy band (oracle vs oracle) {band:.3e}, cos {cos:.4f}")
                    ok = (e < 3e-2) or (e <= 2.0 * band + 1e-2 and cos > 0.95)
                else:
                    # fp32 policy: ReLU masks / max-pool arg-max are discrete and BatchNorm backward is
                    # ill-conditioned, so ANY fp32 evaluation order moves these gradients: the fp32 torch oracle
                    # itself sits 2.8e-3 from the fp64 oracle on this net (printed above)
                    ok = e < 2e-2
                if not (ok or is_pre_bn_bias or g_ref[i].abs().max() < 1e-7):
                    bad.append((w["name"], e))
            i += 1
    assert not bad, bad
    # moving statistics updated as keras does
    k = 0
    for ly in model.layers:
        if type(ly).__name__ == "BatchNormalization":
            mm, mv = ly.get_weights()[2:]
            assert relerr(torch.from_numpy(mm), new_stats[2 * k]) < (2e-2 if bf else 1e-5)
            assert relerr(torch.from_numpy(mv), new_stats[2 * k + 1]) < (2e-2 if bf else 1e-5)
            k += 1


@pytest.mark.parametrize("num_classes", [1, 5])
def test_seg_vanilla_step(num_classes):
    from b200unet import builders as B
    from b200unet.keras import losses as LS
    from b200unet.keras.optimizers import Adam
    from oracle import keras_ops as K, models as M
    _setup("float32")
    depth, base, P, batch = 2, 32, 16, 2
    model = B.build_unet(P, num_classes=num_classes, base_channels=base, depth=depth)
    ws_np = M.init_weights(M.seg_vanilla_spec(depth, base, num_classes), seed=11, jitter=0.05)
    model.set_weights(ws_np)
    rng = np.random.default_rng(5)
    x = rng.random((batch, P, P, 3), dtype=np.float32)
    if num_classes == 1:
        model.compile(optimizer=Adam(1e-3), loss=LS.BinaryCrossentropy())
        t = (rng.random((batch, P, P, 1)) > 0.5).astype(np.float32)
        loss_fn, tt = K.binary_crossentropy, torch.from_numpy(t)
    else:
        model.compile(optimizer=Adam(1e-3), loss=LS.CategoricalCrossentropy())
        t = rng.integers(0, num_classes, (batch, P, P)).astype(np.int32)
        tt = torch.nn.functional.one_hot(torch.from_numpy(t).long(), num_classes).float()
        loss_fn = K.categorical_crossentropy
    fwd = lambda ws, xx: M.seg_vanilla_forward(ws, xx, depth, num_classes)
    y_ref, l_ref, g_ref = _oracle_step(ws_np, torch.from_numpy(x), tt, fwd, loss_fn, None)
    logs = model.train_on_batch(x, t)
    assert abs(logs["loss"] - l_ref) < 1e-5 * max(1, abs(l_ref))
    i = 0
    for ly in model.layers:
        for w in ly.weight_specs:
            e = relerr(model._grad(ly, w["name"].split("/", 1)[1]), g_ref[i])
            assert e < 2e-4 or g_ref[i].abs().max() < 1e-7, (w["name"], e)
            i += 1


def test_save_load_roundtrip(tmp_path):
    """`.keras` archives in the keras-3 layout (zip: config.json, metadata.json, model.weights.h5 with
    layers/<class_snake[_k]>/vars/<i> and optimizer/vars/<i>): weights and Adam state survive a round trip, the bare
    `.weights.h5` form loads too, and a checkpoint of another architecture is refused with the layer named
    (/root/reference: train_adaptive_unet.py:496-531, 617; evaluate_model.py:57-91)."""
    import io, json, zipfile
    from b200unet import builders as B, h5lite
    from b200unet._ffi import B200Error
    from b200unet.keras import clear_session
    from b200unet.keras.optimizers import Adam
    _setup("float32")
    m1, _ = B.build_super_resolution_unet(0.5, depth_override=1, input_size=16)
    loss, metrics = B.build_losses_and_metrics("charbonnier")
    m1.compile(optimizer=Adam(1e-3), loss=loss, metrics=metrics)
    rng = np.random.default_rng(0)
    hr = rng.random((2, 16, 16, 3), dtype=np.float32)
    head = m1.get_layer("residual_rgb")
    m1._ensure_built()
    w = m1.get_weights(); w[-2] = rng.uniform(-0.2, 0.2, w[-2].shape).astype(np.float32); m1.set_weights(w)
    for _ in range(3):
        m1.train_on_batch(hr * 0.9, hr)
    path = tmp_path / "ckpt.keras"
    m1.save(path)
    with zipfile.ZipFile(path) as z:
        assert sorted(z.namelist()) == ["config.json", "metadata.json", "model.weights.h5"]
        assert json.loads(z.read("metadata.json"))["keras_version"] == "3.3.3"
        blob = z.read("model.weights.h5")
    h5lite.validate(blob)
    tree = h5lite.read_h5(blob)
    assert tree["layers"]["conv2d"]["vars"]["0"].shape == (3, 3, 3, 64) and tree["layers"]["input_layer"]["vars"] == {}
    assert int(tree["optimizer"]["vars"]["0"]) == 3
    clear_session()
    m2, _ = B.build_super_resolution_unet(0.5, depth_override=1, input_size=16)
    m2.compile(optimizer=Adam(1e-3), loss=loss, metrics=metrics)
    m2.load_weights(path)
    for a, b in zip(m1.get_weights(), m2.get_weights()):
        assert np.array_equal(a, b)
    assert m2.optimizer.iterations == 3
    assert torch.equal(m1.optimizer._state["m"], m2.optimizer._state["m"])
    assert torch.equal(m1.optimizer._state["v"], m2.optimizer._state["v"])
    l1 = m1.train_on_batch(hr * 0.9, hr)["loss"]           # and training resumes on the same trajectory
    l2 = m2.train_on_batch(hr * 0.9, hr)["loss"]
    assert abs(l1 - l2) <= 1e-6 * max(1.0, abs(l1))
    # bare weights file
    m1.save_weights(tmp_path / "w.weights.h5")
    clear_session()
    m3, _ = B.build_super_resolution_unet(0.5, depth_override=1, input_size=16)
    m3.load_weights(tmp_path / "w.weights.h5")
    assert all(np.array_equal(a, b) for a, b in zip(m1.get_weights(), m3.get_weights()))
    # another depth: refused, naming the first layer that does not fit
    clear_session()
    m4, _ = B.build_super_resolution_unet(0.5, depth_override=2, input_size=16)
    with pytest.raises(B200Error, match="conv2d|expects"):
        m4.load_weights(path)
    (tmp_path / "junk.keras").write_bytes(b"not a checkpoint at all, just bytes" * 40)
    with pytest.raises(B200Error, match="neither"):
        m3.load_weights(tmp_path / "junk.keras")
    # round-1 archives (npz member) still load, and their stored weight names are checked
    buf = io.BytesIO()
    np.savez(buf, **{f"w{i:04d}": a for i, a in enumerate(m1.get_weights())})
    with zipfile.ZipFile(tmp_path / "old.keras", "w") as z:
        z.writestr("config.json", json.dumps({"weight_names": [x["name"] for x in m1.weights]}))
        z.writestr("model.weights.npz", buf.getvalue())
    m3.load_weights(tmp_path / "old.keras")
    with pytest.raises(B200Error, match="another model"):
        m4.load_weights(tmp_path / "old.keras")


def test_sr_vanilla_bn_unet_step():
    """The reference's fixed-depth baseline (Super_resolution/code/u-net-vinillia.py:128-167): BatchNorm blocks, MaxPool,
    bilinear UpSampling2D + Conv3x3/ReLU, concat[up, skip], 3-channel sigmoid head, trained on the MSE term."""
    from b200unet import builders as B
    from b200unet.keras import losses as LS
    from b200unet.keras.optimizers import Adam
    from oracle import keras_ops as K, models as M
    _setup("float32")
    depth, base, P, batch = 2, 64, 32, 4
    model = B.build_vanilla_super_resolution_unet((P, P, 3), base, depth)
    spec = M.sr_vanilla_spec(depth, base)
    assert model.count_params() == M.param_count(spec)
    ws_np = M.init_weights(spec, seed=11, jitter=0.05)
    model.set_weights(ws_np)
    model.compile(optimizer=Adam(1e-3), loss=LS.SRLoss("mse"), metrics=[LS.PSNRMetric()])
    rng = np.random.default_rng(5)
    hr = rng.random((batch, P, P, 3), dtype=np.float32)
    lr = np.clip(hr + 0.05 * rng.standard_normal(hr.shape).astype(np.float32), 0, 1)
    fwd = lambda ws, xx: M.sr_vanilla_forward(ws, xx, depth, True)
    y_ref, l_ref, g_ref = _oracle_step(ws_np, torch.from_numpy(lr), torch.from_numpy(hr), fwd, K.mse_loss, None,
                                       dtype=torch.float64)
    logs = model.train_on_batch(lr, hr)
    print(f"sr vanilla loss {logs['loss']:.7f} vs {l_ref:.7f}")
    assert abs(logs["loss"] - l_ref) < 1e-5
    i, bad = 0, []
    for ly in model.layers:
        for w in ly.weight_specs:
            nm = w["name"].split("/", 1)[1]
            if w["trainable"]:
                e = relerr(model._grad(ly, nm), g_ref[i])
                pre_bn_bias = nm == "bias" and type(ly).__name__ == "Conv2D" and ly.activation is None
                # same bar as the BatchNorm segmentation net: BN backward is ill-conditioned in fp32
                if not (e < 2e-2 or pre_bn_bias or g_ref[i].abs().max() < 1e-7):
                    bad.append((w["name"], e))
            i += 1
    assert not bad, bad
    _setup("mixed_bfloat16")
    model = B.build_vanilla_super_resolution_unet((P, P, 3), base, depth)
    model.compile(optimizer=Adam(1e-3), loss=LS.SRLoss("mse"), metrics=[LS.PSNRMetric()])
    ls = [model.train_on_batch(lr, hr)["loss"] for _ in range(12)]
    assert np.isfinite(ls).all() and ls[-1] < ls[0]
    _setup("float32")


def test_seg_vanilla_base32_bf16_on_tensor_cores():
    """The reference's default width (base_channels=32, BASELINE config 4) under the bf16 policy: the 32-channel level
    (stem 3->32, 32->32, 64->32, Conv2DTranspose 64->32) runs on the tcgen05 kernels through zero-filled 64-channel
    tiles.  Loss and weight gradients against the oracle evaluated with bf16 storage rounding."""
    from b200unet import builders as B
    from b200unet.keras import losses as LS
    from b200unet.keras.engine import Plan
    from b200unet.keras.optimizers import Adam
    from oracle import keras_ops as K, models as M
    _setup("mixed_bfloat16")
    depth, base, P, batch, classes = 2, 32, 32, 4, 5
    model = B.build_unet(P, num_classes=classes, base_channels=base, depth=depth)
    ws_np = M.init_weights(M.seg_vanilla_spec(depth, base, classes), seed=11, jitter=0.05)
    model.set_weights(ws_np)
    model.compile(optimizer=Adam(1e-3), loss=LS.CategoricalCrossentropy())
    rng = np.random.default_rng(5)
    x = rng.random((batch, P, P, 3), dtype=np.float32)
    t = rng.integers(0, classes, (batch, P, P)).astype(np.int32)
    tt = torch.nn.functional.one_hot(torch.from_numpy(t).long(), classes).float()
    fwd = lambda ws, xx: M.seg_vanilla_forward(ws, M.bf16_round(xx), depth, classes, M.bf16_round)
    y_ref, l_ref, g_ref = _oracle_step(ws_np, torch.from_numpy(x), tt, fwd, K.categorical_crossentropy, M.bf16_round)
    logs = model.train_on_batch(x, t)
    plan = model._train_state(batch)["plan"]
    convs = [op for op in plan.ops if op.kind == "conv" and op.layer.kernel_size == (3, 3)]
    assert convs and all(Plan.is_tc(op) for op in convs)          # no 3x3 convolution is left on the SIMT kernels
    print(f"seg vanilla base32 bf16: loss {logs['loss']:.5f} vs {l_ref:.5f}")
    assert abs(logs["loss"] - l_ref) < 2e-2 * max(1, abs(l_ref))
    i, worst = 0, 0.0
    for ly in model.layers:
        for w in ly.weight_specs:
            e = relerr(model._grad(ly, w["name"].split("/", 1)[1]), g_ref[i])
            if g_ref[i].abs().max() >= 1e-7:
                worst = max(worst, e)
                assert e < 5e-2, (w["name"], e)
            i += 1
    print(f"   worst weight-gradient relerr {worst:.3e}")
    _setup("float32")


def test_wgrad_side_stream_matches_single_stream():
    """Filter-gradient kernels forked onto a second stream inside the captured step (Model._run_bwd), and the experimental
    per-layer Adam behind each of them (B200_OVERLAP_ADAM), against the single-stream order -- in LOCKSTEP over four steps:
    before every step the reference run's state (fp32 master, bf16 shadow, Adam moments and step counter) is copied into
    the other runs, then every run replays its own captured graph.

    Forward and activation-gradient kernels use no atomics, so from identical state every activation, every LayerNorm
    statistic and every activation gradient must be BIT-identical whatever the stream schedule -- any race between a
    side-stream wgrad and the main stream would show up as a differing bit.  Parameter gradients are summed with fp32
    atomics (summation order differs from run to run): compared per tensor at 1e-5.

    Why lockstep and not free-running trajectories (round-1 failure, root-caused with tools/trajectory_probe.py,
    profiles/r02_trajectory_probe_*.json): the atomic-order noise leaves ~13 000 of the 8.6 M fp32 master weights one ulp
    apart after a step, which now and then rounds ONE OR TWO bf16 shadow weights the other way; in this fast-moving phase of
    training (loss 0.36 -> 0.17 in four steps) a single such flip moves the step-2 loss by 2e-5 .. 4e-4.  The same spread
    shows up between two runs of the SAME schedule, in isolation and in-suite; it is not a property of the second stream."""
    from b200unet import builders as B
    from b200unet.keras.optimizers import Adam
    rng = np.random.default_rng(2)
    hr = rng.random((8, 64, 64, 3), dtype=np.float32)
    lr = np.clip(hr + 0.05 * rng.standard_normal(hr.shape).astype(np.float32), 0, 1)
    models = []
    for overlap, adam in ((False, False), (True, False), (True, True)):
        _setup("mixed_bfloat16")
        model, _ = B.build_super_resolution_unet(0.5, depth_override=3, input_size=64)
        model.overlap_wgrad, model.overlap_adam = overlap, adam
        head = model.get_layer("residual_rgb")
        head.weight_specs[0]["value"] = np.random.default_rng(3).uniform(-0.2, 0.2, (1, 1, 64, 3)).astype(np.float32)
        loss, metrics = B.build_losses_and_metrics("charbonnier")
        model.compile(optimizer=Adam(learning_rate=1e-4), loss=loss, metrics=metrics)
        models.append(model)

    def bits(t):
        return t.view(torch.int16 if t.dtype == torch.bfloat16 else torch.int32)

    ref = models[0]
    losses = [[], [], []]
    for step in range(4):
        if step:                                   # lockstep: everyone enters the step from the reference state
            rs = ref.optimizer._state
            for m in models[1:]:
                ms = m.optimizer._state
                m.P.copy_(ref.P); m.S.copy_(ref.S)
                ms["m"].copy_(rs["m"]); ms["v"].copy_(rs["v"]); ms["step"].copy_(rs["step"])
        snaps = []
        for k, m in enumerate(models):
            losses[k].append(m.train_on_batch(lr, hr)["loss"])
            torch.cuda.synchronize()
            plan = m._train_state(8)["plan"]
            roots = [v for v in plan.all_vals if v.parent is None]
            snaps.append(([(v.name, v.buf.clone()) for v in roots],
                          [(v.name, v.grad.clone()) for v in roots if v.grad is not None],
                          [(op.output.name, op.mean.clone(), op.rstd.clone()) for op in plan.ops if op.kind == "ln"],
                          m.G.clone(), m.optimizer._state["m"].clone()))
        acts0, grads0, stats0, g0, m0 = snaps[0]
        for k in (1, 2):
            acts, grads, stats, g, mm = snaps[k]
            for (n0, a), (_, b) in zip(acts0, acts):
                assert torch.equal(bits(a), bits(b)), f"step {step}, run {k}: activation {n0} differs"
            for (n0, mu0, r0), (_, mu, r) in zip(stats0, stats):
                assert torch.equal(bits(mu0), bits(mu)) and torch.equal(bits(r0), bits(r)), f"step {step}: LN stats {n0}"
            for (n0, a), (_, b) in zip(grads0, grads):
                assert torch.equal(bits(a), bits(b)), f"step {step}, run {k}: activation gradient {n0} differs"
            assert abs(losses[0][-1] - losses[k][-1]) <= 2e-6       # the loss sum itself is an atomic reduction
            for ly in ref.layers:                                     # parameter gradients: atomic summation order only
                for w in ly.weight_specs:
                    r = ref._grad_range(ly, w["name"].split("/", 1)[1]) if w["trainable"] else None
                    if r is not None and g0[r[0]:r[0] + r[1]].abs().max() > 0:
                        assert relerr(g[r[0]:r[0] + r[1]], g0[r[0]:r[0] + r[1]]) < 1e-5, (step, k, w["name"])
            assert relerr(mm, m0) < 1e-5                              # the optimizer saw the COMPLETE gradients
    print("lockstep losses:", losses)
    assert losses[0][-1] < 0.6 * losses[0][0]                         # and it trains
    _setup("float32")


def test_sync_batchnorm_two_ranks_one_gpu():
    """Synchronised BatchNorm (SURVEY 8e caveat): two data-parallel ranks (both on cuda:0, gloo) x 4 samples against one
    process x 8 samples -- loss, every weight gradient, the moving statistics and the weights after Adam
    (tools/syncbn_check.py; the driver's multi-GPU runs use NCCL on separate devices)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B200_DP_CHECK_ONE_GPU="1", B200_DP_SHARD="0", MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", os.path.join(root, "tools", "syncbn_check.py")],
                       env=env, capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "SYNCBN_OK" in r.stdout


def test_mixed_float16_policy_dynamic_loss_scaling():
    """The reference's policy (`mixed_float16`, train_adaptive_unet.py:471-477): keras wraps Adam in a LossScaleOptimizer.
    Here the kernels stay bf16 and the wrapper's semantics are reproduced on the device: the loss gradient is scaled by
    2**15, Adam un-scales, so the trajectory equals the mixed_bfloat16 one; a batch that produces nan gradients is
    skipped (weights, moments and Adam's step counter untouched) and halves the scale."""
    from b200unet import builders as B
    from b200unet.keras.optimizers import Adam
    rng = np.random.default_rng(5)
    hr = rng.random((4, 32, 32, 3), dtype=np.float32)
    lr = np.clip(hr + 0.05 * rng.standard_normal(hr.shape).astype(np.float32), 0, 1)
    out = {}
    for policy in ("mixed_bfloat16", "mixed_float16"):
        _setup(policy)
        model, _ = B.build_super_resolution_unet(0.5, depth_override=2, input_size=32)
        head = model.get_layer("residual_rgb")
        head.weight_specs[0]["value"] = np.random.default_rng(3).uniform(-0.2, 0.2, (1, 1, 64, 3)).astype(np.float32)
        loss, metrics = B.build_losses_and_metrics("charbonnier")
        model.compile(optimizer=Adam(learning_rate=1e-4), loss=loss, metrics=metrics)
        l0 = model.train_on_batch(lr, hr)["loss"]
        out[policy] = (l0, model.G.clone(), model.P.clone(), model)
    (lb, gb, pb, _), (lf, gf, pf, model) = out["mixed_bfloat16"], out["mixed_float16"]
    opt = model.optimizer
    assert opt.dynamic_loss_scale and opt.loss_scale == 2.0 ** 15
    assert abs(lb - lf) <= 1e-6
    assert relerr(gf / 2.0 ** 15, gb) < 2e-2          # scaled gradients (bf16 rounding of the scaled activations' gradients)
    assert relerr(pf, pb) < 1e-5                       # Adam un-scales: same update
    st = opt._state
    assert st["loss_scale"].tolist()[:3] == [2.0 ** 15, 1.0, 0.0]
    p_before, m_before, step_before = model.P.clone(), st["m"].clone(), int(st["step"])
    bad = hr.copy(); bad[0] = np.nan                   # nan targets for one image: nan loss gradient wherever the output is not
    # clipped (ClippedResidualAdd passes no gradient where input + residual leaves [0, 1]), hence nan weight gradients
    model.train_on_batch(lr, bad)
    torch.cuda.synchronize()
    assert torch.equal(model.P, p_before) and torch.equal(st["m"], m_before) and int(st["step"]) == step_before
    assert st["loss_scale"].tolist() == [2.0 ** 14, 0.0, 0.0, 1.0]
    l2 = model.train_on_batch(lr, hr)["loss"]          # and training goes on
    assert np.isfinite(l2) and not torch.equal(model.P, p_before)
    _setup("float32")


@pytest.mark.parametrize("cfg", [("C2", 4, 0.25, 128, 8), ("C3", 5, 0.25, 128, 4)])
def test_sr_unet_step_at_baseline_topologies(cfg):
    """Whole-model parity at the BASELINE topologies themselves -- C2: depth 4, scale 0.25, 128x128 (128 -> 32 -> 8 -> 2
    -> 1); C3: depth 5 (... -> 1 -> 1, 138 M parameters) -- bf16 policy, at a batch the CPU oracle finishes in seconds.
    These are the graphs bench.py times: CTA-pair convolutions at 128^2 and 32^2, pair tiles at 8^2, the split-K kernels at
    2^2 / 1^2, the same-size resize at the 1x1 levels.  Output, loss, PSNR and EVERY weight gradient against the oracle
    with bf16 storage rounding (tolerances of SURVEY 8c: 3e-2 on weight gradients, 1e-3 relative on the loss)."""
    from b200unet import builders as B
    from b200unet.keras.optimizers import Adam
    from oracle import keras_ops as K, models as M
    name, depth, scale, P, batch = cfg
    _setup("mixed_bfloat16")
    model, _ = B.build_super_resolution_unet(scale, depth_override=depth, input_size=P)
    ws_np = M.init_weights(M.sr_unet_spec(depth), seed=2024, randomize_zero_kernels=True, jitter=0.05)
    model.set_weights(ws_np)
    loss, metrics = B.build_losses_and_metrics("charbonnier")
    model.compile(optimizer=Adam(learning_rate=1e-4), loss=loss, metrics=metrics)
    rng = np.random.default_rng(77)
    hr = rng.random((batch, P, P, 3), dtype=np.float32)
    lr = np.clip(hr + 0.05 * rng.standard_normal(hr.shape).astype(np.float32), 0, 1)
    fwd = lambda ws, x: M.sr_unet_forward(ws, x, scale, depth, rnd=M.bf16_round)
    y_ref, l_ref, g_ref = _oracle_step(ws_np, torch.from_numpy(lr), torch.from_numpy(hr), fwd, K.charbonnier_loss, M.bf16_round)
    # the policy's own band: the same oracle, same bf16 storage points, evaluated in float64.  The levels below 8x8 hold
    # batch * 4 (2x2) or batch (1x1) pixels: a weight gradient there is a sum over a handful of bf16-rounded terms, and a
    # single activation that rounds the other way moves it by percents -- in the oracle as much as in the kernels.
    _, _, g_alt = _oracle_step(ws_np, torch.from_numpy(lr), torch.from_numpy(hr), fwd, K.charbonnier_loss, M.bf16_round,
                               dtype=torch.float64)
    y = model(lr)
    logs = model.train_on_batch(lr, hr)
    torch.cuda.synchronize()
    e_out = relerr(y, y_ref)
    print(f"[{name}] out relerr {e_out:.3e} loss {logs['loss']:.6f} vs {l_ref:.6f} psnr {logs['psnr']:.3f}")
    assert e_out < 3e-2
    assert abs(logs["loss"] - l_ref) < 1e-3 * max(1.0, abs(l_ref))
    assert abs(logs["psnr"] - K.psnr_metric(torch.from_numpy(hr), y_ref).item()) < 0.05
    i, bad, worst, worst_band = 0, [], 0.0, 0.0
    for ly in model.layers:
        for w in ly.weight_specs:
            e = relerr(model._grad(ly, w["name"].split("/", 1)[1]), g_ref[i])
            if g_ref[i].abs().max().item() >= 1e-7:
                band = relerr(g_alt[i], g_ref[i])
                worst, worst_band = max(worst, e), max(worst_band, band)
                if e >= 3e-2:
                    print(f"   {w['name']:<36s} relerr {e:.3e}   oracle fp64-vs-fp32 band {band:.3e}")
                # 3e-2; where the oracle itself moves by more than 1e-2 between two evaluation precisions
                # (the few-pixel deep levels), within three times that band
                if e >= 3e-2 and not (band > 1e-2 and e <= 3.0 * band):
                    bad.append((w["name"], e, band))
            i += 1
    print(f"[{name}] worst oracle band {worst_band:.3e}")
    print(f"[{name}] worst weight-gradient relerr {worst:.3e} over {i} tensors")
    assert not bad, bad
    _setup("float32")


def test_deterministic_filter_gradients():
    """Model.deterministic = True: the filter gradients (the slab + fixed-order reduce path) are BIT-identical between two
    runs from the same state and between the single-stream and the two-stream schedule; they agree with the default
    (atomic) path to summation-order noise."""
    from b200unet import builders as B
    from b200unet.keras.optimizers import Adam
    rng = np.random.default_rng(2)
    hr = rng.random((8, 64, 64, 3), dtype=np.float32)
    lr = np.clip(hr + 0.05 * rng.standard_normal(hr.shape).astype(np.float32), 0, 1)
    grads = []
    for det, overlap in ((True, False), (True, True), (True, True), (False, True)):
        _setup("mixed_bfloat16")
        model, _ = B.build_super_resolution_unet(0.5, depth_override=3, input_size=64)
        model.deterministic, model.overlap_wgrad = det, overlap
        head = model.get_layer("residual_rgb")
        head.weight_specs[0]["value"] = np.random.default_rng(3).uniform(-0.2, 0.2, (1, 1, 64, 3)).astype(np.float32)
        loss, metrics = B.build_losses_and_metrics("charbonnier")
        model.compile(optimizer=Adam(learning_rate=1e-4), loss=loss, metrics=metrics)
        model.train_on_batch(lr, hr)
        torch.cuda.synchronize()
        kernels = torch.cat([model._grad(ly, "kernel").flatten() for ly in model.layers
                             if type(ly).__name__ == "Conv2D" and ly.kernel_size == (3, 3) and ly.name != model.layers[1].name])
        grads.append(kernels.clone())
    assert torch.equal(grads[0].view(torch.int32), grads[1].view(torch.int32))     # single stream == two streams, bitwise
    assert torch.equal(grads[1].view(torch.int32), grads[2].view(torch.int32))     # run to run, bitwise
    assert relerr(grads[3], grads[0]) < 1e-5                                       # atomics: order noise only
    _setup("float32")
