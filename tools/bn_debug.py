"""Per-layer comparison of the BatchNorm segmentation net's backward pass (bf16 policy) with the oracle: where do the
weight gradients part ways -- dy (gradient of the BN+ReLU output), dz (gradient of the conv output) or the wgrad?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200unet import builders as B, ops
from b200unet.keras import clear_session, losses as LS, mixed_precision
from b200unet.keras.optimizers import Adam
from oracle import keras_ops as K, models as M

def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))

clear_session(); mixed_precision.set_global_policy("mixed_bfloat16")
depth, base, P, batch = 2, 64, 32, 4
model = B.build_adaptive_depth_unet(P, base, depth)
ws_np = M.init_weights(M.seg_adaptive_spec(depth, base), seed=7, jitter=0.05)
model.set_weights(ws_np)
model.compile(optimizer=Adam(1e-3), loss=LS.make_hybrid_ce_dice_loss(0.4, 0.6), metrics=[LS.dice_metric, LS.iou_metric])
rng = np.random.default_rng(3)
x = rng.random((batch, P, P, 3), dtype=np.float32)
t = (rng.random((batch, P, P, 1)) > 0.5).astype(np.float32)
model.train_on_batch(x, t)
torch.cuda.synchronize()
plan = model._train_state(batch)["plan"]

# oracle with captured intermediates
zs, ys = [], []
orig = M._conv_bn_relu
def capture(xx, w, rnd, training, new_stats, momentum=0.99):
    k, b = w.take(2); g, be, mm, mv = w.take(4)
    z = rnd(K.conv2d_same(xx, k, b)); z.retain_grad(); zs.append(z)
    y, nm, nv = K.batch_norm_train(z, g, be, mm, mv, momentum)
    y = rnd(K.relu(y)); y.retain_grad(); ys.append(y)
    return y
M._conv_bn_relu = capture
ws = [torch.tensor(w, requires_grad=True) for w in ws_np]
wsr = [M.bf16_round(w) if w.dim() == 4 else w for w in ws]
yo = M.seg_adaptive_forward(wsr, M.bf16_round(torch.from_numpy(x)), depth, True, M.bf16_storage, [])
K.bce_dice_loss(torch.from_numpy(t), yo, 0.4, 0.6).backward()
bn_ops = [op for op in plan.ops if op.kind == "bn"]
convs = [op for op in plan.ops if op.kind == "conv"]
gi = 0
for k, op in enumerate(bn_ops):
    z, y = op.inputs[0], op.output
    cv = convs[k]
    gw = model._grad(cv.layer, "kernel")
    # our wgrad recomputed in fp32 torch from OUR x and OUR dz: isolates the wgrad kernel
    xin = cv.inputs[0].buf.float()
    dz = z.grad.float()
    xr = xin.detach().cpu().requires_grad_(False)
    wref = torch.zeros_like(gw.cpu())
    xp = torch.nn.functional.pad(xr, (0, 0, 1, 1, 1, 1))
    dzc = dz.cpu()
    for kh in range(3):
        for kw in range(3):
            wref[kh, kw] = torch.einsum("nhwc,nhwo->co", xp[:, kh:kh + z.h, kw:kw + z.w, :].double(), dzc.double()).float()
    print(f"layer {k:2d} {cv.layer.name:10s} HxW {z.h:2d} C {z.c:3d}: z {rel(z.buf, zs[k]):.2e} y {rel(y.buf, ys[k]):.2e} | "
          f"dy {rel(y.grad, ys[k].grad):.2e} dz {rel(z.grad, zs[k].grad):.2e} | wgrad kernel vs einsum(our x, our dz) {rel(gw, wref):.2e} "
          f"| dW vs oracle {rel(gw, ws[6 * k].grad):.2e} | sum(dz)/sum|dz| ours {float(dz.sum() / dz.abs().sum()):+.2e} oracle "
          f"{float(zs[k].grad.sum() / zs[k].grad.abs().sum()):+.2e}", flush=True)
