"""torch-CPU restatement of the reference's three U-Net builders.

Oracle code (test infrastructure, see ``oracle/__init__.py``; parity unpinned at
op level, pinned at shape/parameter-count level by the reference's
``model_summary`` dumps).  Each model is a pure function of a flat list of weight
tensors in Keras layer-creation order (Conv2D: kernel HWIO, bias;
LayerNormalization: gamma, beta; BatchNormalization: gamma, beta, moving_mean,
moving_variance; Conv2DTranspose: kernel [kh,kw,Cout,Cin], bias), the same order
the product's Keras-shaped ``Model.get_weights()`` uses.

Restated builders (paths relative to /root/reference):
  * ``build_super_resolution_unet`` -- Super_resolution/code/train_adaptive_unet.py:217-287
    with ``conv_block`` :200-210
  * ``build_adaptive_depth_unet``   -- Segmenation/code/train_adaptive_unet.py:335-362
    with ``conv_block`` :325-332
  * ``build_unet``                  -- Segmenation/code/unet_vinillia.py:72-91
    with ``conv_block`` :42-52, ``encoder_block`` :60-63, ``decoder_block`` :66-69
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import keras_ops as K
from . import resize_np

Spec = List[Tuple[str, Tuple[int, ...], str]]  # (name, shape, init kind)


# --------------------------------------------------------------------------- #
# Depth rules (shared/custom_layers.py:10-82)
# --------------------------------------------------------------------------- #
def custom_depth_from_scale(scale, min_depth=1, max_depth=7, base_resolution=256, min_feature=21):
    """shared/custom_layers.py:42-75: shrink with ceil(extent*scale) (Python
    floats here, exactly as the reference does on the host) until the next
    extent would fall under ``min_feature`` or ``max_depth`` is reached."""
    if not (0.05 < scale < 1.0):
        raise ValueError("Scale should be between 0 and 1 (exclusive).")
    depth = max(min_depth, 1)
    extent = base_resolution
    while depth < max_depth:
        cand = math.ceil(extent * scale)
        if cand < min_feature:
            break
        extent = cand
        depth += 1
    return max(min_depth, min(depth, max_depth))


def infer_depth_from_scale(scale, min_depth=1, max_depth=4):
    """shared/custom_layers.py:10-28."""
    if not (0.05 < scale < 1.0):
        raise ValueError("Scale should be between 0 and 1 (exclusive).")
    depth = 1 if scale <= 0.25 else (2 if scale <= 0.45 else 3)
    return max(min_depth, min(depth, max_depth))


def estimate_bottleneck_size(hr, scale, depth):
    """shared/custom_layers.py:77-82 (uses round, unlike the layer's ceil)."""
    size = hr
    for _ in range(depth):
        size = max(1, int(round(size * scale)))
    return size


# --------------------------------------------------------------------------- #
# Weight specs and initialisation
# --------------------------------------------------------------------------- #
class _SpecBuilder:
    def __init__(self):
        self.spec: Spec = []
        self.counts: Dict[str, int] = {}

    def _name(self, base):
        k = self.counts.get(base, 0)
        self.counts[base] = k + 1
        return base if k == 0 else f"{base}_{k}"

    def conv(self, cin, cout, k, name=None, zeros=False):
        nm = name or self._name("conv2d")
        self.spec.append((nm + "/kernel", (k, k, cin, cout), "zeros" if zeros else "glorot"))
        self.spec.append((nm + "/bias", (cout,), "zeros"))

    def convT(self, cin, cout, k):
        nm = self._name("conv2d_transpose")
        self.spec.append((nm + "/kernel", (k, k, cout, cin), "glorotT"))
        self.spec.append((nm + "/bias", (cout,), "zeros"))

    def ln(self, c):
        nm = self._name("layer_normalization")
        self.spec.append((nm + "/gamma", (c,), "ones"))
        self.spec.append((nm + "/beta", (c,), "zeros"))

    def bn(self, c):
        nm = self._name("batch_normalization")
        self.spec.append((nm + "/gamma", (c,), "ones"))
        self.spec.append((nm + "/beta", (c,), "zeros"))
        self.spec.append((nm + "/moving_mean", (c,), "zeros"))
        self.spec.append((nm + "/moving_variance", (c,), "ones"))


def sr_unet_spec(depth, base_channels=64, residual_head_channels=64) -> Spec:
    b = _SpecBuilder()
    nf, cin = base_channels, 3
    for _ in range(depth):
        b.conv(cin, nf, 3); b.ln(nf); b.conv(nf, nf, 3); b.ln(nf)
        cin, nf = nf, nf * 2
    b.conv(cin, nf, 3); b.ln(nf); b.conv(nf, nf, 3); b.ln(nf)
    cin = nf
    for _ in range(depth):
        nf //= 2
        b.conv(cin, nf, 3)  # post-upsample conv + ReLU
        b.conv(2 * nf, nf, 3); b.ln(nf); b.conv(nf, nf, 3); b.ln(nf)
        cin = nf
    h = residual_head_channels
    b.conv(cin, h, 3); b.ln(h); b.conv(h, h, 3); b.ln(h)
    b.conv(h, 3, 1, name="residual_rgb", zeros=True)
    return b.spec


def seg_adaptive_spec(depth, base_channels) -> Spec:
    b = _SpecBuilder()
    f, cin, chans = base_channels, 3, []
    for _ in range(depth):
        b.conv(cin, f, 3); b.bn(f); b.conv(f, f, 3); b.bn(f)
        chans.append(f)
        cin, f = f, f * 2
    b.conv(cin, f, 3); b.bn(f); b.conv(f, f, 3); b.bn(f)
    cin = f
    for f in reversed(chans):
        b.conv(cin + f, f, 3); b.bn(f); b.conv(f, f, 3); b.bn(f)
        cin = f
    b.conv(cin, 1, 1, name="lesion_mask")
    return b.spec


def sr_vanilla_spec(depth=4, base_channels=64) -> Spec:
    """``build_super_resolution_unet(input_shape)`` of Super_resolution/code/u-net-vinillia.py:128-167."""
    b = _SpecBuilder()
    nf, cin, chans = base_channels, 3, []
    for _ in range(depth):
        b.conv(cin, nf, 3); b.bn(nf); b.conv(nf, nf, 3); b.bn(nf)
        chans.append(nf)
        cin, nf = nf, nf * 2
    b.conv(cin, nf, 3); b.bn(nf); b.conv(nf, nf, 3); b.bn(nf)
    cin = nf
    for nf in reversed(chans):
        b.conv(cin, nf, 3)                       # post-upsample conv + ReLU (:147)
        b.conv(2 * nf, nf, 3); b.bn(nf); b.conv(nf, nf, 3); b.bn(nf)
        cin = nf
    b.conv(cin, 3, 1, name="enhanced_rgb")
    return b.spec


def seg_vanilla_spec(depth, base_channels=32, num_classes=1) -> Spec:
    b = _SpecBuilder()
    nf, cin = base_channels, 3
    for _ in range(depth):
        b.conv(cin, nf, 3); b.ln(nf); b.conv(nf, nf, 3); b.ln(nf)
        cin, nf = nf, nf * 2
    b.conv(cin, nf, 3); b.ln(nf); b.conv(nf, nf, 3); b.ln(nf)
    cin = nf
    for _ in range(depth):
        nf //= 2
        b.convT(cin, nf, 2)
        b.conv(2 * nf, nf, 3); b.ln(nf); b.conv(nf, nf, 3); b.ln(nf)
        cin = nf
    b.conv(cin, num_classes, 1, name="mask_logits")
    return b.spec


def param_count(spec: Spec, trainable_only=False) -> int:
    tot = 0
    for name, shape, _ in spec:
        if trainable_only and ("moving_" in name):
            continue
        tot += int(np.prod(shape))
    return tot


def init_weights(spec: Spec, seed=1234, randomize_zero_kernels=True, jitter=0.0) -> List[np.ndarray]:
    """Glorot-uniform kernels (keras default), zero biases, unit gammas.

    ``randomize_zero_kernels`` replaces the zero-initialised head kernel
    (train_adaptive_unet.py:267-274) with a Glorot draw: with the zero head every
    upstream gradient is exactly 0 and parity would be vacuous.
    ``jitter`` > 0 perturbs biases/gammas/betas so their gradients paths are
    exercised with non-trivial values.
    """
    rng = np.random.default_rng(seed)
    out = []
    for name, shape, kind in spec:
        if kind in ("glorot", "glorotT") or (kind == "zeros" and len(shape) == 4 and randomize_zero_kernels):
            kh, kw, a, b_ = shape
            fan_in, fan_out = (kh * kw * a, kh * kw * b_) if kind != "glorotT" else (kh * kw * b_, kh * kw * a)
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            arr = rng.uniform(-lim, lim, size=shape)
        elif kind == "ones":
            arr = np.ones(shape) + jitter * rng.standard_normal(shape)
            if "moving_variance" in name:
                arr = np.abs(arr)
        else:
            arr = np.zeros(shape) + jitter * rng.standard_normal(shape)
        out.append(arr.astype(np.float32))
    return out


# --------------------------------------------------------------------------- #
# Forward functions
# --------------------------------------------------------------------------- #
class _W:
    """Sequential reader over the flat weight list."""

    def __init__(self, ws: Sequence[torch.Tensor]):
        self.ws, self.i = list(ws), 0

    def take(self, n):
        out = self.ws[self.i:self.i + n]
        self.i += n
        return out

    def done(self):
        assert self.i == len(self.ws), (self.i, len(self.ws))


def _ident(x):
    return x


def _conv_ln_relu(x, w, rnd, cap, tag):
    k, b = w.take(2)
    g, be = w.take(2)
    z = rnd(K.conv2d_same(x, k, b))          # conv output is stored (policy dtype)
    y = rnd(K.relu(K.layer_norm(z, g, be)))  # LN in fp32 on the stored z, stored again
    if cap is not None:
        cap[tag + ".z"] = z
        cap[tag + ".y"] = y
    return y


def sr_unet_forward(ws, x, scale, depth, rnd: Callable = _ident, cap: Optional[dict] = None):
    """Forward of ``build_super_resolution_unet`` (train_adaptive_unet.py:217-287).

    ``rnd`` emulates the storage precision of the mixed-precision policy (identity
    for fp32; round-to-bf16-and-back for the bf16 policy): it is applied to every
    tensor a Keras layer would store in the compute dtype.  ``cap`` (dict) collects
    named intermediates for per-layer parity.
    """
    w = _W(ws)
    inp = x
    skips = []
    li = 0
    for d in range(depth):
        x = _conv_ln_relu(x, w, rnd, cap, f"enc{d}.c1")
        x = _conv_ln_relu(x, w, rnd, cap, f"enc{d}.c2")
        skips.append(x)
        x = rnd(K.resize_by_scale(x, scale))
        if cap is not None:
            cap[f"enc{d}.down"] = x
    x = _conv_ln_relu(x, w, rnd, cap, "mid.c1")
    x = _conv_ln_relu(x, w, rnd, cap, "mid.c2")
    for d in reversed(range(depth)):
        skip = skips[d]
        x = rnd(K.resize_to_match(x, skip))
        k, b = w.take(2)
        x = rnd(K.relu(K.conv2d_same(x, k, b)))
        if cap is not None:
            cap[f"dec{d}.up"] = x
        x = torch.cat([x, skip], dim=-1)
        x = _conv_ln_relu(x, w, rnd, cap, f"dec{d}.c1")
        x = _conv_ln_relu(x, w, rnd, cap, f"dec{d}.c2")
    x = _conv_ln_relu(x, w, rnd, cap, "head.c1")
    x = _conv_ln_relu(x, w, rnd, cap, "head.c2")
    k, b = w.take(2)
    res = rnd(K.conv2d_same(x, k, b))
    if cap is not None:
        cap["residual_rgb"] = res
    w.done()
    return rnd(K.clipped_residual_add(inp, res))


def _conv_bn_relu(x, w, rnd, training, new_stats, momentum=0.99):
    k, b = w.take(2)
    g, be, mm, mv = w.take(4)
    z = rnd(K.conv2d_same(x, k, b))
    if training:
        y, nm, nv = K.batch_norm_train(z, g, be, mm, mv, momentum)
        new_stats.extend([nm, nv])
    else:
        y = K.batch_norm_infer(z, g, be, mm, mv)
    return rnd(K.relu(y))


def seg_adaptive_forward(ws, x, depth, training=True, rnd: Callable = _ident, new_stats=None):
    """``build_adaptive_depth_unet`` (Segmenation/code/train_adaptive_unet.py:335-362)."""
    w = _W(ws)
    new_stats = [] if new_stats is None else new_stats
    skips = []
    for _ in range(depth):
        x = _conv_bn_relu(x, w, rnd, training, new_stats)
        x = _conv_bn_relu(x, w, rnd, training, new_stats)
        skips.append(x)
        x = K.max_pool2(x)
    x = _conv_bn_relu(x, w, rnd, training, new_stats)
    x = _conv_bn_relu(x, w, rnd, training, new_stats)
    for skip in reversed(skips):
        x = rnd(K.upsample2_bilinear(x))
        x = torch.cat([x, skip], dim=-1)
        x = _conv_bn_relu(x, w, rnd, training, new_stats)
        x = _conv_bn_relu(x, w, rnd, training, new_stats)
    k, b = w.take(2)
    w.done()
    return torch.sigmoid(K.conv2d_same(x, k, b))


def sr_vanilla_forward(ws, x, depth=4, training=True, rnd: Callable = _ident, new_stats=None):
    """BatchNorm / MaxPool / bilinear-UpSampling SR U-Net with a 3-channel sigmoid head
    (Super_resolution/code/u-net-vinillia.py:128-167): decoder = UpSampling2D -> Conv3x3+ReLU -> concat[x, skip] -> conv_block."""
    w = _W(ws)
    new_stats = [] if new_stats is None else new_stats
    skips = []
    for _ in range(depth):
        x = _conv_bn_relu(x, w, rnd, training, new_stats)
        x = _conv_bn_relu(x, w, rnd, training, new_stats)
        skips.append(x)
        x = K.max_pool2(x)
    x = _conv_bn_relu(x, w, rnd, training, new_stats)
    x = _conv_bn_relu(x, w, rnd, training, new_stats)
    for skip in reversed(skips):
        x = rnd(K.upsample2_bilinear(x))
        k, b = w.take(2)
        x = rnd(K.relu(K.conv2d_same(x, k, b)))
        x = torch.cat([x, skip], dim=-1)
        x = _conv_bn_relu(x, w, rnd, training, new_stats)
        x = _conv_bn_relu(x, w, rnd, training, new_stats)
    k, b = w.take(2)
    w.done()
    return torch.sigmoid(K.conv2d_same(x, k, b))


def seg_vanilla_forward(ws, x, depth, num_classes=1, rnd: Callable = _ident):
    """``build_unet`` (Segmenation/code/unet_vinillia.py:72-91)."""
    w = _W(ws)
    skips = []
    for d in range(depth):
        x = _conv_ln_relu(x, w, rnd, None, "")
        x = _conv_ln_relu(x, w, rnd, None, "")
        skips.append(x)
        x = K.max_pool2(x)
    x = _conv_ln_relu(x, w, rnd, None, "")
    x = _conv_ln_relu(x, w, rnd, None, "")
    for skip in reversed(skips):
        k, b = w.take(2)
        x = rnd(K.conv2d_transpose_2x2(x, k, b))
        x = torch.cat([x, skip], dim=-1)
        x = _conv_ln_relu(x, w, rnd, None, "")
        x = _conv_ln_relu(x, w, rnd, None, "")
    k, b = w.take(2)
    w.done()
    z = K.conv2d_same(x, k, b)
    return torch.sigmoid(z) if num_classes == 1 else torch.softmax(z, dim=-1)


def bf16_round(x):
    """Storage rounding of the bf16 policy, straight-through for autograd."""
    return x + (x.detach().to(torch.bfloat16).to(x.dtype) - x.detach())


class _Bf16Storage(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def bf16_storage(x):
    """A tensor STORED in bf16 under the mixed policy: its value is rounded, and so is its gradient -- the gradient of a
    bf16 tensor is carried in bf16 between ops (TensorFlow's Cast op back-propagates a Cast; the kernels here write
    activation gradients to bf16 buffers).  ``bf16_round`` above is the straight-through form (fp32 gradients): enough
    for LayerNorm nets, whose per-pixel statistics do not cancel across the batch.  BatchNorm backward subtracts
    per-channel means over N*H*W, so sum(dz) == 0 holds exactly only before dz is rounded; the filter gradient
    sum_p x[p] dz[p] then sees mean(x) * (rounding residue of sum dz), which is tens of percent of the true value on
    post-ReLU inputs.  That residue is a deterministic function of the rounded values, so an oracle that rounds the
    same tensors at the same points reproduces it."""
    return _Bf16Storage.apply(x)


def sr_flops_per_sample(scale, depth, input_size, base_channels=64, head=64) -> float:
    """Algorithmic forward FLOPs of every conv (2*H*W*Cin*Cout*k*k) for one sample."""
    sizes = resize_np.size_chain(input_size, scale, depth)
    fl = 0.0
    nf, cin = base_channels, 3
    for d in range(depth):
        s = sizes[d]
        fl += 2.0 * s * s * 9 * (cin * nf + nf * nf)
        cin, nf = nf, nf * 2
    s = sizes[depth]
    fl += 2.0 * s * s * 9 * (cin * nf + nf * nf)
    cin = nf
    for d in reversed(range(depth)):
        nf //= 2
        s = sizes[d]
        fl += 2.0 * s * s * 9 * (cin * nf + 2 * nf * nf + nf * nf)
        cin = nf
    s = sizes[0]
    fl += 2.0 * s * s * 9 * (cin * head + head * head) + 2.0 * s * s * head * 3
    return fl
