"""The oracle's restatements of the Keras ops against INDEPENDENT implementations of the same published semantics that
exist in this image (torch.nn.functional's library kernels, plain numpy): two code paths written from the same
definition pin each other.  (TensorFlow / Keras themselves are absent; see oracle/__init__.py for what is and is not pinned.)"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import keras_ops as K


def _r(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * scale


def test_conv2d_same_vs_direct_loops_and_library():
    x, w, b = _r(2, 7, 6, 5, seed=1), _r(3, 3, 5, 4, seed=2), _r(4, seed=3)
    y = K.conv2d_same(x, w, b)
    lib = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, padding=1).permute(0, 2, 3, 1)
    assert (y - lib).abs().max() < 1e-12
    xp = np.pad(x.numpy(), ((0, 0), (1, 1), (1, 1), (0, 0)))
    ref = np.zeros((2, 7, 6, 4))
    for kh in range(3):
        for kw in range(3):
            ref += np.einsum("nhwc,co->nhwo", xp[:, kh:kh + 7, kw:kw + 6, :], w.numpy()[kh, kw])     # cross-correlation
    assert np.abs(y.numpy() - (ref + b.numpy())).max() < 1e-12
    w1 = _r(1, 1, 5, 3, seed=4)
    assert (K.conv2d_same(x, w1) - torch.einsum("nhwc,co->nhwo", x, w1[0, 0])).abs().max() < 1e-12


def test_conv2d_transpose_vs_library():
    x, k, b = _r(2, 4, 5, 6, seed=5), _r(2, 2, 3, 6, seed=6), _r(3, seed=7)         # keras kernel [kh, kw, Cout, Cin]
    y = K.conv2d_transpose_2x2(x, k, b)
    lib = F.conv_transpose2d(x.permute(0, 3, 1, 2), k.permute(3, 2, 0, 1), b, stride=2).permute(0, 2, 3, 1)
    assert y.shape == (2, 8, 10, 3) and (y - lib).abs().max() < 1e-12
    i, j, a, c = 2, 3, 1, 0
    want = (x[0, i, j] * k[a, c, 1]).sum() + b[1]                                      # out[2i+a, 2j+b, o] = sum_c in[i,j,c] K[a,b,o,c]
    assert abs(y[0, 2 * i + a, 2 * j + c, 1] - want) < 1e-12


def test_normalisations_vs_library():
    x, g, be = _r(3, 4, 5, 16, seed=8, scale=3), 1 + _r(16, seed=9, scale=0.2), _r(16, seed=10, scale=0.2)
    assert (K.layer_norm(x, g, be, 1e-3) - F.layer_norm(x, (16,), g, be, eps=1e-3)).abs().max() < 1e-12
    mm, mv = _r(16, seed=11), 1 + _r(16, seed=12, scale=0.5)
    y, nm, nv = K.batch_norm_train(x, g, be, mm, mv, momentum=0.99, eps=1e-3)
    rm, rv = mm.clone(), mv.clone()
    lib = F.batch_norm(x.permute(0, 3, 1, 2), rm, rv, g, be, training=True, momentum=0.01, eps=1e-3).permute(0, 2, 3, 1)
    assert (y - lib).abs().max() < 1e-12 and (nm - rm).abs().max() < 1e-12
    n = 3 * 4 * 5                                           # torch tracks the UNBIASED variance, keras 3 the biased one
    batch_var_unbiased = (rv - mv * 0.99) / 0.01
    assert (nv - (mv * 0.99 + batch_var_unbiased * (n - 1) / n * 0.01)).abs().max() < 1e-12
    yi = K.batch_norm_infer(x, g, be, mm, mv, 1e-3)
    libi = F.batch_norm(x.permute(0, 3, 1, 2), mm, mv, g, be, training=False, eps=1e-3).permute(0, 2, 3, 1)
    assert (yi - libi).abs().max() < 1e-12


def test_pool_and_upsample_vs_library():
    x = _r(2, 7, 9, 3, seed=13)
    y = K.max_pool2(x)
    ref = x[:, :6, :8].reshape(2, 3, 2, 4, 2, 3).amax(dim=(2, 4))                      # valid, stride 2, floor
    assert y.shape == (2, 3, 4, 3) and torch.equal(y, ref)
    up = K.upsample2_bilinear(x.float())
    lib = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    assert up.shape == (2, 14, 18, 3) and (up - lib).abs().max() < 1e-6                # half-pixel centres, edges clamped


def test_losses_vs_library():
    t = (torch.rand(3, 6, 6, 1, generator=torch.Generator().manual_seed(14)) > 0.5).double()
    p = torch.rand(3, 6, 6, 1, generator=torch.Generator().manual_seed(15), dtype=torch.float64) * 0.98 + 0.01
    assert abs(K.binary_crossentropy(t, p) - F.binary_cross_entropy(p, t)) < 1e-12
    p[0, 0, 0, 0], p[0, 0, 1, 0] = 0.0, 1.0                                             # clipped to [1e-7, 1 - 1e-7], not to log >= -100
    want = -(t * torch.log(p.clamp(1e-7, 1 - 1e-7)) + (1 - t) * torch.log(1 - p.clamp(1e-7, 1 - 1e-7))).mean()
    assert abs(K.binary_crossentropy(t, p) - want) < 1e-12
    logits = _r(2, 5, 4, 7, seed=16, scale=3)
    prob = torch.softmax(logits, dim=-1)
    labels = torch.randint(0, 7, (2, 5, 4), generator=torch.Generator().manual_seed(17))
    onehot = F.one_hot(labels, 7).double()
    lib = F.cross_entropy(logits.permute(0, 3, 1, 2), labels)
    assert abs(K.categorical_crossentropy(onehot, prob) - lib) < 1e-9
    a, b = _r(2, 8, 8, 3, seed=18).abs(), _r(2, 8, 8, 3, seed=19).abs()
    assert abs(K.mse_loss(a, b) - F.mse_loss(b, a)) < 1e-12 and abs(K.l1_loss(a, b) - F.l1_loss(b, a)) < 1e-12
    mse = ((a.clamp(0, 1) - b.clamp(0, 1)) ** 2).mean(dim=(1, 2, 3))
    assert abs(K.psnr_metric(a.clamp(0, 1), b) - (10 * torch.log10(1 / mse)).mean()) < 1e-9


def test_adam_vs_closed_form_and_library_in_the_eps_free_limit():
    """keras Adam: p -= lr * sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps).  torch.optim.Adam puts eps elsewhere
    (sqrt(v / (1 - b2^t)) + eps), so the two coincide only when eps is negligible: checked with large gradients."""
    p0, g = _r(50, seed=20), _r(50, seed=21, scale=5.0)
    g = torch.sign(g) * (5.0 + g.abs())                    # bounded away from zero: eps stays negligible
    p, m, v = p0.clone(), torch.zeros(50, dtype=torch.float64), torch.zeros(50, dtype=torch.float64)
    tp = p0.clone().requires_grad_()
    opt = torch.optim.Adam([tp], lr=1e-3, betas=(0.9, 0.999), eps=1e-7)
    for step in range(1, 6):
        gi = g * (1 + 0.1 * step)
        p, m, v = K.adam_step(p, gi, m, v, step, 1e-3)
        tp.grad = gi.clone()
        opt.step()
        mm = sum((1 - 0.9) * 0.9 ** (step - s) * g * (1 + 0.1 * s) for s in range(1, step + 1))          # closed forms
        vv = sum((1 - 0.999) * 0.999 ** (step - s) * (g * (1 + 0.1 * s)) ** 2 for s in range(1, step + 1))
        assert (m - mm).abs().max() < 1e-12 and (v - vv).abs().max() < 1e-12
    assert (p - tp.detach()).abs().max() < 1e-8
