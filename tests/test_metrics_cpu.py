"""Evaluation-metric oracle (oracle/metrics_ref.py, torch) against an independent numpy/scipy restatement of the
tf.image algorithms written the way TF writes them (2-D softmax Gaussian, one 11x11 VALID filter per moment),
and the host-side rules of the product module.  TensorFlow is absent: two restatements pin each other."""
import numpy as np
import pytest
import torch

from oracle import metrics_ref as MR


def _tf_gauss2d(size=11, sigma=1.5):
    c = np.arange(size, dtype=np.float64) - (size - 1) / 2.0
    g = -0.5 * c * c / (sigma * sigma)
    g = g[None, :] + g[:, None]
    e = np.exp(g - g.max())
    return e / e.sum()


def _valid_filter(x, k):
    from scipy.signal import correlate2d
    return correlate2d(x, k, mode="valid")


def _np_ssim_cs(a, b, max_val=1.0):
    k = _tf_gauss2d()
    c1, c2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    m0, m1 = _valid_filter(a, k), _valid_filter(b, k)
    num0, den0 = 2.0 * m0 * m1, m0 * m0 + m1 * m1
    lum = (num0 + c1) / (den0 + c1)
    num1 = 2.0 * _valid_filter(a * b, k)
    den1 = _valid_filter(a * a + b * b, k)
    cs = (num1 - num0 + c2) / (den1 - den0 + c2)
    return (lum * cs).mean(), cs.mean()


def _np_pool(x):
    h, w = x.shape
    x = np.pad(x, ((0, h % 2), (0, w % 2)), mode="symmetric")
    return 0.25 * (x[0::2, 0::2] + x[0::2, 1::2] + x[1::2, 0::2] + x[1::2, 1::2])


def _np_msssim(a, b):
    w = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)
    out = 1.0
    for i, wi in enumerate(w):
        s, cs = _np_ssim_cs(a, b)
        out *= max(s if i == len(w) - 1 else cs, 0.0) ** wi
        if i < len(w) - 1:
            a, b = _np_pool(a), _np_pool(b)
    return out


def _pair(h, w, seed, noise=0.08):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    a = 0.5 + 0.3 * np.sin(xx / 6.0) * np.cos(yy / 9.0) + 0.1 * rng.standard_normal((h, w))
    b = a + noise * rng.standard_normal((h, w))
    return np.clip(a, 0, 1), np.clip(b, 0, 1)


def test_luma_and_psnr_rules():
    rgb = torch.tensor([[[[1.0, 1.0, 1.0], [0.0, 0.0, 0.0], [0.2, 0.5, 0.9]]]])
    y = MR.rgb_to_luma_bt601(rgb).numpy().ravel()
    want = (np.array([[1, 1, 1], [0, 0, 0], [0.2, 0.5, 0.9]]) @ np.array([65.481, 128.553, 24.966]) + 16.0) / 255.0
    assert np.allclose(y, want, atol=1e-6)                 # white = 235/255, black = 16/255 (studio range)
    a = torch.zeros(1, 4, 4, 1)
    b = torch.full((1, 4, 4, 1), 0.1)
    assert abs(MR.psnr(a, b).item() - 20.0) < 1e-4


@pytest.mark.parametrize("h,w", [(40, 52), (11, 11), (64, 33)])
def test_ssim_oracle_vs_numpy(h, w):
    a, b = _pair(h, w, h * 100 + w)
    ta, tb = torch.from_numpy(a)[None, :, :, None], torch.from_numpy(b)[None, :, :, None]    # float64
    s, cs = MR._ssim_cs(ta, tb)
    ns, ncs = _np_ssim_cs(a, b)
    assert tuple(s.shape) == (1, 1) and abs(s.item() - ns) < 1e-10 and abs(cs.item() - ncs) < 1e-10
    assert abs(MR.ssim(ta.float(), tb.float()).item() - ns) < 2e-4       # fp32 evaluation (cancellation in E[x^2]-E[x]^2)
    assert abs(MR.ssim(ta, ta).item() - 1.0) < 1e-12


@pytest.mark.parametrize("h,w", [(176, 176), (181, 203)])
def test_msssim_oracle_vs_numpy(h, w):
    a, b = _pair(h, w, 5)
    ta, tb = torch.from_numpy(a)[None, :, :, None], torch.from_numpy(b)[None, :, :, None]
    assert abs(MR.ssim_multiscale(ta, tb).item() - _np_msssim(a, b)) < 1e-10
    assert torch.isnan(MR.ssim_multiscale(ta[:, :175], tb[:, :175])).all()   # < 11 * 2^4 pixels: TF raises


def test_multichannel_is_per_channel_then_mean():
    """tf.image.ssim / ssim_multiscale on RGB (u-net-vinillia.py:222-230): every channel on its own, then the channel mean
    (for MS-SSIM the product over scales is taken per channel first)."""
    a0, b0 = _pair(180, 190, 1)
    a1, b1 = _pair(180, 190, 2, noise=0.2)
    ta = torch.from_numpy(np.stack([a0, a1], -1))[None]
    tb = torch.from_numpy(np.stack([b0, b1], -1))[None]
    want = 0.5 * (_np_ssim_cs(a0, b0)[0] + _np_ssim_cs(a1, b1)[0])
    assert abs(MR.ssim(ta, tb).item() - want) < 1e-10
    want = 0.5 * (_np_msssim(a0, b0) + _np_msssim(a1, b1))
    assert abs(MR.ssim_multiscale(ta, tb).item() - want) < 1e-10


def test_product_metrics_need_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from b200unet import _ffi, metrics as MT
    with pytest.raises(_ffi.B200Error):
        MT.eval_luma_metrics(np.zeros((1, 16, 16, 3), np.float32), np.zeros((1, 16, 16, 3), np.float32))
