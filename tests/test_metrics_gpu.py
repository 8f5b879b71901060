"""Evaluation-metric kernels (csrc/metrics.cu) through the C ABI against the fp64 oracle (oracle/metrics_ref.py):
BT.601 luma + shave + MSE/PSNR to 1e-6 / 1e-4 dB, SSIM and MS-SSIM to 2e-4 (fp32 E[x^2]-E[x]^2 cancellation; the
reference's tf.image runs the same arithmetic in fp32)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rgb_pair(n, h, w, seed, noise=0.05):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    base = torch.stack([0.5 + 0.3 * torch.sin(xx / 7) * torch.cos(yy / 5), 0.5 + 0.4 * torch.sin((xx + yy) / 11),
                        0.5 + 0.3 * torch.cos(xx / 4 - yy / 9)], dim=-1)
    hr = (base[None] + 0.1 * torch.randn((n, h, w, 3), generator=g)).clamp(0, 1)
    pred = hr + noise * torch.randn((n, h, w, 3), generator=g)          # leaves [0,1] here and there
    return pred, hr


def _oracle(pred, hr, shave):
    from oracle import metrics_ref as MR
    p = MR.rgb_to_luma_bt601(pred.double().clamp(0, 1))
    h = MR.rgb_to_luma_bt601(hr.double())
    if shave:
        p, h = p[:, shave:-shave, shave:-shave], h[:, shave:-shave, shave:-shave]
    return p, h


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shave", [0, 4])
def test_luma_pair_vs_oracle(dtype, shave):
    from b200unet import metrics as MT
    pred, hr = _rgb_pair(3, 37, 50, 1)
    pred = pred.to(dtype)
    py, hy, sse = MT.luma_planes(pred.cuda(), hr.cuda(), shave)
    p, h = _oracle(pred.float(), hr, shave)
    assert (py.cpu().double() - p[..., 0]).abs().max().item() <= 1e-6
    assert (hy.cpu().double() - h[..., 0]).abs().max().item() <= 1e-6
    want = ((p - h) ** 2).sum(dim=(1, 2, 3))
    assert ((sse.cpu().double() - want).abs() / want).max().item() <= 1e-5
    y = MT.rgb_to_luma_bt601(hr.cuda())
    assert tuple(y.shape) == (3, 37, 50, 1) and (y.cpu().double() - _oracle(hr, hr, 0)[1]).abs().max().item() <= 1e-6


@pytest.mark.parametrize("h,w", [(11, 11), (24, 24), (42, 43), (64, 33), (128, 128), (75, 140)])
def test_ssim_vs_oracle(h, w):
    from b200unet import metrics as MT
    from oracle import metrics_ref as MR
    pred, hr = _rgb_pair(3, h, w, h + w)
    p, t = _oracle(pred, hr, 0)
    got = MT.ssim(t.float().cuda(), p.float().cuda())
    want = MR.ssim(t, p).numpy()
    assert np.abs(got - want).max() <= 2e-4, (got, want)
    same = MT.ssim(t.float().cuda(), t.float().cuda())
    assert np.abs(same - 1.0).max() <= 1e-5


@pytest.mark.parametrize("h,w", [(176, 176), (181, 203), (256, 256)])
def test_msssim_vs_oracle(h, w):
    from b200unet import metrics as MT
    from oracle import metrics_ref as MR
    pred, hr = _rgb_pair(2, h, w, 7)
    p, t = _oracle(pred, hr, 0)
    got = MT.ssim_multiscale(t.float().cuda(), p.float().cuda())
    want = MR.ssim_multiscale(t, p).numpy()
    assert np.abs(got - want).max() <= 2e-4, (got, want)
    assert np.isnan(MT.ssim_multiscale(t[:, :175].float().cuda(), p[:, :175].float().cuda())).all()


def test_rgb_metrics_vs_oracle():
    """The vanilla baseline evaluates on RGB (u-net-vinillia.py:222-230): channel-interleaved tensors go through the same
    kernels, one plane per (image, channel)."""
    from b200unet import metrics as MT
    from oracle import metrics_ref as MR
    pred, hr = _rgb_pair(2, 181, 190, 3)
    pred = pred.clamp(0, 1)
    assert np.abs(MT.ssim(hr.cuda(), pred.cuda()) - MR.ssim(hr.double(), pred.double()).numpy()).max() <= 2e-4
    assert np.abs(MT.ssim_multiscale(hr.cuda(), pred.cuda()) - MR.ssim_multiscale(hr.double(), pred.double()).numpy()).max() <= 2e-4
    assert np.abs(MT.psnr_rgb(hr.cuda(), pred.cuda()) - MR.psnr(hr.double(), pred.double()).numpy()).max() <= 1e-3


def test_eval_luma_metrics_full_size():
    """One eval-loop iteration at the reference's evaluation patch size (256x256, shave 4, batch 8)."""
    from b200unet import metrics as MT
    from oracle import metrics_ref as MR
    pred, hr = _rgb_pair(8, 256, 256, 11, noise=0.02)
    out = MT.eval_luma_metrics(pred.cuda().to(torch.bfloat16), hr.cuda(), 4)
    p, t = _oracle(pred.to(torch.bfloat16).float(), hr, 4)
    mse = ((p - t) ** 2).mean(dim=(1, 2, 3)).numpy()
    assert np.abs(out["mse"] - mse).max() / mse.max() <= 1e-5
    assert np.abs(out["psnr"] - MR.psnr(t, p).numpy()).max() <= 1e-3          # dB
    assert np.abs(out["ssim"] - MR.ssim(t, p).numpy()).max() <= 2e-4
    assert np.abs(out["msssim"] - MR.ssim_multiscale(t, p).numpy()).max() <= 2e-4
    small = MT.eval_luma_metrics(pred[:, :16, :16].cuda(), hr[:, :16, :16].cuda(), 4)     # 8x8 planes: no SSIM window fits
    assert np.isnan(small["ssim"]).all() and np.isnan(small["msssim"]).all() and np.isfinite(small["psnr"]).all()
