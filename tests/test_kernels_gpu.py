"""GPU parity of every kernel against the CPU oracle (bit-for-tolerance).

Tolerances (rel-L2 per tensor): fp32 kernels 2e-5 against the fp32 oracle;
bf16 kernels 1e-2 against the fp32 oracle fed the same bf16-rounded inputs.
"""
import numpy as np
import pytest
import torch

from util import conv_ws, TOL, f32, rand, relerr

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.bfloat16]


def _ops():
    import b200unet.ops as ops
    return ops


def _ffi():
    import b200unet._ffi as ffi
    return ffi


def _K():
    from oracle import keras_ops
    return keras_ops


# ----------------------------------------------------------------------------- conv (SIMT)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 13, 9, 3, 64, 3), (1, 16, 16, 64, 32, 3), (2, 8, 8, 20, 3, 1), (1, 5, 7, 64, 21, 1), (1, 9, 8, 64, 192, 3), (1, 8, 9, 192, 64, 3)])
def test_conv_simt(dtype, shape):
    ops, K = _ops(), _K()
    n, h, w, ci, co, ks = shape
    x = rand((n, h, w, ci), 1, dtype)
    wt = rand((ks, ks, ci, co), 2, dtype, 0.2)
    b = rand((co,), 3, torch.float32, 0.5)
    dy = rand((n, h, w, co), 4, dtype)
    filt = ops.ConvFilter(wt, packed=False)
    y = torch.empty((n, h, w, co), dtype=dtype, device="cuda")
    ops.conv2d_fprop(x, filt, b, y, ops.ACT_RELU, ops.ALGO_SIMT)
    dx = torch.empty_like(x)
    ops.conv2d_dgrad(dy, filt, dx, False, ops.ALGO_SIMT)
    dw = torch.empty((ks, ks, ci, co), dtype=torch.float32, device="cuda")
    ops.conv2d_wgrad(x, dy, ks, ks, dw, None, ops.ALGO_SIMT)
    xr, wr = f32(x).requires_grad_(), f32(wt).requires_grad_()
    yr = K.conv2d_same(xr, wr, f32(b))
    (yr * f32(dy)).sum().backward()
    assert relerr(y, torch.relu(yr)) < TOL[dtype]
    assert relerr(dx, xr.grad) < TOL[dtype]
    assert relerr(dw, wr.grad) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 37, 21, 3, 64, 3), (1, 16, 16, 3, 32, 3), (2, 19, 23, 64, 3, 1), (3, 9, 7, 64, 1, 1), (1, 8, 8, 128, 3, 1),
                                   # multi-class heads (4 <= Cout <= 32 from 32 / 64 channels): ragged pixel counts, odd Cout
                                   (2, 19, 23, 32, 21, 1), (1, 5, 5, 32, 21, 1), (1, 16, 16, 64, 5, 1), (3, 9, 7, 64, 24, 1),
                                   (2, 8, 8, 32, 4, 1), (2, 32, 32, 64, 21, 1)])
def test_conv_small_specialised(dtype, shape):
    """ALGO_AUTO routes the RGB stem (Cin=3), the Cout<=3 1x1 heads and the multi-class 1x1 heads to the specialised SIMT
    kernels."""
    ops, K = _ops(), _K()
    n, h, w, ci, co, ks = shape
    x = rand((n, h, w, ci), 1, dtype)
    wt = rand((ks, ks, ci, co), 2, dtype, 0.2)
    b = rand((co,), 3, torch.float32, 0.5)
    dy = rand((n, h, w, co), 4, dtype)
    filt = ops.ConvFilter(wt, packed=False)
    y = torch.empty((n, h, w, co), dtype=dtype, device="cuda")
    ops.conv2d_fprop(x, filt, b, y, ops.ACT_RELU)
    base = rand((n, h, w, ci), 5, dtype)
    dx = base.clone()
    ops.conv2d_dgrad(dy, filt, dx, True)
    dw = torch.full((ks, ks, ci, co), 3.0, dtype=torch.float32, device="cuda")
    ops.conv2d_wgrad(x, dy, ks, ks, dw, None)
    xr, wr = f32(x).requires_grad_(), f32(wt).requires_grad_()
    yr = K.conv2d_same(xr, wr, f32(b))
    (yr * f32(dy)).sum().backward()
    assert relerr(y, torch.relu(yr)) < TOL[dtype]
    assert relerr(dx, xr.grad + f32(base)) < TOL[dtype]
    assert relerr(dw, wr.grad) < TOL[dtype]


# ----------------------------------------------------------------------------- UMMA descriptor probe
def _probe_expected(a, b, start_rows, sbo_rows):
    rows = [start_rows + (r // 8) * sbo_rows + (r % 8) for r in range(128)]
    return a[rows].float() @ b.float().t()


@pytest.mark.parametrize("start_bytes,sbo_bytes", [(0, 1024), (1024, 1024), (128, 1024), (1408, 1024), (0, 1280), (1408, 1280), (2816, 1280)])
def test_umma_probe_kmajor(start_bytes, sbo_bytes):
    """K-major SW128 operand read through a shifted start address / non-1024 SBO.
    Expected under the absolute-address swizzle model: row r of the operand is the
    128-byte smem row  start/128 + (r/8)*(SBO/128) + r%8."""
    ops = _ops()
    a = rand((320, 64), 5, torch.bfloat16)
    b = rand((64, 64), 6, torch.bfloat16)
    out = torch.zeros((128, 64), dtype=torch.float32, device="cuda")
    ops.umma_probe(a, b, start_bytes, sbo_bytes, 0, False, out)
    torch.cuda.synchronize()
    exp = _probe_expected(a.cpu(), b.cpu(), start_bytes // 128, sbo_bytes // 128)
    err = relerr(out, exp)
    print(f"probe kmajor start={start_bytes} sbo={sbo_bytes} relerr={err:.3e}")
    assert err < 1e-5


@pytest.mark.parametrize("start_bytes,sbo_bytes,lbo_bytes", [(0, 1024, 1024), (0, 1024, 128), (1408, 1280, 128), (1408, 1280, 1024)])
def test_umma_probe_mnmajor(start_bytes, sbo_bytes, lbo_bytes):
    """wgrad-style MN-major operands: D[m][n] = sum_k A[m][k] B[n][k], m = atom*64 + c,
    A[m][k] = smem row (start + atom*LBO)/128 + (k/8)*(SBO/128) + k%8, channel c; B[n][k] = b[k][n]."""
    ops = _ops()
    a = rand((320, 64), 7, torch.bfloat16)
    b = rand((64, 64), 8, torch.bfloat16)
    out = torch.zeros((128, 64), dtype=torch.float32, device="cuda")
    ops.umma_probe(a, b, start_bytes, sbo_bytes, lbo_bytes, True, out)
    torch.cuda.synchronize()
    ac, bc = a.cpu().float(), b.cpu().float()
    exp = torch.zeros(128, 64)
    for atom in range(2):
        rows = [start_bytes // 128 + atom * (lbo_bytes // 128) + (k // 8) * (sbo_bytes // 128) + k % 8 for k in range(64)]
        exp[atom * 64:(atom + 1) * 64] = ac[rows].t() @ bc  # [c][k] @ [k][n]
    err = relerr(out, exp)
    print(f"probe mnmajor start={start_bytes} sbo={sbo_bytes} lbo={lbo_bytes} relerr={err:.3e}")
    assert err < 1e-5


# ----------------------------------------------------------------------------- conv (tcgen05)
TC_SHAPES = [
    (2, 32, 32, 64, 64), (1, 20, 13, 64, 64), (2, 16, 16, 128, 64), (2, 8, 8, 64, 128),
    (3, 40, 24, 128, 128), (4, 2, 2, 128, 256), (3, 1, 1, 256, 128), (1, 45, 45, 64, 64),
    # small images: stacked several per tile (H <= 7) / flattened (1x1)
    (16, 1, 1, 128, 128), (13, 1, 1, 64, 64), (9, 2, 2, 64, 64), (7, 3, 5, 64, 128), (5, 4, 4, 128, 64),
    (3, 5, 11, 64, 64), (5, 6, 6, 64, 64), (3, 7, 7, 64, 64), (11, 1, 6, 64, 64), (1, 2, 2, 64, 64),
    # 8-row images: two per tile with interleaved rows (pair tiles), odd batch, several tiles wide, streamed weights
    (3, 8, 8, 128, 128), (4, 8, 13, 64, 64), (5, 8, 8, 256, 256), (2, 8, 20, 128, 64),
    # 32-channel tensors on the 64-channel tiles (TMA zero fill / clipped stores): full-size, mixed widths, pair tiles,
    # stacked and flattened small images
    (2, 32, 32, 32, 32), (1, 20, 13, 64, 32), (2, 16, 16, 32, 64), (2, 24, 16, 32, 128), (2, 24, 16, 128, 32),
    (3, 8, 8, 32, 32), (9, 2, 2, 32, 32), (16, 1, 1, 32, 32), (5, 6, 6, 64, 32),
]


CG2_SHAPES = [  # CTA-pair (cta_group::2) coverage: odd tile counts (a padding tile in the last pair), several N tiles,
    # resident and streamed weight halves, every (Cin, Cout) class of the model
    (1, 16, 16, 64, 64), (3, 16, 8, 64, 64), (2, 48, 40, 128, 64), (1, 33, 9, 64, 128), (2, 32, 32, 128, 128),
    (1, 32, 24, 256, 128), (1, 16, 24, 128, 256), (1, 20, 20, 256, 256), (2, 128, 128, 64, 64),
]


@pytest.mark.parametrize("algo", ["pairs", "1cta"])
@pytest.mark.parametrize("shape", TC_SHAPES + CG2_SHAPES)
def test_conv_tc_fprop_dgrad(shape, algo):
    """tcgen05 implicit GEMM, fprop (+bias+ReLU) and dgrad, against the oracle: the default dispatch (CTA pairs =
    cta_group::2 wherever the shape allows: plain 16x8 tiles, channel counts multiples of 64) and the single-CTA kernel
    (B200_ALGO_TCGEN05_1CTA); the two must also agree with each other to bf16 rounding of identical fp32 sums."""
    ops, K = _ops(), _K()
    n, h, w, ci, co = shape
    dt = torch.bfloat16
    x = rand((n, h, w, ci), 11, dt)
    wt = rand((3, 3, ci, co), 12, dt, 0.1)
    b = rand((co,), 13, torch.float32, 0.5)
    dy = rand((n, h, w, co), 14, dt)
    filt = ops.ConvFilter(wt, packed=True)       # + the K-major pack the Cout = 64 pair kernel reads
    a = ops.ALGO_TCGEN05 if algo == "pairs" else ops.ALGO_TCGEN05_1CTA
    y = torch.full((n, h, w, co), 7.0, dtype=dt, device="cuda")
    cws = conv_ws(ops, x, filt, dy)
    ops.conv2d_fprop(x, filt, b, y, ops.ACT_RELU, a, ws=cws)
    dx = torch.full((n, h, w, ci), 7.0, dtype=dt, device="cuda")
    ops.conv2d_dgrad(dy, filt, dx, False, a, ws=cws)
    torch.cuda.synchronize()
    xr, wr = f32(x).requires_grad_(), f32(wt).requires_grad_()
    yr = K.conv2d_same(xr, wr, f32(b))
    (yr * f32(dy)).sum().backward()
    e1, e2 = relerr(y, torch.relu(yr)), relerr(dx, xr.grad)
    print(f"tc conv {shape} [{algo}]: fprop {e1:.3e} dgrad {e2:.3e}")
    assert e1 < 1e-2 and e2 < 1e-2
    if algo == "pairs":                          # same K order, same fp32 accumulation: the two kernels agree almost bit for bit
        y1, dx1 = torch.empty_like(y), torch.empty_like(dx)
        ops.conv2d_fprop(x, filt, b, y1, ops.ACT_RELU, ops.ALGO_TCGEN05_1CTA, ws=cws)
        ops.conv2d_dgrad(dy, filt, dx1, False, ops.ALGO_TCGEN05_1CTA, ws=cws)
        assert relerr(y, y1) < 2e-3 and relerr(dx, dx1) < 2e-3


@pytest.mark.parametrize("shape", [(2, 32, 32, 64, 64), (1, 20, 13, 128, 64), (2, 16, 24, 64, 128), (3, 40, 24, 128, 128),
                                   (9, 2, 2, 64, 64), (16, 1, 1, 128, 128), (5, 6, 6, 64, 128),
                                   (2, 8, 8, 64, 256), (2, 9, 9, 3, 64), (5, 8, 8, 64, 64), (3, 8, 11, 128, 128)])
@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("packed", [False, True])
def test_conv_ln_fused(shape, relu, packed):
    """Conv2D -> LayerNormalization -> ReLU as one call: fused tcgen05 epilogue (Cout 64/128; as CTA pairs when the
    filter carries its K-major pack or Cout is 128) and the in-library composition for the other shapes; z, y, mean,
    rstd against the oracle."""
    ops, K = _ops(), _K()
    n, h, w, ci, co = shape
    dt = torch.bfloat16
    x = rand((n, h, w, ci), 61, dt)
    wt = rand((3, 3, ci, co), 62, dt, 0.1)
    b = rand((co,), 63, torch.float32, 0.5)
    g = (1 + 0.3 * rand((co,), 64)).contiguous(); be = rand((co,), 65, scale=0.3)
    filt = ops.ConvFilter(wt, packed=packed)
    z = torch.full((n, h, w, co), 7.0, dtype=dt, device="cuda"); y = torch.full_like(z, 7.0)
    mean = torch.zeros(n * h * w, device="cuda"); rstd = torch.zeros_like(mean)
    ops.conv2d_ln_fprop(x, filt, b, g, be, 1e-3, relu, z, y, mean, rstd, ws=conv_ws(ops, x, filt))
    torch.cuda.synchronize()
    zr = K.conv2d_same(f32(x), f32(wt), f32(b))
    zr16 = zr.to(torch.bfloat16).float()
    yr = K.layer_norm(zr16, f32(g), f32(be))
    yr = torch.relu(yr) if relu else yr
    mr = zr16.mean(dim=-1).flatten()
    rr = torch.rsqrt(zr16.var(dim=-1, unbiased=False) + 1e-3).flatten()
    e = (relerr(z, zr), relerr(y, yr), relerr(mean, mr), relerr(rstd, rr))
    print(f"conv+ln {shape} relu={relu}: z {e[0]:.2e} y {e[1]:.2e} mean {e[2]:.2e} rstd {e[3]:.2e}")
    assert e[0] < 1e-2 and e[1] < 1.5e-2 and e[3] < 1e-2
    assert (f32(mean) - mr).abs().max() < 2e-2
    # inference form: no z kept
    y2 = torch.empty_like(y)
    ops.conv2d_ln_fprop(x, filt, b, g, be, 1e-3, relu, None, y2, mean, rstd, ws=conv_ws(ops, x, filt))
    assert relerr(y2, y) < 1e-6 or co > 128 or ci == 3


def test_conv_tc_strided_concat_and_accumulate():
    """Output written into a channel slice of a wider (concat) buffer; dgrad accumulating."""
    ops, K = _ops(), _K()
    dt = torch.bfloat16
    n, h, w = 2, 24, 16
    x = rand((n, h, w, 128), 21, dt)
    wt = rand((3, 3, 128, 64), 22, dt, 0.1)
    filt = ops.ConvFilter(wt, packed=True)
    cat = torch.zeros((n, h, w, 128), dtype=dt, device="cuda")
    ops.conv2d_fprop(x, filt, None, cat[..., 64:], ops.ACT_NONE, ops.ALGO_TCGEN05)
    yr = K.conv2d_same(f32(x), f32(wt))
    assert relerr(cat[..., 64:], yr) < 1e-2
    assert cat[..., :64].abs().max().item() == 0.0
    # a conv reading the slice as its input
    w2 = rand((3, 3, 64, 64), 23, dt, 0.1)
    f2 = ops.ConvFilter(w2)
    y2 = torch.empty((n, h, w, 64), dtype=dt, device="cuda")
    ops.conv2d_fprop(cat[..., 64:], f2, None, y2, ops.ACT_NONE, ops.ALGO_TCGEN05)
    assert relerr(y2, K.conv2d_same(f32(cat[..., 64:]), f32(w2))) < 1e-2
    # accumulate
    dy = rand((n, h, w, 64), 24, dt)
    base = rand((n, h, w, 128), 25, dt)
    dx = base.clone()
    ops.conv2d_dgrad(dy, filt, dx, True, ops.ALGO_TCGEN05)
    xr = f32(x).requires_grad_()
    (K.conv2d_same(xr, f32(wt)) * f32(dy)).sum().backward()
    assert relerr(dx, xr.grad + f32(base)) < 1e-2


def test_conv_tc_32_channel_slices_and_accumulate():
    """The 32-channel level of the reference's default segmentation U-Net (base_channels=32): producers write the two
    32-channel halves of a 64-channel concat buffer in place (the 64-wide store box is clipped at the slice's 32
    channels), a consumer reads one half (zero-filled to 64 by the load box), dgrad accumulates into a half."""
    ops, K = _ops(), _K()
    dt = torch.bfloat16
    n, h, w = 2, 24, 16
    x = rand((n, h, w, 32), 61, dt)
    w1 = rand((3, 3, 32, 32), 62, dt, 0.1)
    f1 = ops.ConvFilter(w1)
    cat = torch.full((n, h, w, 64), 3.0, dtype=dt, device="cuda")
    ops.conv2d_fprop(x, f1, None, cat[..., 32:], ops.ACT_NONE, ops.ALGO_TCGEN05)
    assert relerr(cat[..., 32:], K.conv2d_same(f32(x), f32(w1))) < 1e-2
    assert torch.equal(cat[..., :32], torch.full_like(cat[..., :32], 3.0))          # the other half is untouched
    ops.conv2d_fprop(x, f1, None, cat[..., :32], ops.ACT_RELU, ops.ALGO_TCGEN05)
    assert relerr(cat[..., :32], torch.relu(K.conv2d_same(f32(x), f32(w1)))) < 1e-2
    assert relerr(cat[..., 32:], K.conv2d_same(f32(x), f32(w1))) < 1e-2              # ... and so is the first one now
    y2 = torch.empty((n, h, w, 32), dtype=dt, device="cuda")
    ops.conv2d_fprop(cat[..., 32:], f1, None, y2, ops.ACT_NONE, ops.ALGO_TCGEN05)    # reads ONLY its half
    assert relerr(y2, K.conv2d_same(f32(cat[..., 32:]), f32(w1))) < 1e-2
    dy = rand((n, h, w, 32), 63, dt)
    base = rand((n, h, w, 64), 64, dt)
    dcat = base.clone()
    ops.conv2d_dgrad(dy, f1, dcat[..., 32:], True, ops.ALGO_TCGEN05)                 # accumulate into a half
    xr = f32(x).requires_grad_()
    (K.conv2d_same(xr, f32(w1)) * f32(dy)).sum().backward()
    assert relerr(dcat[..., 32:], xr.grad + f32(base[..., 32:])) < 1e-2
    assert torch.equal(dcat[..., :32], base[..., :32])


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_conv_tc_wgrad(shape):
    ops, K = _ops(), _K()
    n, h, w, ci, co = shape
    dt = torch.bfloat16
    x = rand((n, h, w, ci), 31, dt)
    dy = rand((n, h, w, co), 32, dt)
    dw = torch.full((3, 3, ci, co), 7.0, dtype=torch.float32, device="cuda")
    nbytes = ops.conv2d_wgrad_workspace(x, dy, 3, 3, ops.ALGO_TCGEN05)
    ws = torch.empty(max(nbytes, 16) // 4, dtype=torch.float32, device="cuda")
    ops.conv2d_wgrad(x, dy, 3, 3, dw, ws, ops.ALGO_TCGEN05)
    torch.cuda.synchronize()
    wr = torch.zeros(3, 3, ci, co, requires_grad=True)
    (K.conv2d_same(f32(x), wr) * f32(dy)).sum().backward()
    e = relerr(dw, wr.grad)
    print(f"tc wgrad {shape}: {e:.3e}")
    assert e < 2e-3


@pytest.mark.parametrize("shape", [(2, 32, 32, 64, 64), (1, 20, 13, 128, 64), (3, 16, 24, 64, 128), (2, 9, 7, 128, 256),
                                   (9, 2, 2, 64, 64), (16, 1, 1, 128, 128), (5, 6, 6, 64, 128),
                                   (2, 32, 32, 64, 32), (2, 9, 7, 32, 32), (1, 16, 24, 32, 64)])
def test_conv1x1_tc(shape):
    """1x1 convolutions on the tcgen05 kernels (centre tap of the 3x3 machinery): fprop (+ReLU), dgrad, wgrad."""
    ops, K = _ops(), _K()
    n, h, w, ci, co = shape
    dt = torch.bfloat16
    x = rand((n, h, w, ci), 41, dt)
    wt = rand((1, 1, ci, co), 42, dt, 0.1)
    b = rand((co,), 43, torch.float32, 0.5)
    dy = rand((n, h, w, co), 44, dt)
    filt = ops.ConvFilter(wt)
    y = torch.full((n, h, w, co), 7.0, dtype=dt, device="cuda")
    cws = conv_ws(ops, x, filt, dy)
    ops.conv2d_fprop(x, filt, b, y, ops.ACT_RELU, ops.ALGO_TCGEN05, ws=cws)
    dx = torch.full((n, h, w, ci), 7.0, dtype=dt, device="cuda")
    ops.conv2d_dgrad(dy, filt, dx, False, ops.ALGO_TCGEN05, ws=cws)
    dw = torch.full((1, 1, ci, co), 7.0, dtype=torch.float32, device="cuda")
    nbytes = ops.conv2d_wgrad_workspace(x, dy, 1, 1, ops.ALGO_TCGEN05)
    ws = torch.empty(max(nbytes, 16) // 4, dtype=torch.float32, device="cuda")
    ops.conv2d_wgrad(x, dy, 1, 1, dw, ws, ops.ALGO_TCGEN05)
    torch.cuda.synchronize()
    xr, wr = f32(x).requires_grad_(), f32(wt).requires_grad_()
    yr = K.conv2d_same(xr, wr, f32(b))
    (yr * f32(dy)).sum().backward()
    assert relerr(y, torch.relu(yr)) < 1e-2
    assert relerr(dx, xr.grad) < 1e-2
    assert relerr(dw, wr.grad) < 2e-3


@pytest.mark.parametrize("shape", [(8, 2, 2, 128, 256, 3), (16, 1, 1, 256, 128, 3), (4, 8, 8, 128, 128, 3), (3, 5, 7, 128, 128, 3),
                                   (2, 4, 4, 256, 128, 3), (64, 2, 2, 128, 128, 3), (5, 3, 3, 256, 64, 3), (7, 4, 2, 128, 256, 1),
                                   (130, 1, 1, 128, 128, 3), (1, 8, 8, 256, 384, 3)])
def test_conv_small_spatial_splitk(shape):
    """Deep-level layers (images <= 8x8): densely packed pixel boxes + split-K over (channel block, tap) units with
    an fp32 workspace and a finalize kernel -- fprop (+ReLU), dgrad (plain and accumulating), wgrad."""
    ops, K = _ops(), _K()
    n, h, w, ci, co, ks = shape
    dt = torch.bfloat16
    x = rand((n, h, w, ci), 61, dt)
    wt = rand((ks, ks, ci, co), 62, dt, 0.05)
    b = rand((co,), 63, torch.float32, 0.5)
    dy = rand((n, h, w, co), 64, dt)
    filt = ops.ConvFilter(wt)
    need = max(ops.conv2d_workspace(x, filt, False), ops.conv2d_workspace(dy, filt, True))
    assert need > 0 or max(h, w) > 4, "images of at most 4x4 pixels should take the split-K path"
    cws = ops.new_workspace(need, "cuda")       # (8x8 layers: split-K wgrad only; fprop/dgrad on the window kernel)
    if cws is not None:
        cws.fill_(float("nan"))                 # scratch contents must not matter
    wide = torch.zeros((n, h, w, co + 64), dtype=dt, device="cuda")
    y = wide[..., 64:]                       # channel slice: concat-in-place store
    if need > 0:
        # the path is chosen from the shapes alone: without (enough) scratch the call fails loudly, it does not switch kernels
        with pytest.raises(_ffi().B200Error, match="scratch"):
            ops.conv2d_fprop(x, filt, b, y, ops.ACT_RELU, ops.ALGO_TCGEN05)
    ops.conv2d_fprop(x, filt, b, y, ops.ACT_RELU, ops.ALGO_TCGEN05, ws=cws)
    dx = torch.full((n, h, w, ci), 7.0, dtype=dt, device="cuda")
    ops.conv2d_dgrad(dy, filt, dx, False, ops.ALGO_TCGEN05, ws=cws)
    base = rand((n, h, w, ci), 65, dt)
    dx2 = base.clone()
    ops.conv2d_dgrad(dy, filt, dx2, True, ops.ALGO_TCGEN05, ws=cws)
    dw = torch.full((ks, ks, ci, co), 7.0, dtype=torch.float32, device="cuda")
    nbytes = ops.conv2d_wgrad_workspace(x, dy, ks, ks, ops.ALGO_TCGEN05)
    ws = torch.empty(max(nbytes, 16) // 4, dtype=torch.float32, device="cuda")
    ops.conv2d_wgrad(x, dy, ks, ks, dw, ws, ops.ALGO_TCGEN05)
    torch.cuda.synchronize()
    xr, wr = f32(x).requires_grad_(), f32(wt).requires_grad_()
    yr = K.conv2d_same(xr, wr, f32(b))
    (yr * f32(dy)).sum().backward()
    assert relerr(y, torch.relu(yr)) < 1e-2
    assert float(wide[..., :64].abs().max()) == 0.0
    assert relerr(dx, xr.grad) < 1e-2
    assert relerr(dx2, xr.grad + f32(base)) < 1e-2
    assert relerr(dw, wr.grad) < 2e-3


@pytest.mark.parametrize("shape", [(2, 37, 21, 3), (3, 16, 16, 3), (1, 5, 9, 7), (4, 2, 2, 3), (2, 1, 1, 3), (2, 20, 13, 3, 32)])
@pytest.mark.parametrize("src_dtype", DTYPES)
def test_stem_im2col_tensor_core_path(shape, src_dtype):
    """The RGB stem as im2col (64 bf16 channels) + 1x1 tcgen05 convolution with the HWIO kernel zero-padded
    to [64][Cout]: conv + LayerNorm + ReLU forward and the filter gradient against the 3x3 oracle."""
    ops, K = _ops(), _K()
    n, h, w, ci = shape[:4]
    co = shape[4] if len(shape) > 4 else 64      # 32: composed conv + stand-alone LayerNorm (no fused epilogue)
    dt = torch.bfloat16
    x = rand((n, h, w, ci), 51, src_dtype)
    xb = x.to(dt)                                   # the engine feeds the stem the bf16 cast of the input
    wt = rand((3, 3, ci, co), 52, dt, 0.2)
    b = rand((co,), 53, torch.float32, 0.5)
    g, be = 1 + rand((co,), 54, torch.float32, 0.1), rand((co,), 55, torch.float32, 0.1)
    dy = rand((n, h, w, co), 56, dt)
    xcol = torch.full((n, h, w, 64), 7.0, dtype=dt, device="cuda")
    ops.im2col3x3(xb, xcol)
    # oracle im2col
    xp = torch.nn.functional.pad(f32(xb), (0, 0, 1, 1, 1, 1))
    ref = torch.zeros(n, h, w, 64)
    for kh in range(3):
        for kw in range(3):
            ref[..., (kh * 3 + kw) * ci:(kh * 3 + kw + 1) * ci] = xp[:, kh:kh + h, kw:kw + w, :]
    assert torch.equal(f32(xcol), ref)
    wpad = torch.zeros((1, 1, 64, co), dtype=dt, device="cuda")
    wpad[0, 0, :9 * ci] = wt.reshape(9 * ci, co)
    filt = ops.ConvFilter(wpad)
    z = torch.empty((n, h, w, co), dtype=dt, device="cuda"); y = torch.empty_like(z)
    mean = torch.empty(n * h * w, device="cuda"); rstd = torch.empty_like(mean)
    ops.conv2d_ln_fprop(xcol, filt, b, g, be, 1e-3, True, z, y, mean, rstd)
    dw = torch.full((64, co), 7.0, dtype=torch.float32, device="cuda")
    nbytes = ops.conv2d_wgrad_workspace(xcol, dy, 1, 1)
    ws = torch.empty(max(nbytes, 16) // 4, dtype=torch.float32, device="cuda")
    ops.conv2d_wgrad(xcol, dy, 1, 1, dw, ws)
    torch.cuda.synchronize()
    wr = f32(wt).requires_grad_()
    zr = K.conv2d_same(f32(xb), wr, f32(b))
    (zr * f32(dy)).sum().backward()
    assert relerr(z, zr) < 1e-2
    zq = f32(z)
    yr = torch.relu(K.layer_norm(zq, f32(g), f32(be), 1e-3))
    assert relerr(y, yr) < 1e-2
    assert relerr(dw[:9 * ci], wr.grad.reshape(9 * ci, co)) < 2e-3
    assert float(dw[9 * ci:].abs().max()) == 0.0


@pytest.mark.parametrize("shape", [(2, 32, 32, 64, 64, 3), (1, 20, 13, 128, 64, 3), (3, 40, 24, 128, 128, 3), (4, 2, 2, 128, 256, 3),
                                   (16, 1, 1, 128, 128, 3), (3, 8, 8, 128, 256, 3), (2, 16, 16, 64, 64, 1),
                                   (2, 32, 32, 32, 32, 3), (2, 16, 16, 64, 32, 3), (2, 16, 16, 32, 64, 1)])
def test_conv_wgrad_atomic(shape):
    """dw += wgrad through vector atomics (no workspace / reduce launch): exact accumulate semantics on a pre-filled dw."""
    ops, K = _ops(), _K()
    n, h, w, ci, co, ks = shape
    dt = torch.bfloat16
    x = rand((n, h, w, ci), 131, dt)
    dy = rand((n, h, w, co), 132, dt)
    base = rand((ks, ks, ci, co), 133, torch.float32)
    dw = base.clone()
    ops.conv2d_wgrad_atomic(x, dy, ks, ks, dw)
    torch.cuda.synchronize()
    wr = torch.zeros(ks, ks, ci, co, requires_grad=True)
    (K.conv2d_same(f32(x), wr) * f32(dy)).sum().backward()
    assert relerr(dw - base, wr.grad) < 2e-3


# ----------------------------------------------------------------------------- layer norm
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 7, 5, 64), (1, 3, 3, 8), (2, 4, 4, 256), (1, 2, 3, 1024), (1, 1, 5, 2048)])
@pytest.mark.parametrize("relu", [True, False])
def test_layernorm(dtype, shape, relu):
    ops, K = _ops(), _K()
    n, h, w, c = shape
    z = rand(shape, 41, dtype, 2.0)
    g = (1 + 0.3 * rand((c,), 42)).contiguous()
    b = rand((c,), 43, scale=0.3)
    dy = rand(shape, 44, dtype)
    y = torch.empty_like(z)
    mean = torch.empty(n * h * w, device="cuda"); rstd = torch.empty_like(mean)
    ops.layernorm_fwd(z, g, b, 1e-3, relu, y, mean, rstd)
    dz = torch.empty_like(z)
    dg = torch.zeros(c, device="cuda"); db = torch.zeros(c, device="cuda"); dbias = torch.zeros(c, device="cuda")
    ops.layernorm_bwd(dy, z, mean, rstd, g, b, relu, dz, dg, db, dbias)
    zr, gr, br = f32(z).requires_grad_(), f32(g).requires_grad_(), f32(b).requires_grad_()
    yr = K.layer_norm(zr, gr, br)
    yr = torch.relu(yr) if relu else yr
    (yr * f32(dy)).sum().backward()
    tol = TOL[dtype]
    assert relerr(y, yr) < tol
    assert relerr(dz, zr.grad) < tol * 2
    assert relerr(dg, gr.grad) < 1e-4 and relerr(db, br.grad) < 1e-4
    assert relerr(dbias, zr.grad.sum(dim=(0, 1, 2))) < 5e-2 or zr.grad.sum(dim=(0, 1, 2)).abs().max() < 1e-3


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("c,act", [(64, 1), (3, 0), (1, 2), (24, 1)])
def test_bias_act_bwd(dtype, c, act):
    ops = _ops()
    shape = (2, 5, 6, c)
    y = rand(shape, 51, dtype).abs() * (rand(shape, 52, dtype) > 0)
    if act == 2:
        y = torch.sigmoid(rand(shape, 51, torch.float32)).to(dtype)
    dy = rand(shape, 53, dtype)
    dz = torch.empty_like(dy)
    db = torch.zeros(c, device="cuda")
    ops.bias_act_bwd(dy, y, act, dz, db)
    yf, df = f32(y), f32(dy)
    exp = df if act == 0 else (df * (yf > 0) if act == 1 else df * yf * (1 - yf))
    assert relerr(dz, exp) < TOL[dtype]
    assert relerr(db, exp.sum(dim=(0, 1, 2))) < 1e-2


@pytest.mark.parametrize("shape", [(2, 37, 21, 3), (64, 16, 16, 3), (1, 1, 5, 3), (3, 8, 8, 3)])
def test_bias_grad_of_linear_rgb_head_in_place(shape):
    """bias_act_bwd(ACT_NONE) in place on a dense bf16 [..,3] tensor only has to produce the bias gradient:
    the column-sum kernel must leave the tensor untouched and accumulate (+=) into dbias."""
    ops = _ops()
    dy = rand(shape, 91, torch.bfloat16)
    keep = dy.clone()
    db = torch.full((3,), 0.5, device="cuda")
    ops.bias_act_bwd(dy, dy, 0, dy, db)
    torch.cuda.synchronize()
    assert torch.equal(dy, keep)
    assert relerr(db, f32(keep).sum(dim=(0, 1, 2)) + 0.5) < 1e-4


# ----------------------------------------------------------------------------- batch norm
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("c", [64, 24, 512])
def test_batchnorm(dtype, c):
    ops, K = _ops(), _K()
    shape = (3, 6, 5, c)
    z = rand(shape, 61, dtype, 2.0) + 0.5
    g = (1 + 0.3 * rand((c,), 62)).contiguous(); b = rand((c,), 63, scale=0.3)
    mm = rand((c,), 64, scale=0.1); mv = (1 + 0.2 * rand((c,), 65)).contiguous()
    dy = rand(shape, 66, dtype)
    y = torch.empty_like(z)
    sm = torch.empty(c, device="cuda"); sr = torch.empty(c, device="cuda")
    ws = torch.empty(2 * c, dtype=torch.float64, device="cuda")
    mm2, mv2 = mm.clone(), mv.clone()
    ops.batchnorm_fwd_train(z, g, b, 1e-3, 0.99, True, y, sm, sr, mm2, mv2, ws)
    dz = torch.empty_like(z); dg = torch.zeros(c, device="cuda"); db = torch.zeros(c, device="cuda")
    ops.batchnorm_bwd(dy, z, sm, sr, g, b, True, dz, dg, db, ws)
    yi = torch.empty_like(z)
    ops.batchnorm_fwd_infer(z, g, b, 1e-3, True, mm, mv, yi)
    zr, gr, br = f32(z).requires_grad_(), f32(g).requires_grad_(), f32(b).requires_grad_()
    yr, nm, nv = K.batch_norm_train(zr, gr, br, f32(mm), f32(mv))
    yr = torch.relu(yr)
    (yr * f32(dy)).sum().backward()
    tol = TOL[dtype]
    assert relerr(y, yr) < tol
    assert relerr(mm2, nm) < 1e-5 and relerr(mv2, nv) < 1e-5
    assert relerr(dz, zr.grad) < tol * 3
    assert relerr(dg, gr.grad) < 1e-3 and relerr(db, br.grad) < 1e-3
    assert relerr(yi, torch.relu(K.batch_norm_infer(f32(z), f32(g), f32(b), f32(mm), f32(mv)))) < tol


# ----------------------------------------------------------------------------- resampling / pooling
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("sizes", [((16, 16), (8, 8)), ((13, 9), (4, 3)), ((8, 8), (32, 32)), ((45, 45), (63, 63)), ((2, 2), (1, 1)), ((1, 1), (2, 2)), ((20, 20), (14, 14))])
@pytest.mark.parametrize("c", [64, 3])
@pytest.mark.parametrize("aa", [True, False])
def test_resample(dtype, sizes, c, aa):
    ops, K = _ops(), _K()
    (h, w), (oh, ow) = sizes
    x = rand((2, h, w, c), 71, dtype)
    dy = rand((2, oh, ow, c), 72, dtype)
    ph = ops.ResamplePlan(h, oh, aa, "cuda"); pw = ops.ResamplePlan(w, ow, aa, "cuda")
    y = torch.empty((2, oh, ow, c), dtype=dtype, device="cuda")
    ops.resample2d(x, y, ph, pw)
    base = rand((2, h, w, c), 73, dtype)
    dx = base.clone()
    ops.resample2d_bwd(dy, dx, ph, pw, accumulate=True)
    xr = f32(x).requires_grad_()
    yr = K.resize_bilinear(xr, oh, ow, aa)
    (yr * f32(dy)).sum().backward()
    assert relerr(y, yr) < TOL[dtype]
    assert relerr(dx, xr.grad + f32(base)) < TOL[dtype]


@pytest.mark.parametrize("sizes", [((128, 128), (32, 32)), ((32, 32), (128, 128)), ((63, 63), (45, 45)), ((45, 45), (63, 63)),
                                   ((100, 60), (90, 54)), ((90, 54), (100, 60)), ((33, 17), (9, 5)), ((9, 5), (33, 17)),
                                   ((64, 64), (13, 13)), ((256, 256), (180, 180))])
@pytest.mark.parametrize("c", [64, 128, 8])
def test_resample_marching_kernels(sizes, c):
    """bf16, C % 8 == 0: the row-marching kernels (up / down, 4 and 6 slots, several row segments) against
    the oracle, forward and accumulate-backward, on dense tensors and on a channel slice of a wider one."""
    ops, K = _ops(), _K()
    (h, w), (oh, ow) = sizes
    dtype = torch.bfloat16
    n = 3
    x = rand((n, h, w, c), 81, dtype)
    dy = rand((n, oh, ow, c), 82, dtype)
    ph = ops.ResamplePlan(h, oh, True, "cuda"); pw = ops.ResamplePlan(w, ow, True, "cuda")
    assert ph.mode in (1, 2, 3) and ph.t_mode in (1, 2, 3), (ph.mode, ph.t_mode)
    wide = torch.zeros((n, oh, ow, 2 * c), dtype=dtype, device="cuda")
    y = wide[..., c:]
    ops.resample2d(x, y, ph, pw)
    base = rand((n, h, w, c), 83, dtype)
    dx = base.clone()
    ops.resample2d_bwd(dy, dx, ph, pw, accumulate=True)
    dx2 = torch.empty_like(dx)
    ops.resample2d_bwd(dy, dx2, ph, pw, accumulate=False)
    xr = f32(x).requires_grad_()
    yr = K.resize_bilinear(xr, oh, ow, True)
    (yr * f32(dy)).sum().backward()
    assert relerr(y, yr) < TOL[dtype]
    assert float(wide[..., :c].abs().max()) == 0.0
    assert relerr(dx, xr.grad + f32(base)) < TOL[dtype]
    assert relerr(dx2, xr.grad) < TOL[dtype]


def test_resample_tables_match_oracle():
    ops = _ops()
    from oracle import resize_np
    for (a, b) in [(256, 180), (180, 126), (126, 89), (89, 63), (63, 45), (128, 32), (32, 8), (8, 2), (2, 1), (1, 2), (32, 128), (50, 16)]:
        for aa in (True, False):
            plan = ops.ResamplePlan(a, b, aa, "cuda")
            st, wt = resize_np.triangle_spans(a, b, aa)
            assert np.array_equal(plan.host[0], st)
            assert np.array_equal(plan.host[1], wt), (a, b, aa)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("hw", [(8, 8), (7, 5)])
def test_maxpool(dtype, hw):
    ops, K = _ops(), _K()
    h, w = hw
    x = rand((2, h, w, 16), 81, dtype)
    y = torch.empty((2, h // 2, w // 2, 16), dtype=dtype, device="cuda")
    ops.maxpool2_fwd(x, y)
    dy = rand(tuple(y.shape), 82, dtype)
    dx = torch.empty_like(x)
    ops.maxpool2_bwd(x, y, dy, dx)
    xr = f32(x).requires_grad_()
    yr = K.max_pool2(xr)
    (yr * f32(dy)).sum().backward()
    assert relerr(y, yr) < 1e-7
    assert relerr(dx, xr.grad) < 1e-7


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("c", [32, 3])
def test_maxpool_ties_slices_accumulate(dtype, c):
    """Vector (C % 8 == 0) and scalar kernels: windows with tied maxima (gradient to the FIRST maximum, row-major, as TF's
    MaxPoolGrad), inputs / gradients that are channel slices of wider buffers, accumulation into an existing gradient."""
    ops, K = _ops(), _K()
    n, h, w = 3, 12, 10
    g = torch.Generator().manual_seed(7)
    wide = (torch.randint(0, 3, (n, h, w, c + 8), generator=g).float() * 0.5).to(dtype).cuda()     # many ties
    x = wide[..., 8:] if c % 8 == 0 else wide[..., :c]
    y = torch.empty((n, h // 2, w // 2, c), dtype=dtype, device="cuda")
    ops.maxpool2_fwd(x, y)
    dy = rand(tuple(y.shape), 83, dtype)
    base = rand((n, h, w, c + 8), 84, dtype)
    dwide = base.clone()
    dx = dwide[..., 8:] if c % 8 == 0 else dwide[..., :c]
    ops.maxpool2_bwd(x, y, dy, dx, accumulate=True)
    xr = f32(x).requires_grad_()
    yr = K.max_pool2(xr)
    (yr * f32(dy)).sum().backward()
    assert torch.equal(f32(y), yr.detach())
    keep = base[..., :8] if c % 8 == 0 else base[..., c:]
    kept = dwide[..., :8] if c % 8 == 0 else dwide[..., c:]
    assert torch.equal(kept, keep)
    ref = xr.grad + f32(base[..., 8:] if c % 8 == 0 else base[..., :c])
    assert relerr(dx, ref) < (1e-7 if dtype == torch.float32 else 4e-3)
    # each window hands its gradient to exactly one element
    dx0 = torch.empty_like(x.contiguous())
    ops.maxpool2_bwd(x.contiguous(), y, dy, dx0)
    assert torch.equal(f32(dx0), xr.grad)


@pytest.mark.parametrize("dtype", DTYPES)
def test_conv_transpose(dtype):
    ops, K = _ops(), _K()
    x = rand((2, 5, 4, 16), 91, dtype); k = rand((2, 2, 8, 16), 92, dtype, 0.3); b = rand((8,), 93, scale=0.2)
    y = torch.empty((2, 10, 8, 8), dtype=dtype, device="cuda")
    ops.convT2x2_fprop(x, k, b, y)
    dy = rand((2, 10, 8, 8), 94, dtype)
    dx = torch.empty_like(x); dk = torch.empty((2, 2, 8, 16), device="cuda"); dbias = torch.zeros(8, device="cuda")
    ops.convT2x2_dgrad(dy, k, dx); ops.convT2x2_wgrad(x, dy, dk, dbias)
    xr, kr, br = f32(x).requires_grad_(), f32(k).requires_grad_(), f32(b).requires_grad_()
    yr = K.conv2d_transpose_2x2(xr, kr, br)
    (yr * f32(dy)).sum().backward()
    tol = TOL[dtype]
    assert relerr(y, yr) < tol and relerr(dx, xr.grad) < tol and relerr(dk, kr.grad) < tol and relerr(dbias, br.grad) < tol


@pytest.mark.parametrize("shape", [(2, 8, 8, 128, 64), (1, 4, 6, 64, 128), (3, 1, 1, 128, 128), (2, 16, 16, 256, 128), (5, 2, 2, 64, 64),
                                   (2, 8, 8, 64, 32), (1, 4, 6, 32, 32), (2, 16, 16, 32, 64)])
def test_conv_transpose_tensor_core(shape):
    """Conv2DTranspose(k2, s2) as four 1x1 tcgen05 convolutions over the parity views y[:, a::2, b::2, :] (fprop writes
    them, dgrad / wgrad read them), on a channel slice of a wider output buffer (concat in place)."""
    ops, K = _ops(), _K()
    n, h, w, ci, co = shape
    dt = torch.bfloat16
    x = rand((n, h, w, ci), 95, dt); k = rand((2, 2, co, ci), 96, dt, 0.1); b = rand((co,), 97, scale=0.2)
    wide = torch.zeros((n, 2 * h, 2 * w, co + 64), dtype=dt, device="cuda")
    y = wide[..., :co]
    dy = rand((n, 2 * h, 2 * w, co), 98, dt)
    cws = ops.new_workspace(max(ops.convT2x2_workspace(x, ci, co, False), ops.convT2x2_workspace(dy, co, ci, True)), "cuda")
    ops.convT2x2_fprop(x, k, b, y, ws=cws)
    dx = torch.full_like(x, 7.0); dk = torch.full((2, 2, co, ci), 7.0, device="cuda"); dbias = torch.zeros(co, device="cuda")
    ops.convT2x2_dgrad(dy, k, dx, ws=cws); ops.convT2x2_wgrad(x, dy, dk, dbias)
    xr, kr, br = f32(x).requires_grad_(), f32(k).requires_grad_(), f32(b).requires_grad_()
    yr = K.conv2d_transpose_2x2(xr, kr, br)
    (yr * f32(dy)).sum().backward()
    assert relerr(y, yr) < 1e-2 and float(wide[..., co:].abs().max()) == 0.0
    assert relerr(dx, xr.grad) < 1e-2 and relerr(dk, kr.grad) < 2e-3 and relerr(dbias, br.grad) < 1e-3


# ----------------------------------------------------------------------------- head / losses / optimiser
@pytest.mark.parametrize("dtype", DTYPES)
def test_clipadd(dtype):
    ops, K = _ops(), _K()
    inp = rand((2, 6, 6, 3), 101, torch.float32).abs()
    res = rand((2, 6, 6, 3), 102, dtype, 0.8)
    y = torch.empty((2, 6, 6, 3), dtype=dtype, device="cuda")
    ops.clipadd_fwd(inp, res, y)
    dy = rand((2, 6, 6, 3), 103, dtype)
    dres = torch.empty_like(res)
    ops.clipadd_bwd(inp, res, dy, dres)
    rr = f32(res).requires_grad_()
    yr = K.clipped_residual_add(f32(inp), rr)
    (yr * f32(dy)).sum().backward()
    assert relerr(y, yr) < TOL[dtype] and relerr(dres, rr.grad) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_sr_loss(dtype, kind):
    ops, K = _ops(), _K()
    pred = rand((3, 9, 7, 3), 111, torch.float32).abs().to(dtype)
    tgt = rand((3, 9, 7, 3), 112, torch.float32).abs()
    out = torch.zeros(2, device="cuda"); ws = torch.zeros(2 + 3, device="cuda")
    dp = torch.empty_like(pred)
    ops.sr_loss(pred, tgt, kind, 1e-3, 0.5, out, dp, ws)
    pr = f32(pred).requires_grad_()
    fn = [K.charbonnier_loss, K.l1_loss, K.mse_loss][kind]
    l = fn(f32(tgt), pr) if kind else K.charbonnier_loss(f32(tgt), pr, 1e-3)
    l.backward()
    assert abs(out[0].item() - l.item()) < 1e-5 * max(1, abs(l.item()))
    assert abs(out[1].item() - K.psnr_metric(f32(tgt), f32(pred)).item()) < 1e-3
    assert relerr(dp, 0.5 * pr.grad) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
def test_bce_dice(dtype):
    ops, K = _ops(), _K()
    pred = torch.sigmoid(3 * rand((3, 8, 8, 1), 121)).to(dtype)
    tgt = (rand((3, 8, 8, 1), 122) > 0).float()
    out = torch.zeros(5, device="cuda"); ws = torch.zeros(1 + 9, device="cuda"); dp = torch.empty_like(pred)
    ops.bce_dice_loss(pred, tgt, 0.4, 0.6, 1.0, out, dp, ws)
    pr = f32(pred).requires_grad_()
    l = K.bce_dice_loss(f32(tgt), pr, 0.4, 0.6)
    l.backward()
    assert abs(out[0].item() - l.item()) < 1e-5
    assert abs(out[2].item() - K.dice_coefficient(f32(tgt), f32(pred)).item()) < 1e-5
    assert abs(out[3].item() - K.iou_score(f32(tgt), f32(pred)).item()) < 1e-5
    assert abs(out[4].item() - K.dice_coefficient_global(f32(tgt), f32(pred)).item()) < 1e-5     # unet_vinillia's Dice
    assert relerr(dp, pr.grad) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
def test_softmax_ce(dtype):
    ops, K = _ops(), _K()
    z = rand((2, 5, 4, 21), 131, dtype, 3.0)
    labels = torch.randint(0, 21, (2, 5, 4), generator=torch.Generator().manual_seed(5), dtype=torch.int32).cuda()
    p = torch.empty_like(z)
    ops.softmax_fwd(z, p)
    out = torch.zeros(1, device="cuda"); ws = torch.zeros(1, device="cuda"); dz = torch.empty_like(z)
    ops.softmax_ce_loss(p, labels, 1.0, out, dz, ws)
    zr = f32(z).requires_grad_()
    pr = torch.softmax(zr, dim=-1)
    onehot = torch.nn.functional.one_hot(labels.cpu().long(), 21).float()
    l = K.categorical_crossentropy(onehot, pr)
    l.backward()
    assert relerr(p, pr) < TOL[dtype]
    assert abs(out[0].item() - l.item()) < (1e-5 if dtype == torch.float32 else 2e-2)
    assert relerr(dz, zr.grad) < (1e-4 if dtype == torch.float32 else 3e-2)
    # labels outside [0, C) (an "ignore" id such as 255, or -1): an all-zero one-hot row -- no loss, no gradient, no
    # out-of-bounds read; the mean keeps its denominator
    lab2 = labels.clone()
    lab2[0, 0, :2] = 255
    lab2[1, 2, 1] = -1
    ops.softmax_ce_loss(p, lab2, 1.0, out, dz, ws)
    oh2 = onehot.clone()
    oh2[0, 0, :2] = 0
    oh2[1, 2, 1] = 0
    pr2 = torch.softmax(f32(z), dim=-1)
    l2 = -(oh2 * torch.log(pr2.clamp(1e-7, 1 - 1e-7))).sum(-1).mean()
    assert abs(out[0].item() - l2.item()) < (1e-5 if dtype == torch.float32 else 2e-2)
    assert float(f32(dz)[0, 0, :2].abs().max()) == 0.0 and float(f32(dz)[1, 2, 1].abs().max()) == 0.0


def test_adam():
    ops, K = _ops(), _K()
    n = 1003
    p = rand((n,), 141); g = rand((n,), 142, scale=0.1); m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-7, 1 - 0.9, 1 - 0.999], device="cuda"); step = torch.zeros(1, dtype=torch.int32, device="cuda")
    shadow = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    pr, mr, vr = f32(p), torch.zeros(n), torch.zeros(n)
    for t in range(1, 4):
        ops.adam_advance(step)
        ops.adam_step(p, g, m, v, hyper, step, shadow)
        pr, mr, vr = K.adam_step(pr, f32(g), mr, vr, t, 1e-3)
    assert relerr(p, pr) < 1e-6 and relerr(m, mr) < 1e-6 and relerr(v, vr) < 1e-6
    assert relerr(shadow, pr) < 4e-3


def test_dynamic_loss_scale_kernels():
    """Device-resident LossScaleOptimizer state {scale, streak, found_inf, skipped}: the finite check, the skipped Adam step
    and the scale dynamics (x2 after `growth` finite steps in a row, /2 floor 1 on inf / nan) -- keras semantics of the
    reference's mixed_float16 policy (train_adaptive_unet.py:471-477)."""
    ops, K = _ops(), _K()
    n = 1003
    p = rand((n,), 141); g = rand((n,), 142, scale=0.1); m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-7, 1 - 0.9, 1 - 0.999], device="cuda"); step = torch.zeros(1, dtype=torch.int32, device="cuda")
    ls = torch.tensor([4.0, 0.0, 0.0, 0.0], device="cuda")
    gs = (g * 4.0).contiguous()                      # what backward produces from a loss gradient scaled by 4
    ops.loss_scale_check(gs, ls)
    assert ls.tolist() == [4.0, 0.0, 0.0, 0.0]
    ops.adam_advance(step, ls); ops.adam_step(p, gs, m, v, hyper, step, None, ls); ops.loss_scale_update(ls, 3)
    pr, mr, vr = K.adam_step(f32(rand((n,), 141)), f32(g), torch.zeros(n), torch.zeros(n), 1, 1e-3)
    assert relerr(p, pr) < 1e-6 and relerr(m, mr) < 1e-6 and int(step) == 1 and ls.tolist() == [4.0, 1.0, 0.0, 0.0]
    # a non-finite gradient: nothing moves, the scale halves, the streak restarts
    bad = gs.clone(); bad[517] = float("inf")
    p0, m0, v0 = p.clone(), m.clone(), v.clone()
    ops.loss_scale_check(bad, ls)
    assert ls[2].item() == 1.0
    ops.adam_advance(step, ls); ops.adam_step(p, bad, m, v, hyper, step, None, ls); ops.loss_scale_update(ls, 3)
    assert torch.equal(p, p0) and torch.equal(m, m0) and torch.equal(v, v0) and int(step) == 1
    assert ls.tolist() == [2.0, 0.0, 0.0, 1.0]
    bad[517] = float("nan"); bad[1002] = 1.0        # nan in the unaligned tail region too
    ops.loss_scale_check(bad, ls); ops.loss_scale_update(ls, 3)
    assert ls.tolist() == [1.0, 0.0, 0.0, 2.0]
    ops.loss_scale_check(bad, ls); ops.loss_scale_update(ls, 3)
    assert ls[0].item() == 1.0                       # floor
    for k in range(3):                               # three finite steps in a row double the scale
        ops.loss_scale_check(gs, ls); ops.loss_scale_update(ls, 3)
    assert ls[0].item() == 2.0 and ls[1].item() == 0.0
    x = rand((5, 7), 143, torch.bfloat16); x0 = x.clone()
    ops.loss_scale_apply(x, ls)
    assert torch.equal(x.float(), x0.float() * 2.0)


@pytest.mark.parametrize("dtype", DTYPES)
def test_binary_confusion_counters(dtype):
    """keras BinaryAccuracy / Precision / Recall (unet_vinillia.py:266-270) as tp / fp / fn / correct counters."""
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    pred = torch.rand((3, 17, 13, 1), generator=g).to(dtype).cuda()
    tgt = (torch.rand((3, 17, 13, 1), generator=g) > 0.6).float().cuda()
    counts = torch.full((4,), 7.0, device="cuda")
    ops.binary_confusion(pred, tgt, counts)
    pos, lab = pred.float().cpu() > 0.5, tgt.cpu() != 0
    want = [float((pos & lab).sum()), float((pos & ~lab).sum()), float((~pos & lab).sum()), float((pos.float() == tgt.cpu()).sum())]
    assert counts.tolist() == want


@pytest.mark.parametrize("shape", [(2, 32, 32, 64), (1, 20, 13, 128), (3, 16, 24, 64), (1, 33, 9, 128), (2, 128, 128, 64)])
@pytest.mark.parametrize("relu", [True, False])
def test_conv_dgrad_ln_bwd_fused(shape, relu):
    """dgrad with the LayerNorm(+ReLU) backward of its input's producer fused into the epilogue (CTA-pair kernel, 64-channel
    input): dz, d(gamma), d(beta) and d(bias) = sum(dz) against torch autograd through relu(LN(z)) -> conv, and against
    the unfused pair of kernels (b200_conv2d_dgrad + b200_layernorm_bwd) on the same buffers."""
    ops, K = _ops(), _K()
    n, h, w, co = shape
    ci, dt = 64, torch.bfloat16
    z = rand((n, h, w, ci), 171, dt, 2.0)
    wt = rand((3, 3, ci, co), 172, dt, 0.1)
    dy = rand((n, h, w, co), 173, dt)
    g = (1 + 0.3 * rand((ci,), 174)).contiguous(); be = rand((ci,), 175, scale=0.3)
    filt = ops.ConvFilter(wt)
    y = torch.empty_like(z); mean = torch.empty(n * h * w, device="cuda"); rstd = torch.empty_like(mean)
    ops.layernorm_fwd(z, g, be, 1e-3, relu, y, mean, rstd)
    dz = torch.full_like(z, 7.0)
    assert ops.conv2d_dgrad_ln_bwd_supported(dy, filt, dz)
    dg = torch.full((ci,), 0.5, device="cuda"); db = torch.full((ci,), -0.25, device="cuda"); dbias = torch.full((ci,), 2.0, device="cuda")
    ops.conv2d_dgrad_ln_bwd(dy, filt, z, mean, rstd, g, be, relu, dz, dg, db, dbias)
    torch.cuda.synchronize()
    # oracle: autograd through relu(LN(z)) -> conv
    zr, gr, br = f32(z).requires_grad_(), f32(g).requires_grad_(), f32(be).requires_grad_()
    yr = K.layer_norm(zr, gr, br)
    yr = torch.relu(yr) if relu else yr
    (K.conv2d_same(yr, f32(wt)) * f32(dy)).sum().backward()
    e = (relerr(dz, zr.grad), relerr(dg - 0.5, gr.grad), relerr(db + 0.25, br.grad), relerr(dbias - 2.0, zr.grad.sum(dim=(0, 1, 2))))
    print(f"dgrad+LN bwd {shape} relu={relu}: dz {e[0]:.2e} dgamma {e[1]:.2e} dbeta {e[2]:.2e} dbias {e[3]:.2e}")
    assert e[0] < 1e-2 and e[1] < 1e-2 and e[2] < 1e-2
    assert e[3] < 2e-2 or float(zr.grad.sum(dim=(0, 1, 2)).abs().max()) < 1e-3      # sum(dz): a cancellation (about 0 per pixel)
    # the unfused kernels on the same buffers: dz agrees to bf16 rounding of the intermediate dy they store
    dx = torch.empty_like(z); dz2 = torch.empty_like(z)
    dg2 = torch.zeros(ci, device="cuda"); db2 = torch.zeros(ci, device="cuda"); dbias2 = torch.zeros(ci, device="cuda")
    ops.conv2d_dgrad(dy, filt, dx, False)
    ops.layernorm_bwd(dx, z, mean, rstd, g, be, relu, dz2, dg2, db2, dbias2)
    assert relerr(dz, dz2) < 1e-2 and relerr(dg - 0.5, dg2) < 1e-2 and relerr(db + 0.25, db2) < 1e-2
    # shapes outside the fused kernel are refused, not mis-computed
    z32 = rand((n, h, w, 128), 176, dt)
    f2 = ops.ConvFilter(rand((3, 3, 128, 64), 177, dt, 0.1))
    assert not ops.conv2d_dgrad_ln_bwd_supported(rand((n, h, w, 64), 178, dt), f2, z32)
