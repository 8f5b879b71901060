"""Host-side operator wrappers: torch tensors in, C-ABI calls out.

PyTorch is used here only as the owner of device memory and streams; every
computation is a launch of one of the library's sm_100a kernels on torch's
current CUDA stream.  Tensors are NHWC (channel stride 1); channel-slice views of
wider buffers are passed with their real strides (concat written in place).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from . import _ffi
from ._ffi import (ACT_NONE, ACT_RELU, ACT_SIGMOID, ALGO_AUTO, ALGO_SIMT, ALGO_TCGEN05, ALGO_TCGEN05_1CTA, BF16, CV_INTER_AREA, CV_INTER_CUBIC,
                   F32, U8, Filter, Scratch, Tensor, check)

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def lib():
    return _ffi.load()


def launch_count(reset: bool = False) -> int:
    return int(lib().b200_launch_count(int(reset)))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def tdesc(t: torch.Tensor) -> Tensor:
    """b200_tensor for an NHWC torch tensor (any strides with channel stride 1)."""
    if t.dim() != 4:
        raise ValueError(f"expected NHWC tensor, got shape {tuple(t.shape)}")
    if t.shape[3] > 1 and t.stride(3) != 1:
        raise ValueError("channel stride must be 1")
    if not t.is_cuda:
        raise _ffi.B200Error("b200 ops need CUDA tensors: there is no CPU fallback")
    n, h, w, c = t.shape
    return Tensor(t.data_ptr(), n, h, w, c, t.stride(0), t.stride(1), t.stride(2), _DT[t.dtype], 0)


class ConvFilter:
    """A Conv2D kernel in the layouts the device code consumes (struct b200_filter)."""

    def __init__(self, hwio: torch.Tensor, packed: bool = False):
        self.hwio = hwio.contiguous()
        self.kh, self.kw, self.cin, self.cout = self.hwio.shape
        self.ohwi = None
        if packed:
            self.ohwi = torch.empty((self.kh, self.kw, self.cout, self.cin), dtype=hwio.dtype, device=hwio.device)
            self.repack()

    def repack(self):
        if self.ohwi is not None:
            check(lib().b200_filter_pack(_ptr(self.hwio), _ptr(self.ohwi), self.kh, self.kw, self.cin, self.cout,
                                         _DT[self.hwio.dtype], _stream()), "filter_pack")

    def struct(self) -> Filter:
        return Filter(self.hwio.data_ptr(), None if self.ohwi is None else self.ohwi.data_ptr(), self.kh, self.kw,
                      self.cin, self.cout, _DT[self.hwio.dtype], 0)


# ---- convolution -------------------------------------------------------------------
def conv2d_workspace(x, filt: "ConvFilter", dgrad: bool = False) -> int:
    """Bytes of split-K scratch this layer needs in `ws` (0 = it does not take the small-spatial path)."""
    return int(lib().b200_conv2d_workspace(tdesc(x), filt.struct(), int(dgrad)))


def convT2x2_workspace(x, cin: int, cout: int, dgrad: bool = False) -> int:
    return int(lib().b200_convT2x2_workspace(tdesc(x), int(cin), int(cout), int(dgrad)))


def _scratch(ws: Optional[torch.Tensor]):
    """b200_scratch for a caller-owned buffer (None -> NULL): passed per call, never registered."""
    if ws is None:
        return None
    return C.byref(Scratch(ws.data_ptr(), ws.numel() * ws.element_size()))


def new_workspace(nbytes: int, device) -> Optional[torch.Tensor]:
    """A scratch buffer of `nbytes` (None when 0); NaN-filled under B200_POISON=1."""
    if nbytes <= 0:
        return None
    buf = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=device)
    if os.environ.get("B200_POISON", "0") == "1":
        buf.fill_(float("nan"))
    return buf


def conv2d_fprop(x, filt: ConvFilter, bias, y, act=ACT_NONE, algo=ALGO_AUTO, ws=None):
    check(lib().b200_conv2d_fprop(tdesc(x), filt.struct(), _ptr(bias), tdesc(y), act, algo, _scratch(ws), _stream()),
          "conv2d_fprop")
    return y


def conv2d_ln_fprop(x, filt: ConvFilter, bias, gamma, beta, eps, relu, z, y, mean, rstd, algo=ALGO_AUTO, ws=None):
    """Conv2D -> LayerNormalization -> (ReLU) in one call; z (pre-norm, kept for backward) may be None."""
    check(lib().b200_conv2d_ln_fprop(tdesc(x), filt.struct(), _ptr(bias), _ptr(gamma), _ptr(beta), eps, int(relu),
                                     _opt_desc(z), tdesc(y), _ptr(mean), _ptr(rstd), algo, _scratch(ws), _stream()),
          "conv2d_ln_fprop")
    return y


def im2col3x3(x, xcol):
    """Narrow-input 3x3 im2col into 64 bf16 channels (the stem's tensor-core path)."""
    check(lib().b200_im2col3x3(tdesc(x), tdesc(xcol), _stream()), "im2col3x3")
    return xcol


def conv2d_dgrad(dy, filt: ConvFilter, dx, accumulate=False, algo=ALGO_AUTO, ws=None):
    check(lib().b200_conv2d_dgrad(tdesc(dy), filt.struct(), tdesc(dx), int(accumulate), algo, _scratch(ws), _stream()),
          "conv2d_dgrad")
    return dx


def conv2d_dgrad_ln_bwd_supported(dy, filt: ConvFilter, dz) -> bool:
    return bool(lib().b200_conv2d_dgrad_ln_bwd_supported(tdesc(dy), filt.struct(), tdesc(dz)))


def conv2d_dgrad_ln_bwd(dy, filt: ConvFilter, z, mean, rstd, gamma, beta, relu, dz, dgamma, dbeta, dbias):
    """dz = LayerNorm(+ReLU) backward of (conv dgrad of dy), the dgrad result never leaving the kernel; the parameter
    gradients are ADDED to dgamma / dbeta / dbias."""
    check(lib().b200_conv2d_dgrad_ln_bwd(tdesc(dy), filt.struct(), tdesc(z), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(beta),
                                         int(relu), tdesc(dz), _ptr(dgamma), _ptr(dbeta), _ptr(dbias), _stream()),
          "conv2d_dgrad_ln_bwd")
    return dz


def conv2d_wgrad_workspace(x, dy, kh, kw, algo=ALGO_AUTO) -> int:
    return int(lib().b200_conv2d_wgrad_workspace(tdesc(x), tdesc(dy), kh, kw, algo))


def conv2d_wgrad(x, dy, kh, kw, dw, workspace=None, algo=ALGO_AUTO):
    ws_bytes = 0 if workspace is None else workspace.numel() * workspace.element_size()
    check(lib().b200_conv2d_wgrad(tdesc(x), tdesc(dy), kh, kw, _ptr(dw), _ptr(workspace), ws_bytes, algo, _stream()),
          "conv2d_wgrad")
    return dw


def conv2d_wgrad_atomic(x, dy, kh, kw, dw):
    """dw += filter gradient through vector atomics (no workspace, no reduce launch); dw must be pre-zeroed."""
    check(lib().b200_conv2d_wgrad_atomic(tdesc(x), tdesc(dy), kh, kw, _ptr(dw), _stream()), "conv2d_wgrad_atomic")
    return dw


def convT2x2_fprop(x, kernel, bias, y, ws=None):
    check(lib().b200_convT2x2_fprop(tdesc(x), _ptr(kernel), _ptr(bias), kernel.shape[2], tdesc(y), _scratch(ws), _stream()),
          "convT2x2_fprop")
    return y


def convT2x2_dgrad(dy, kernel, dx, ws=None):
    check(lib().b200_convT2x2_dgrad(tdesc(dy), _ptr(kernel), kernel.shape[2], tdesc(dx), _scratch(ws), _stream()),
          "convT2x2_dgrad")
    return dx


def convT2x2_wgrad(x, dy, dkernel, dbias):
    check(lib().b200_convT2x2_wgrad(tdesc(x), tdesc(dy), _ptr(dkernel), _ptr(dbias), _stream()), "convT2x2_wgrad")


# ---- normalisation / activation -------------------------------------------------------
def bias_act_bwd(dy, y, act, dz, dbias=None):
    check(lib().b200_bias_act_bwd(tdesc(dy), tdesc(y), act, tdesc(dz), _ptr(dbias), _stream()), "bias_act_bwd")
    return dz


def layernorm_fwd(z, gamma, beta, eps, relu, y, mean, rstd):
    check(lib().b200_layernorm_fwd(tdesc(z), _ptr(gamma), _ptr(beta), eps, int(relu), tdesc(y), _ptr(mean), _ptr(rstd),
                                   _stream()), "layernorm_fwd")
    return y


def layernorm_bwd(dy, z, mean, rstd, gamma, beta, relu, dz, dgamma, dbeta, dbias):
    check(lib().b200_layernorm_bwd(tdesc(dy), tdesc(z), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(beta), int(relu),
                                   tdesc(dz), _ptr(dgamma), _ptr(dbeta), _ptr(dbias), _stream()), "layernorm_bwd")
    return dz


def batchnorm_fwd_train(z, gamma, beta, eps, momentum, relu, y, save_mean, save_rstd, moving_mean, moving_var, stats_ws):
    check(lib().b200_batchnorm_fwd_train(tdesc(z), _ptr(gamma), _ptr(beta), eps, momentum, int(relu), tdesc(y),
                                         _ptr(save_mean), _ptr(save_rstd), _ptr(moving_mean), _ptr(moving_var),
                                         _ptr(stats_ws), _stream()), "batchnorm_fwd_train")
    return y


def batchnorm_fwd_infer(z, gamma, beta, eps, relu, moving_mean, moving_var, y):
    check(lib().b200_batchnorm_fwd_infer(tdesc(z), _ptr(gamma), _ptr(beta), eps, int(relu), _ptr(moving_mean),
                                         _ptr(moving_var), tdesc(y), _stream()), "batchnorm_fwd_infer")
    return y


def batchnorm_bwd(dy, z, save_mean, save_rstd, gamma, beta, relu, dz, dgamma, dbeta, stats_ws):
    check(lib().b200_batchnorm_bwd(tdesc(dy), tdesc(z), _ptr(save_mean), _ptr(save_rstd), _ptr(gamma), _ptr(beta),
                                   int(relu), tdesc(dz), _ptr(dgamma), _ptr(dbeta), None, _ptr(stats_ws), _stream()),
          "batchnorm_bwd")
    return dz


def batchnorm_stats(z, stats_ws, dy=None, save_mean=None, save_rstd=None, gamma=None, beta=None, relu=False, dgamma=None,
                    dbeta=None):
    """Phase 1 of synchronised BatchNorm: per-channel (sum z, sum z^2) -- or, with dy, (sum g, sum g*xhat) plus the local
    dgamma / dbeta -- into stats_ws [2C] float64; the caller all-reduces stats_ws before the apply phase."""
    check(lib().b200_batchnorm_stats(tdesc(z), _opt_desc(dy), _ptr(save_mean), _ptr(save_rstd), _ptr(gamma), _ptr(beta),
                                     int(relu), _ptr(stats_ws), _ptr(dgamma), _ptr(dbeta), _stream()), "batchnorm_stats")


def batchnorm_fwd_apply(z, gamma, beta, eps, momentum, relu, y, save_mean, save_rstd, moving_mean, moving_var, stats_ws, count):
    check(lib().b200_batchnorm_fwd_apply(tdesc(z), _ptr(gamma), _ptr(beta), eps, momentum, int(relu), tdesc(y),
                                         _ptr(save_mean), _ptr(save_rstd), _ptr(moving_mean), _ptr(moving_var),
                                         _ptr(stats_ws), float(count), _stream()), "batchnorm_fwd_apply")
    return y


def batchnorm_bwd_apply(dy, z, save_mean, save_rstd, gamma, beta, relu, dz, stats_ws, count):
    check(lib().b200_batchnorm_bwd_apply(tdesc(dy), tdesc(z), _ptr(save_mean), _ptr(save_rstd), _ptr(gamma), _ptr(beta),
                                         int(relu), tdesc(dz), _ptr(stats_ws), float(count), _stream()), "batchnorm_bwd_apply")
    return dz


# ---- resampling --------------------------------------------------------------------------
def resize_extent(extent: int, scale: float) -> int:
    return int(lib().b200_resize_extent(int(extent), float(scale)))


class ResamplePlan:
    """Span tables (forward and transposed) for one axis, uploaded to the device."""

    def __init__(self, in_size: int, out_size: int, antialias: bool, device):
        L = lib()
        self.in_size, self.out_size = in_size, out_size
        taps = int(L.b200_resample_taps(in_size, out_size, int(antialias)))
        starts = np.zeros(out_size, dtype=np.int32)
        weights = np.zeros((out_size, taps), dtype=np.float32)
        check(L.b200_resample_plan(in_size, out_size, int(antialias), starts.ctypes.data, weights.ctypes.data, taps),
              "resample_plan")
        t_taps = int(L.b200_resample_plan_transpose(in_size, out_size, taps, starts.ctypes.data, weights.ctypes.data,
                                                    None, None, 0))
        t_starts = np.zeros(in_size, dtype=np.int32)
        t_weights = np.zeros((in_size, t_taps), dtype=np.float32)
        rc = L.b200_resample_plan_transpose(in_size, out_size, taps, starts.ctypes.data, weights.ctypes.data,
                                            t_starts.ctypes.data, t_weights.ctypes.data, t_taps)
        if rc < 0:
            check(rc, "resample_plan_transpose")
        self.host = (starts, weights, t_starts, t_weights)   # the raw ScaleAndTranslate tables (oracle-comparable)
        # device tables are the compacted ones (leading zero weights shifted out, rows re-packed to the
        # effective tap count); `mode` / `t_mode` name the row-marching kernel that may walk them
        cs, cw, self.taps, self.mode = self._compact(starts, weights)
        ts, tw, self.t_taps, self.t_mode = self._compact(t_starts, t_weights)
        self.compact_host = (cs, cw, ts, tw)
        self.starts = torch.from_numpy(cs).to(device)
        self.weights = torch.from_numpy(cw).to(device)
        self.t_starts = torch.from_numpy(ts).to(device)
        self.t_weights = torch.from_numpy(tw).to(device)

    @staticmethod
    def _compact(starts, weights):
        L = lib()
        st, w = starts.copy(), np.ascontiguousarray(weights.copy())
        n_out, taps = w.shape
        eff = int(L.b200_resample_compact(n_out, taps, st.ctypes.data, w.ctypes.data))
        if eff < 0:
            check(eff, "resample_compact")
        w = np.ascontiguousarray(w[:, :eff])
        mode = int(L.b200_resample_mode(n_out, eff, st.ctypes.data))
        return st, w, eff, max(mode, 0)


def resample2d(x, y, plan_h: ResamplePlan, plan_w: ResamplePlan, accumulate=False):
    check(lib().b200_resample2d_ex(tdesc(x), tdesc(y), _ptr(plan_h.starts), _ptr(plan_h.weights), plan_h.taps,
                                   _ptr(plan_w.starts), _ptr(plan_w.weights), plan_w.taps, int(accumulate), plan_h.mode,
                                   _stream()), "resample2d")
    return y


def resample2d_bwd(dy, dx, plan_h: ResamplePlan, plan_w: ResamplePlan, accumulate=False):
    """dx (+)= R^T dy: the same gather kernel over the transposed tables."""
    check(lib().b200_resample2d_ex(tdesc(dy), tdesc(dx), _ptr(plan_h.t_starts), _ptr(plan_h.t_weights), plan_h.t_taps,
                                   _ptr(plan_w.t_starts), _ptr(plan_w.t_weights), plan_w.t_taps, int(accumulate),
                                   plan_h.t_mode, _stream()), "resample2d(bwd)")
    return dx


def maxpool2_fwd(x, y):
    check(lib().b200_maxpool2_fwd(tdesc(x), tdesc(y), _stream()), "maxpool2_fwd")
    return y


def maxpool2_bwd(x, y, dy, dx, accumulate=False):
    check(lib().b200_maxpool2_bwd(tdesc(x), tdesc(y), tdesc(dy), tdesc(dx), int(accumulate), _stream()), "maxpool2_bwd")
    return dx


# ---- head / losses ---------------------------------------------------------------------------
def clipadd_fwd(inp, res, y):
    check(lib().b200_clipadd_fwd(tdesc(inp), tdesc(res), tdesc(y), _stream()), "clipadd_fwd")
    return y


def clipadd_bwd(inp, res, dy, dres):
    check(lib().b200_clipadd_bwd(tdesc(inp), tdesc(res), tdesc(dy), tdesc(dres), _stream()), "clipadd_bwd")
    return dres


def _opt_desc(t):
    return tdesc(t) if t is not None else Tensor(None, 0, 0, 0, 0, 0, 0, 0, 0, 0)


def sr_loss(pred, target, kind, eps, grad_scale, out, dpred, ws):
    check(lib().b200_sr_loss(tdesc(pred), tdesc(target), kind, eps, grad_scale, _ptr(out), _opt_desc(dpred), _ptr(ws),
                             _stream()), "sr_loss")
    return out


def bce_dice_loss(pred, target, bce_w, dice_w, grad_scale, out, dpred, ws):
    check(lib().b200_bce_dice_loss(tdesc(pred), tdesc(target), bce_w, dice_w, grad_scale, _ptr(out), _opt_desc(dpred),
                                   _ptr(ws), _stream()), "bce_dice_loss")
    return out


def binary_confusion(pred, target, counts, threshold=0.5):
    """counts[4] = {tp, fp, fn, correct} of (pred > threshold) against (target != 0)."""
    check(lib().b200_binary_confusion(tdesc(pred), tdesc(target), float(threshold), _ptr(counts), _stream()), "binary_confusion")
    return counts


def softmax_fwd(z, p):
    check(lib().b200_softmax_fwd(tdesc(z), tdesc(p), _stream()), "softmax_fwd")
    return p


def softmax_ce_loss(prob, labels, grad_scale, out, dlogits, ws):
    check(lib().b200_softmax_ce_loss(tdesc(prob), _ptr(labels), grad_scale, _ptr(out), _opt_desc(dlogits), _ptr(ws),
                                     _stream()), "softmax_ce_loss")
    return out


# ---- optimiser / utilities ------------------------------------------------------------------------
def adam_advance(step, loss_scale=None):
    check(lib().b200_adam_advance(_ptr(step), _ptr(loss_scale), _stream()), "adam_advance")


def adam_step(p, g, m, v, hyper, step, shadow=None, loss_scale=None):
    check(lib().b200_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), _ptr(hyper), _ptr(step), _ptr(shadow),
                               _ptr(loss_scale), _stream()), "adam_step")


def loss_scale_apply(t, loss_scale):
    """t *= scale (device-resident loss scale; t: the contiguous loss-gradient tensor)."""
    check(lib().b200_loss_scale_apply(_ptr(t), _DT[t.dtype], t.numel(), _ptr(loss_scale), _stream()), "loss_scale_apply")


def loss_scale_check(g, loss_scale):
    check(lib().b200_loss_scale_check(_ptr(g), g.numel(), _ptr(loss_scale), _stream()), "loss_scale_check")


def loss_scale_update(loss_scale, growth_interval=2000):
    check(lib().b200_loss_scale_update(_ptr(loss_scale), float(growth_interval), _stream()), "loss_scale_update")


def cast(src, dst):
    check(lib().b200_cast(_ptr(src), _DT[src.dtype], _ptr(dst), _DT[dst.dtype], src.numel(), _stream()), "cast")
    return dst


def copy_tensor(src, dst):
    check(lib().b200_copy_tensor(tdesc(src), tdesc(dst), _stream()), "copy_tensor")
    return dst


def scale_inplace(p, s):
    check(lib().b200_scale_inplace(_ptr(p), p.numel(), float(s), _stream()), "scale_inplace")


# ---- patch pipeline (shared/pipeline.py:79-136 of the reference, on the device) -----------------
class CvResizePlan:
    """OpenCV tap table of one axis (INTER_AREA shrink or INTER_CUBIC), built by the library, uploaded once."""

    def __init__(self, in_size: int, out_size: int, interp: int, device):
        L = lib()
        self.in_size, self.out_size, self.interp = int(in_size), int(out_size), int(interp)
        taps = int(L.b200_cv_resize_taps(self.in_size, self.out_size, self.interp))
        if taps < 0:
            check(taps, "cv_resize_taps")
        idx = np.zeros((self.out_size, taps), dtype=np.int32)
        w = np.zeros((self.out_size, taps), dtype=np.float32)
        check(L.b200_cv_resize_plan(self.in_size, self.out_size, self.interp, idx.ctypes.data, w.ctypes.data, taps),
              "cv_resize_plan")
        self.taps, self.host = taps, (idx, w)
        self.idx = torch.from_numpy(idx).to(device)
        self.weights = torch.from_numpy(w).to(device)


def patch_extract(image: torch.Tensor, origins: torch.Tensor, hr: torch.Tensor):
    """hr[i] = image[top_i:top_i+P, left_i:left_i+P] as fp32; image: cuda uint8/float32 [H,W,3]; origins: cuda int32 [n,2]."""
    if not image.is_cuda or image.dim() != 3 or image.shape[2] != 3 or not image.is_contiguous():
        raise ValueError("image must be a contiguous CUDA HxWx3 tensor")
    if image.dtype not in (torch.uint8, torch.float32):
        raise ValueError("image must be uint8 or float32")
    if origins.dtype != torch.int32 or not origins.is_cuda or tuple(origins.shape) != (hr.shape[0], 2):
        raise ValueError("origins must be a CUDA int32 [n,2] tensor")
    check(lib().b200_patch_extract(_ptr(image), U8 if image.dtype == torch.uint8 else F32, int(image.shape[0]),
                                   int(image.shape[1]), _ptr(origins.contiguous()), tdesc(hr), _stream()), "patch_extract")
    return hr


def gather2d(x, y, plan_h: CvResizePlan, plan_w: CvResizePlan, clip01: bool = False):
    check(lib().b200_gather2d(tdesc(x), tdesc(y), _ptr(plan_h.idx), _ptr(plan_h.weights), plan_h.taps, _ptr(plan_w.idx),
                              _ptr(plan_w.weights), plan_w.taps, int(clip01), _stream()), "gather2d")
    return y


def copy_rows(src, src_rows, dst, dst_rows, n_rows: int):
    """dst[dst_rows[r]] = src[src_rows[r]] over fp32 rows (first axis); either index tensor may be None (= r)."""
    row = src[0].numel()
    if dst[0].numel() != row or src.dtype != torch.float32 or dst.dtype != torch.float32:
        raise ValueError("copy_rows: fp32 tensors with equal row sizes expected")
    if not src.is_contiguous() or not dst.is_contiguous():
        raise ValueError("copy_rows: contiguous tensors expected")
    check(lib().b200_copy_rows(_ptr(src), _ptr(src_rows), _ptr(dst), _ptr(dst_rows), int(n_rows), row, _stream()),
          "copy_rows")
    return dst


# ---- evaluation metrics (train_adaptive_unet.py:144-157, 673-721) --------------------------------
def luma_pair(pred, hr, shave, pred_y, hr_y, sse):
    check(lib().b200_luma_pair(tdesc(pred), tdesc(hr), int(shave), _ptr(pred_y), _ptr(hr_y), _ptr(sse), _stream()),
          "luma_pair")


def ssim_planes(a, b, out, max_val: float = 1.0):
    """a, b: contiguous fp32 [n,h,w,ch]; out: fp32 [n*ch, 2] (sums of the SSIM / contrast-structure maps per plane)."""
    n, h, w, ch = a.shape
    check(lib().b200_ssim_planes(_ptr(a), _ptr(b), n, h, w, ch, float(max_val), _ptr(out), _stream()), "ssim_planes")
    return out


def avgpool2_planes(x, y):
    n, h, w, ch = x.shape
    check(lib().b200_avgpool2_planes(_ptr(x), n, h, w, ch, _ptr(y), _stream()), "avgpool2_planes")
    return y


def umma_probe(a, b, start_bytes, sbo_bytes, lbo_bytes, mn_major, out):
    check(lib().b200_debug_umma_probe(_ptr(a), a.shape[0], _ptr(b), start_bytes, sbo_bytes, lbo_bytes, int(mn_major),
                                      _ptr(out), _stream()), "umma_probe")
    return out
