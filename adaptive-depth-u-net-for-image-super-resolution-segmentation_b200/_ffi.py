"""ctypes binding of libb200unet.so (the C ABI declared in include/b200_unet.h).

The library is built in-tree by ``build.py`` (nvcc, sm_100a).  There is no
fallback: if the shared object is missing the import raises, and every failing
call raises ``B200Error`` with the library's message.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "lib" / "libb200unet.so"

F32, BF16, U8 = 0, 1, 2
CV_INTER_AREA, CV_INTER_CUBIC = 0, 1
ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
ALGO_AUTO, ALGO_SIMT, ALGO_TCGEN05, ALGO_TCGEN05_1CTA = 0, 1, 2, 3
LOSS_CHARBONNIER, LOSS_L1, LOSS_MSE = 0, 1, 2


class B200Error(RuntimeError):
    pass


class Tensor(C.Structure):
    """struct b200_tensor"""
    _fields_ = [
        ("data", C.c_void_p),
        ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
        ("stride_n", C.c_int64), ("stride_h", C.c_int64), ("stride_w", C.c_int64),
        ("dtype", C.c_int32), ("reserved", C.c_int32),
    ]


class Filter(C.Structure):
    """struct b200_filter"""
    _fields_ = [
        ("hwio", C.c_void_p), ("ohwi", C.c_void_p),
        ("kh", C.c_int32), ("kw", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32),
        ("dtype", C.c_int32), ("reserved", C.c_int32),
    ]


class Scratch(C.Structure):
    """struct b200_scratch"""
    _fields_ = [("ptr", C.c_void_p), ("bytes", C.c_size_t)]


_TP = C.POINTER(Tensor)
_SP = C.POINTER(Scratch)
_FP = C.POINTER(Filter)
_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol of include/b200_unet.h
SIGNATURES = {
    "b200_version": (C.c_char_p, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_device_info": (_i, [C.POINTER(_i)] * 3),
    "b200_launch_count": (C.c_longlong, [_i]),
    "b200_conv2d_fprop": (_i, [_TP, _FP, _vp, _TP, _i, _i, _SP, _vp]),
    "b200_conv2d_ln_fprop": (_i, [_TP, _FP, _vp, _vp, _vp, _f, _i, _TP, _TP, _vp, _vp, _i, _SP, _vp]),
    "b200_conv2d_dgrad": (_i, [_TP, _FP, _TP, _i, _i, _SP, _vp]),
    "b200_conv2d_dgrad_ln_bwd_supported": (_i, [_TP, _FP, _TP]),
    "b200_conv2d_dgrad_ln_bwd": (_i, [_TP, _FP, _TP, _vp, _vp, _vp, _vp, _i, _TP, _vp, _vp, _vp, _vp]),
    "b200_conv2d_workspace": (_sz, [_TP, C.POINTER(Filter), _i]),
    "b200_conv2d_wgrad_workspace": (_sz, [_TP, _TP, _i, _i, _i]),
    "b200_conv2d_wgrad": (_i, [_TP, _TP, _i, _i, _vp, _vp, _sz, _i, _vp]),
    "b200_conv2d_wgrad_atomic": (_i, [_TP, _TP, _i, _i, _vp, _vp]),
    "b200_filter_pack": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "b200_im2col3x3": (_i, [_TP, _TP, _vp]),
    "b200_convT2x2_fprop": (_i, [_TP, _vp, _vp, _i, _TP, _SP, _vp]),
    "b200_convT2x2_dgrad": (_i, [_TP, _vp, _i, _TP, _SP, _vp]),
    "b200_convT2x2_workspace": (_sz, [_TP, _i, _i, _i]),
    "b200_convT2x2_wgrad": (_i, [_TP, _TP, _vp, _vp, _vp]),
    "b200_bias_act_bwd": (_i, [_TP, _TP, _i, _TP, _vp, _vp]),
    "b200_layernorm_fwd": (_i, [_TP, _vp, _vp, _f, _i, _TP, _vp, _vp, _vp]),
    "b200_layernorm_bwd": (_i, [_TP, _TP, _vp, _vp, _vp, _vp, _i, _TP, _vp, _vp, _vp, _vp]),
    "b200_batchnorm_fwd_train": (_i, [_TP, _vp, _vp, _f, _f, _i, _TP, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200_batchnorm_fwd_infer": (_i, [_TP, _vp, _vp, _f, _i, _vp, _vp, _TP, _vp]),
    "b200_batchnorm_bwd": (_i, [_TP, _TP, _vp, _vp, _vp, _vp, _i, _TP, _vp, _vp, _vp, _vp, _vp]),
    "b200_batchnorm_stats": (_i, [_TP, _TP, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "b200_batchnorm_fwd_apply": (_i, [_TP, _vp, _vp, _f, _f, _i, _TP, _vp, _vp, _vp, _vp, _vp, C.c_double, _vp]),
    "b200_batchnorm_bwd_apply": (_i, [_TP, _TP, _vp, _vp, _vp, _vp, _i, _TP, _vp, C.c_double, _vp]),
    "b200_resize_extent": (_i, [_i, _f]),
    "b200_resample_taps": (_i, [_i, _i, _i]),
    "b200_resample_plan": (_i, [_i, _i, _i, _vp, _vp, _i]),
    "b200_resample_plan_transpose": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _i]),
    "b200_resample2d": (_i, [_TP, _TP, _vp, _vp, _i, _vp, _vp, _i, _i, _vp]),
    "b200_resample_compact": (_i, [_i, _i, _vp, _vp]),
    "b200_resample_mode": (_i, [_i, _i, _vp]),
    "b200_resample2d_ex": (_i, [_TP, _TP, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp]),
    "b200_maxpool2_fwd": (_i, [_TP, _TP, _vp]),
    "b200_maxpool2_bwd": (_i, [_TP, _TP, _TP, _TP, _i, _vp]),
    "b200_clipadd_fwd": (_i, [_TP, _TP, _TP, _vp]),
    "b200_clipadd_bwd": (_i, [_TP, _TP, _TP, _TP, _vp]),
    "b200_sr_loss": (_i, [_TP, _TP, _i, _f, _f, _vp, _TP, _vp, _vp]),
    "b200_bce_dice_loss": (_i, [_TP, _TP, _f, _f, _f, _vp, _TP, _vp, _vp]),
    "b200_binary_confusion": (_i, [_TP, _TP, _f, _vp, _vp]),
    "b200_softmax_fwd": (_i, [_TP, _TP, _vp]),
    "b200_softmax_ce_loss": (_i, [_TP, _vp, _f, _vp, _TP, _vp, _vp]),
    "b200_adam_advance": (_i, [_vp, _vp, _vp]),
    "b200_adam_step": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "b200_loss_scale_apply": (_i, [_vp, _i, _sz, _vp, _vp]),
    "b200_loss_scale_check": (_i, [_vp, _sz, _vp, _vp]),
    "b200_loss_scale_update": (_i, [_vp, _f, _vp]),
    "b200_peer_alloc": (_i, [C.POINTER(C.c_void_p), _sz]),
    "b200_peer_free": (_i, [_vp]),
    "b200_peer_export": (_i, [_vp, _vp]),
    "b200_peer_open": (_i, [_vp, C.POINTER(C.c_void_p)]),
    "b200_peer_close": (_i, [_vp]),
    "b200_peer_signal": (_i, [_vp, _vp]),
    "b200_peer_wait": (_i, [_vp, _i, _vp, C.c_double, _vp]),
    "b200_peer_pull": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "b200_peer_gather_sum": (_i, [_vp, _vp, _vp, _vp, _i, _sz, _i, _vp]),
    "b200_cast": (_i, [_vp, _i, _vp, _i, _sz, _vp]),
    "b200_copy_tensor": (_i, [_TP, _TP, _vp]),
    "b200_scale_inplace": (_i, [_vp, _sz, _f, _vp]),
    "b200_cv_resize_taps": (_i, [_i, _i, _i]),
    "b200_cv_resize_plan": (_i, [_i, _i, _i, _vp, _vp, _i]),
    "b200_patch_extract": (_i, [_vp, _i, _i, _i, _vp, _TP, _vp]),
    "b200_gather2d": (_i, [_TP, _TP, _vp, _vp, _i, _vp, _vp, _i, _i, _vp]),
    "b200_copy_rows": (_i, [_vp, _vp, _vp, _vp, _i, C.c_longlong, _vp]),
    "b200_luma_pair": (_i, [_TP, _TP, _i, _vp, _vp, _vp, _vp]),
    "b200_ssim_planes": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp]),
    "b200_avgpool2_planes": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "b200_debug_umma_probe": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "b200_debug_umma_rate": (_i, [_i, _i, _i, _vp, _i, _vp]),
}

_lib = None


def load():
    """Load the shared library (once) and attach prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise B200Error(
            f"{LIB_PATH} is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for the hot path."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().b200_last_error().decode("utf-8", "replace")
        raise B200Error(f"{what or 'b200 call'} failed ({status}): {msg}")
