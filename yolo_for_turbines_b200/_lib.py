"""ctypes binding of libyolo_b200.so (the C-ABI declared in include/yolo_b200.h).

The library is built in-tree by `python -m yolo_for_turbines_b200.build` (or
`__graft_entry__.build()`).  There is deliberately NO fallback: if the shared object is missing
or a call fails, an exception is raised -- the product path never routes through torch ops or the
CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libyolo_b200.so")

YB_OK = 0
STATUS_NAN_INPUT = 1
STATUS_NAN_LAYER = 2
ACT_CODES = {None: 0, "none": 0, "leaky_relu": 1, "mish": 2}
BOX_CENTER, BOX_CORNERS = 0, 1


class YoloB200Error(RuntimeError):
    pass


class ConvDesc(C.Structure):
    """Mirror of `yolo_conv_desc` (include/yolo_b200.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "batch", "h_in", "w_in", "c_in", "in_pitch", "c_out", "c_out_pad", "out_pitch", "ksize", "stride", "pad",
        "act", "has_residual", "res_pitch", "upsample2x", "out_fp32", "check_nan", "a_mode", "block_n_hint",
        "stages_hint", "impl_hint", "cta_pair_hint", "ksize_w", "stride_w", "pad_w_hi_plus1", "stem_c", "want_stats", "pad_h_hi_plus1", "s2_parity", "s2_cin",
        "pdl_hint", "tail_split_hint", "row_hint", "decode_mode", "dec_nc", "dec_rows_per_image", "dec_row_offset")] + [
        ("dec_anchor_bits", C.c_int32 * 6), ("mc_hint", C.c_int32)]


class BnFinalizeDesc(C.Structure):
    """Mirror of `yolo_bn_finalize_desc` (include/yolo_b200.h)."""
    _fields_ = [("P", C.c_longlong), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float), ("momentum", C.c_float),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("mean", C.c_void_p), ("rstd", C.c_void_p),
                ("scale", C.c_void_p), ("bias", C.c_void_p), ("counter", C.c_void_p)]


_P, _I, _F, _D, _SZ, _LL = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t, C.c_longlong

# name -> (restype, argtypes); every int-returning entry is error-checked by _Lib.__getattr__
SIGNATURES = {
    "yolo_last_error": (C.c_char_p, []),
    "yolo_version": (_I, []),
    "yolo_device_info": (_I, [_I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "yolo_conv_plan_bytes": (_SZ, []),
    "yolo_conv_plan_init": (_I, [_P, _SZ, C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P]),
    "yolo_conv_fwd": (_I, [_P, _P, _P]),
    "yolo_conv_fwd_stats": (_I, [_P, _P, _P, C.POINTER(BnFinalizeDesc), _P]),
    "yolo_conv_fwd_stem": (_I, [_P, _P, _P, _P]),
    "yolo_conv_plan_info": (_I, [_P, C.POINTER(C.c_int32)]),
    "yolo_conv_max_clusters": (_I, [_I, C.POINTER(_I)]),
    "yolo_conv_fwd_trace": (_I, [_P, _P, _P, _I, _P]),
    "yolo_conv_fwd_simt": (_I, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P, _P]),
    "yolo_pack_weights": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "yolo_pack_stem_weights": (_I, [_P, _I, _I, _I, _P, _P]),
    "yolo_fold_bn": (_I, [_P, _P, _P, _P, _P, _F, _I, _I, _P, _P, _P]),
    "yolo_nchw_to_nhwc_bf16": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "yolo_nhwc_to_nchw_f32": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "yolo_input_patchify": (_I, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "yolo_letterbox_desc_bytes": (_SZ, []),
    "yolo_letterbox_u8": (_I, [_P, _I, _I, _I, _P, _P]),
    "yolo_decode": (_I, [_P, C.POINTER(C.c_int64), _I, _I, _I, C.POINTER(_F), _I, _I, _P, _I, _I, _P]),
    "yolo_decode_multi": (_I, [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), _I, C.POINTER(C.c_int32), _I, C.POINTER(_F), _I, _P, _I, _P]),
    "yolo_nms_workspace_bytes": (_SZ, [_I, _I]),
    "yolo_nms": (_I, [_P, _P, _I, _I, _F, _D, _I, _I, _P, _P, _P, _SZ, _P]),
    "yolo_iou": (_I, [_P, _I, _I, _P, _I, _I, _I, _I, _P, _P]),
    "yolo_map_match": (_I, [_P, _I, _P, _I, _P, _P, _P, _F, _I, _P, _P, _P, _P, _P]),
    "yolo_map_ap": (_I, [_P, _P, _P, _P, _I, _P, _P]),
    "yolo_accuracy_counts": (_I, [_P, C.POINTER(C.c_int64), _P, C.POINTER(C.c_int64), _I, _I, _I, _F, _P, _P]),
    "yolo_loss_fwd": (_I, [_P, C.POINTER(C.c_int64), _P, C.POINTER(C.c_int64), _I, _I, _I, C.POINTER(_F), _I, _P, _P]),
    "yolo_bn_stats": (_I, [_P, _LL, _I, _I, _P, _P]),
    "yolo_bn_finalize": (_I, [_P, _LL, _I, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P]),
    "yolo_bn_act_fwd": (_I, [_P, _LL, _I, _I, _P, _P, _I, _P, _I, _P, _I, _I, _I, _I, _P]),
    "yolo_bn_act_bwd": (_I, [_P, _I, _I, _P, _I, _LL, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _I, _P, _I, _P]),
    "yolo_bn_stats_finalize": (_I, [_P, _LL, _I, _I, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P]),
    "yolo_pack_weights_train": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "yolo_bias_grad": (_I, [_P, _LL, _I, _I, _I, _P, _P, _P]),
    "yolo_wgrad_plan_bytes": (_SZ, []),
    "yolo_wgrad_plan_init": (_I, [_P, _SZ, C.POINTER(ConvDesc), _P, _P, _I, _P, _I]),
    "yolo_wgrad": (_I, [_P, _P]),
    "yolo_wgrad_plan_info": (_I, [_P, C.POINTER(C.c_int32)]),
    "yolo_unpack_wgrad": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "yolo_pack_weights_dgrad": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "yolo_pack_weights_dgrad_s2": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "yolo_sgd_step": (_I, [_P, _P, _P, _LL, _F, _F, _F, _F, _I, _P]),
    "yolo_sgd_step_dev": (_I, [_P, _P, _P, _LL, _P, _F, _F, _F, _I, _P]),
    "yolo_loss_bwd": (_I, [_P, C.POINTER(C.c_int64), _P, C.POINTER(C.c_int64), _I, _I, _I, C.POINTER(_F), _P, _F,
                           C.POINTER(_F), _P, C.POINTER(C.c_int64), _I, _P]),
    "yolo_encode_targets": (_I, [_P, _P, _I, C.POINTER(_F), _I, _I, _I, _F, _P, _P, _P, _P]),
    "yolo_sort_workspace_bytes": (_SZ, [_I]),
    "yolo_sort_pairs": (_I, [_P, _P, _P, _I, _I, _P, _SZ, _P]),
}


class _Lib:
    def __init__(self):
        self._dll = None

    def _load(self):
        if self._dll is None:
            if not os.path.isfile(LIB_PATH):
                raise YoloB200Error(
                    f"{LIB_PATH} is missing: build it with `python -m yolo_for_turbines_b200.build` "
                    "(nvcc, sm_100a).  There is no CPU/PyTorch fallback for this path.")
            dll = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(dll, name)  # AttributeError => header and library disagree
                fn.restype, fn.argtypes = res, args
            self._dll = dll
        return self._dll

    def raw(self, name):
        return getattr(self._load(), name)

    def __getattr__(self, name):
        if name.startswith("_") or name not in SIGNATURES:
            raise AttributeError(name)
        fn = self.raw(name)
        if SIGNATURES[name][0] is not _I or name == "yolo_version":
            self.__dict__[name] = fn      # cached: later lookups never reach __getattr__
            return fn

        def checked(*args):
            rc = fn(*args)
            if rc != YB_OK:
                msg = self.raw("yolo_last_error")().decode(errors="replace")
                raise YoloB200Error(f"{name} failed ({rc}): {msg}")
            return rc

        checked.__name__ = name
        self.__dict__[name] = checked     # cached: later lookups never reach __getattr__
        return checked


lib = _Lib()


def ptr(t):
    """Device/host pointer of a tensor (or NULL)."""
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise YoloB200Error(
            f"{what} must be a CUDA tensor: this package is the sm_100a path only and has no CPU fallback "
            f"(got device {t.device}).")
