"""ctypes binding of the C-ABI CUDA library (include/semivl_b200.h).

There is no fallback: importing this module raises if the library has not been built
(`python -m semivl_b200.build`), and every call raises SvlError on a non-zero status.
Tensors are passed as raw device pointers; torch only owns memory and streams.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libsemivl_b200.so")

F32, BF16, BF16X2 = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_RELU, ACT_GELU_DSAVE, ACT_SAVED = 0, 1, 2, 3, 4
OUT_LINEAR, OUT_CONVT2X2 = 0, 1
MAX_TAPS = 32


class SvlError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is missing: build the CUDA extension first (python -m semivl_b200.build); "
                      "semivl_b200 has no CPU / PyTorch fallback path")
_lib = C.CDLL(LIB_PATH)

_TapArr = C.c_int * MAX_TAPS


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_conv", C.c_int), ("m", C.c_int64), ("lda", C.c_int64), ("a_cols", C.c_int64),
        ("nb", C.c_int), ("h", C.c_int), ("w", C.c_int), ("a_map_w", C.c_int),
        ("b", C.c_void_p), ("b_rows", C.c_int64), ("ldb", C.c_int64),
        ("n", C.c_int), ("k_per_tap", C.c_int), ("num_taps", C.c_int),
        ("tap_dy", _TapArr), ("tap_dx", _TapArr), ("tap_a_koff", _TapArr), ("tap_b_row", _TapArr), ("tap_b_col", _TapArr),
        ("out", C.c_void_p), ("out_dtype", C.c_int), ("ldc", C.c_int64), ("out_mode", C.c_int), ("out_h", C.c_int), ("out_w", C.c_int),
        ("alpha", C.c_float), ("bias", C.c_void_p), ("row_bias", C.c_void_p), ("row_bias_div", C.c_int64), ("row_bias_ld", C.c_int64),
        ("act", C.c_int), ("preact_out", C.c_void_p), ("preact_dtype", C.c_int), ("ld_preact", C.c_int64),
        ("dact_src", C.c_void_p), ("dact_dtype", C.c_int), ("dact_kind", C.c_int), ("ld_dact", C.c_int64),
        ("residual", C.c_void_p), ("res_dtype", C.c_int), ("ldres", C.c_int64),
        ("accumulate", C.c_int), ("block_n", C.c_int), ("gn_part", C.c_void_p),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("dy", C.c_void_p), ("ld_dy", C.c_int64), ("dy_cols", C.c_int64),
        ("x", C.c_void_p), ("ld_x", C.c_int64), ("x_cols", C.c_int64),
        ("conv", C.c_int), ("rows", C.c_int64), ("nb", C.c_int), ("h", C.c_int), ("w", C.c_int),
        ("m", C.c_int), ("n", C.c_int), ("num_taps", C.c_int),
        ("tap_dy", _TapArr), ("tap_dx", _TapArr), ("tap_dy_koff", _TapArr), ("tap_x_koff", _TapArr), ("tap_slot", _TapArr), ("x_map_w", C.c_int),
        ("dw", C.c_void_p), ("ld_dw", C.c_int64), ("slot_stride", C.c_int64),
        ("alpha", C.c_float), ("splits", C.c_int),
    ]


_lib.svl_last_error.restype = C.c_char_p
_lib.svl_version.restype = C.c_int
_lib.svl_check_device.restype = C.c_int
_lib.svl_gemm.restype = C.c_int
_lib.svl_gemm.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
_lib.svl_conv_gn_splits.restype = C.c_int
_lib.svl_conv_gn_splits.argtypes = [C.POINTER(GemmDesc)]

_lib.svl_attention_bwd_workspace.restype = C.c_size_t
_lib.svl_attention_bwd_workspace.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
_lib.svl_gn_workspace.restype = C.c_size_t
_lib.svl_gn_workspace.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int]
_lib.svl_bn_workspace.restype = C.c_size_t
_lib.svl_bn_workspace.argtypes = [C.c_int64, C.c_int]
_lib.svl_wgrad.restype = C.c_int
_lib.svl_wgrad.argtypes = [C.POINTER(WgradDesc), C.c_void_p]

_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
_PROTOS = {
    "svl_patchify": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "svl_assemble_tokens": [_P, _P, _P, _P, _I, _I, _I, _P],
    "svl_pos_resize_fwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    "svl_pos_resize_bwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    "svl_layernorm_fwd": [_P, _L, _P, _P, _P, _I, _L, _P, _P, _L, _I, _F, _P],
    "svl_layernorm_bwd": [_P, _I, _L, _P, _L, _P, _P, _P, _P, _P, _P, _P, _I, _L, _P, _P, _L, _I, _P],
    "svl_l2norm_fwd": [_P, _L, _P, _P, _I, _L, _P, _L, _I, _F, _P],
    "svl_l2norm_bwd": [_P, _I, _L, _P, _P, _P, _L, _I, _L, _I, _P],
    "svl_cast": [_P, _I, _L, _L, _P, _I, _L, _L, _I, _L, _I, _F, _P],
    "svl_param_jobs": [_P, _P, _I, _I, _I, _P],
    "svl_colsum": [_P, _I, _L, _L, _I, _P, _P],
    "svl_batch_sum": [_P, _P, _I, _L, _I, _P],
    "svl_axpy": [_P, _P, _F, _L, _P],
    "svl_gn_relu_fwd": [_P, _I, _L, _P, _P, _P, _I, _L, _P, _I, _L, _P, _P, _P, _L, _I, _I, _I, _F, _I, _P],
    "svl_gn_relu_bwd": [_P, _I, _L, _P, _I, _L, _P, _P, _P, _P, _P, _I, _L, _P, _P, _P, _L, _I, _I, _I, _P],
    "svl_stem_im2col": [_P, _P, _I, _L, _I, _I, _I, _I, _I, _P],
    "svl_maxpool3s2_fwd": [_P, _I, _L, _P, _I, _L, _P, _I, _I, _I, _I, _I, _I, _P],
    "svl_maxpool3s2_bwd": [_P, _I, _L, _P, _P, _I, _L, _I, _I, _I, _I, _I, _I, _P],
    "svl_bn_stats": [_P, _I, _L, _L, _I, _P, _P, _P],
    "svl_bn_finalize": [_P, _F, _F, _F, _P, _P, _P, _P, _I, _P],
    "svl_bn_apply": [_P, _I, _L, _P, _P, _P, _P, _P, _I, _L, _P, _I, _L, _I, _L, _I, _P],
    "svl_bn_bwd_stats": [_P, _I, _L, _P, _I, _L, _P, _I, _L, _P, _P, _L, _I, _P, _P, _P],
    "svl_bn_bwd_apply": [_P, _I, _L, _P, _I, _L, _P, _I, _L, _P, _P, _P, _P, _F, _P, _I, _L, _P, _I, _L, _L, _I, _P],
    "svl_sim_im2col": [_P, _L, _P, _I, _L, _I, _I, _I, _I, _I, _I, _P],
    "svl_sim_col2im": [_P, _I, _L, _P, _I, _L, _I, _I, _I, _I, _I, _I, _P],
    "svl_map_sum": [_P, _I, _L, _P, _L, _I, _I, _F, _P],
    "svl_map_bcast_add": [_P, _P, _I, _L, _L, _I, _I, _F, _P],
    "svl_pool_tokens": [_P, _I, _L, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "svl_pool_tokens_bwd": [_P, _L, _P, _I, _I, _I, _I, _I, _I, _P],
    "svl_pool_tokens_bwd_from": [_P, _I, _L, _P, _L, _P, _I, _I, _I, _I, _I, _I, _P],
    "svl_unpool_add": [_P, _I, _L, _P, _L, _P, _I, _L, _I, _I, _I, _I, _I, _I, _I, _P],
    "svl_unpool_bwd": [_P, _I, _L, _P, _L, _I, _I, _I, _I, _I, _I, _I, _P],
    "svl_skip_fill": [_P, _I, _L, _P, _I, _L, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "svl_class_sum": [_P, _I, _L, _I, _P, _I, _I, _L, _I, _P],
    "svl_skip_grad": [_P, _I, _L, _I, _P, _I, _L, _P, _I, _L, _I, _I, _I, _I, _I, _I, _I, _P],
    "svl_conv_out1_fwd": [_P, _I, _L, _P, _P, _P, _L, _I, _I, _I, _P],
    "svl_conv_out1_bwd": [_P, _P, _I, _L, _P, _P, _I, _L, _P, _P, _L, _I, _I, _I, _P],
    "svl_upsample_bilinear": [_P, _P, _L, _I, _I, _I, _I, _P],
    "svl_upsample_bilinear_bwd": [_P, _P, _L, _I, _I, _I, _I, _P],
    "svl_resize_bilinear_ac": [_P, _P, _L, _I, _I, _I, _I, _P],
    "svl_softmax_max": [_P, _P, _P, _L, _I, _I, _I, _I, _I, _F, _F, _P],
    "svl_group_max": [_P, _L, _P, _P, _L, _I, _I, _P],
    "svl_upsample_ce": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _F, _I, _P],
    "svl_count_valid": [_P, _L, _I, _P, _P],
    "svl_reciprocal": [_P, _P, _F, _F, _P],
    "svl_cutmix_weights": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _F, _P],
    "svl_conf_stats": [_P, _P, _P, _P, _P, _P, _I, _L, _F, _P],
    "svl_conf_coef": [_P, _I, _I, _F, _P, _P, _P],
    "svl_fill_rows": [_P, _P, _I, _L, _P],
    "svl_cutmix_img": [_P, _P, _P, _P, _I, _I, _L, _P],
    "svl_adamw": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _F, _P],
    "svl_adamw_dev": [_P, _P, _P, _P, _L, _P, _I, _F, _F, _F, _F, _F, _P],
    "svl_window_accumulate": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "svl_divide_count": [_P, _P, _I, _I, _L, _P],
    "svl_argmax_classes": [_P, _P, _I, _I, _L, _P],
    "svl_intersection_union": [_P, _P, _L, _I, _I, _P, _P],
    "svl_crop_flip_normalize": [_P, _I, _I, _P, _I, _I, _I, _I, C.POINTER(C.c_float), C.POINTER(C.c_float), _P],
    "svl_crop_flip_mask": [_P, _I, _I, _P, _P, _I, _I, _I, _I, _I, _P],
    "svl_cutmix_box": [_P, _I, _I, _I, _I, _I, _P],
    "svl_resample_pass_u8": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "svl_gather_nearest_u8": [_P, _P, _P, _P, _I, _I, _I, _P],
    "svl_color_op_u8": [_P, _L, _I, _F, _I, _P, _P],
    "svl_box_blur_pass_u8": [_P, _P, _I, _I, _I, _I, _I, C.c_uint, C.c_uint, _P],
    "svl_attention_fwd": [_P, _I, _P, _P, _I, _I, _I, _F, _P],
    "svl_attention_bwd": [_P, _P, _P, _I, _P, _P, _P, _I, _L, _P, _I, _I, _I, _F, _P],
}
for _name, _args in _PROTOS.items():
    _fn = getattr(_lib, _name)
    _fn.restype = C.c_int
    _fn.argtypes = _args


def call(name, *args, n_launch=1):
    """Call an svl_* entry point; tensors are passed as data pointers, the current stream is appended."""
    global launches
    conv = [a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args]
    check(getattr(_lib, name)(*conv, torch.cuda.current_stream().cuda_stream), name)
    launches += n_launch


launches = 0     # number of kernels launched through this binding (bench.py reports it as gpu_launches)


def check(rc, what=""):
    if rc != 0:
        raise SvlError(f"{what}: status {rc}: {_lib.svl_last_error().decode()}")


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def version():
    return _lib.svl_version()


def check_device():
    check(_lib.svl_check_device(), "svl_check_device")


def dtype_of(t, split=False):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16X2 if split else BF16
    raise SvlError(f"unsupported dtype {t.dtype}")


def raw_gemm(desc):
    global launches
    check(_lib.svl_gemm(C.byref(desc), stream()), "svl_gemm")
    launches += 1


def raw_wgrad(desc):
    global launches
    check(_lib.svl_wgrad(C.byref(desc), stream()), "svl_wgrad")
    launches += 1


def lib():
    return _lib
