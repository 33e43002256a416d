"""ctypes binding of libecgmm.so (C ABI declared in include/ecgmm.h).

The library is the only compute path of this package: if it is missing or the device is not
sm_100 every operator raises -- there is no CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_int, c_void_p, c_float, c_double, c_longlong, c_ulonglong, c_char_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libecgmm.so")

_p = c_void_p
_i = c_int
_f = c_float
_ll = c_longlong
_d = c_double

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "ecgmm_version": [],
    "ecgmm_last_error": [],
    "ecgmm_check_device": [],
    "ecgmm_launch_count": [],
    "ecgmm_nchw_f32_to_nhwc_bf16": [_p, _p, _i, _i, _i, _i, _p],
    "ecgmm_nhwc_bf16_to_nchw_f32": [_p, _p, _i, _i, _i, _i, _p],
    "ecgmm_conv_weight_prep": [_p, _p, _p, _i, _i, _i, _i, _p],
    "ecgmm_conv_weight_prep_batch": [_p, _i, _p],
    "ecgmm_stem_s2d_dims": [_i, _i, POINTER(c_int), POINTER(c_int)],
    "ecgmm_stem_s2d": [_p, _i, _p, _i, _i, _i, _p],
    "ecgmm_stem_weight_prep": [_p, _p, _p],
    "ecgmm_stem_conv_fwd": [_p, _p, _p, _i, _i, _i, _p],
    "ecgmm_stem_conv_fwd_stats_rows": [_i, _i, _i],
    "ecgmm_stem_conv_fwd_stats": [_p, _p, _p, _p, _p, _i, _i, _i, _p],
    "ecgmm_stem_conv_wgrad_workspace": [_i, _i, _i],
    "ecgmm_stem_conv_wgrad": [_p, _p, _p, _i, _i, _i, _p, _ll, _p],
    "ecgmm_conv2d_fwd": [_p, _p, _p] + [_i] * 10 + [_p],
    "ecgmm_conv2d_fwd_stats_rows": [_i] * 10,
    "ecgmm_conv2d_fwd_stats": [_p, _p, _p, _p, _p] + [_i] * 10 + [_p],
    "ecgmm_conv2d_fwd_bn": [_p, _p, _p, _p, _p, _p, _i] + [_i] * 10 + [_p],
    "ecgmm_conv2d_dgrad": [_p, _p, _p] + [_i] * 11 + [_p],
    "ecgmm_conv2d_dgrad_reduce_rows": [_i] * 10,
    "ecgmm_conv2d_dgrad_reduce": [_p] * 9 + [_i] * 11 + [_p],
    "ecgmm_conv2d_wgrad": [_p, _p, _p] + [_i] * 10 + [_p, _ll, _p],
    "ecgmm_conv2d_wgrad_workspace": [_i] * 10,
    # BatchNorm / ReLU / pooling
    "ecgmm_reduce_split": [_i, _i, _i],
    "ecgmm_chan_stats": [_p, _p, _p, _i, _i, _i, _i, _p],
    "ecgmm_bn_finalize": [_p, _p, _i, _i, _i, _ll, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "ecgmm_bn_eval_coeffs": [_i, _p, _p, _p, _p, _p, _f, _p, _p, _p],
    "ecgmm_bn_apply": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "ecgmm_bn_relu_maxpool": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "ecgmm_bn_bwd_reduce": [_p] * 10 + [_i] * 6 + [_p],
    "ecgmm_bn_bwd_finalize": [_p, _p, _i, _i, _i, _ll] + [_p] * 11 + [_ll, _p],
    "ecgmm_bn_bwd_apply": [_p] * 13 + [_i] * 5 + [_p],
    "ecgmm_avgpool_fwd": [_p, _p, _i, _i, _i, _p],
    "ecgmm_avgpool_bwd": [_p, _p, _i, _i, _i, _p],
    # 1-D ResNet-SE specifics
    "ecgmm_signal_stem_fwd": [_p, _p, _p, _i, _i, _i, _p],
    "ecgmm_signal_s4d_len": [_i],
    "ecgmm_signal_s4d": [_p, _p, _i, _i, _i, _p],
    "ecgmm_signal_stem_w4": [_p, _p, _i, _p],
    "ecgmm_signal_stem_dw4_fold": [_p, _p, _i, _p],
    "ecgmm_signal_stem_wgrad_workspace": [_i, _i, _i],
    "ecgmm_signal_stem_wgrad": [_p, _p, _p, _i, _i, _i, _p, _ll, _p],
    "ecgmm_se_fwd": [_p] * 10 + [_i] * 4 + [_p],
    "ecgmm_se_bwd": [_p, _p, _i] + [_p] * 9 + [_i] * 4 + [_p],
    # dense tails / fusion head / losses
    "ecgmm_sgemm": [_p, _p, _p, _p] + [_i] * 7 + [_p],
    "ecgmm_colsum": [_p, _p, _i, _i, _i, _p],
    "ecgmm_layernorm_fwd": [_p] * 6 + [_i, _i, _f, _p],
    "ecgmm_layernorm_bwd": [_p] * 8 + [_i, _i, _i, _p],
    "ecgmm_fusion_gate_fwd": [_p] * 6 + [_i] * 4 + [_p],
    "ecgmm_fusion_gate_bwd": [_p] * 9 + [_i] * 5 + [_p],
    "ecgmm_var_loss_fwd": [_p] * 6 + [_i] * 4 + [_p],
    "ecgmm_var_loss_bwd": [_p] * 5 + [_i] * 3 + [_p],
    "ecgmm_ce_loss": [_p, _p, _p, _p, _i, _i, _i, _f, _f, _f, _ll, _p, _p],
    "ecgmm_dropout_fwd": [_p, _p, _p, _p, _ll, _f, c_ulonglong, _p, _p],
    "ecgmm_mask_bwd": [_p, _p, _p, _p, _ll, _p],
    "ecgmm_zscore": [_p, _p, _ll, _i, _f, _p],
    "ecgmm_butter_lowpass": [_i, _d, POINTER(c_double), POINTER(c_double), POINTER(c_double)],
    "ecgmm_signal_preprocess_workspace": [_ll, _i, _i],
    "ecgmm_signal_preprocess": [_p, _i, _p, _p, _ll, _ll, _i, _i, _i, _d, _i, _d, _p],
    "ecgmm_perturb_build": [_p, _p, _p, _p, _ll, _i, _i, _p],
    "ecgmm_head_tail": [_p, _p, _p, _p, _p, _ll, _i, _i, _i, _p],
    "ecgmm_perturb_head_fused_supported": [_i, _i, _i],
    "ecgmm_perturb_pack_masks": [_p, _p, _i, _i, _p],
    "ecgmm_perturb_head_fused": [_p, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _i, _i, _i, _p],
    "ecgmm_f32_to_bf16": [_p, _p, _ll, _p],
    "ecgmm_eg_points": [_p, _p, _p, _p, _p, _ll, _i, _i, _i, _p],
    "ecgmm_eg_gate": [_p, _p, _p, _ll, _i, _i, _p],
    "ecgmm_eg_reduce": [_p, _p, _p, _p, _p, _ll, _i, _i, _i, _i, _p],
    "ecgmm_modality_share": [_p, _p, _ll, _i, _i, _i, _i, _i, _p],
    "ecgmm_ridge_operator": [_p, _p, _i, _i, _d, _p],
    "ecgmm_softmax_rows": [_p, _p, _p, _ll, _i, _p],
    "ecgmm_gather_rows": [_p, _p, _p, _ll, _i, _i, _p],
    "ecgmm_gradcam": [_p, _p, _p, _i, _i, _i, _f, _p],
    "ecgmm_bn_rows_fwd": [_p] * 9 + [_i, _i, _f, _f, _i, _i, _p],
    "ecgmm_bn_rows_bwd": [_p] * 9 + [_i, _i, _i, _p],
    # optimizer
    "ecgmm_adam_chunk_bytes": [],
    "ecgmm_adam_step": [_p, _i, _f, _f, _f, _f, _f, _ll, _f, _p],
    "ecgmm_step_advance": [_p, _p],
    "ecgmm_adam_step_dev": [_p, _i, _p, _f, _f, _f, _f, _p, _f, _p],
}
_RESTYPES = {"ecgmm_last_error": c_char_p, "ecgmm_stem_s2d_dims": None, "ecgmm_launch_count": c_ulonglong,
             "ecgmm_conv2d_wgrad_workspace": c_longlong, "ecgmm_signal_preprocess_workspace": c_longlong,
             "ecgmm_stem_conv_wgrad_workspace": c_longlong, "ecgmm_signal_stem_wgrad_workspace": c_longlong}


class WeightPrepDesc(ctypes.Structure):
    """ecgmm_weight_prep_desc (include/ecgmm.h)."""
    _fields_ = [("w", _p), ("w_fwd", _p), ("w_dgrad", _p), ("O", _i), ("I", _i), ("R", _i), ("S", _i)]


class EcgmmError(RuntimeError):
    """Raised when a libecgmm entry point returns a negative status."""


_lib = None
CALLS = 0  # C-ABI calls issued so far (each launches at least one kernel); bench.py reports the delta


def load() -> ctypes.CDLL:
    """dlopen libecgmm.so and attach the prototypes; fails loudly when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C ecg-multimodal-model_b200/csrc`). There is no fallback path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


def last_error() -> str:
    return load().ecgmm_last_error().decode("utf-8", "replace")


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point and raise EcgmmError on failure."""
    global CALLS
    CALLS += 1
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise EcgmmError(f"{name} failed with status {rc}: {last_error()}")


def launch_count() -> int:
    """Kernels launched by libecgmm in this process so far."""
    return int(load().ecgmm_launch_count())


def stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def require_device() -> None:
    """Raise unless a CUDA device of compute capability 10.x is current."""
    import torch

    if not torch.cuda.is_available():
        raise EcgmmError("ecgmm needs a CUDA device (sm_100); no CPU fallback exists")
    call("ecgmm_check_device")
