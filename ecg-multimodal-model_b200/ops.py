"""Thin tensor-level wrappers over the C ABI: allocate outputs with torch, pass raw pointers.

Activations are channels-last bf16 tensors shaped [N, H, W, C] (contiguous).  Nothing here
computes with torch operators; torch only owns the memory and the stream.
"""
from __future__ import annotations

import ctypes

import torch

from . import lib

BF16 = torch.bfloat16


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _chk(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise lib.EcgmmError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise lib.EcgmmError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise lib.EcgmmError(f"{name} must be contiguous")


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---------------------------------------------------------------- layout
def nchw_to_nhwc_bf16(x: torch.Tensor) -> torch.Tensor:
    _chk(x, torch.float32, "x")
    N, C, H, W = x.shape
    y = torch.empty((N, H, W, C), dtype=BF16, device=x.device)
    lib.call("ecgmm_nchw_f32_to_nhwc_bf16", _ptr(x), _ptr(y), N, C, H, W, _s())
    return y


def nhwc_to_nchw_f32(x: torch.Tensor) -> torch.Tensor:
    _chk(x, BF16, "x")
    N, H, W, C = x.shape
    y = torch.empty((N, C, H, W), dtype=torch.float32, device=x.device)
    lib.call("ecgmm_nhwc_bf16_to_nchw_f32", _ptr(x), _ptr(y), N, C, H, W, _s())
    return y


def conv_weight_prep(w: torch.Tensor, need_dgrad: bool = True):
    """fp32 OIHW (or OIK for 1-D) -> (w_fwd [O][R][S][I], w_dgrad [I][R][S][O]) bf16."""
    _chk(w, torch.float32, "w")
    if w.dim() == 3:
        O, I, S = w.shape
        R = 1
    else:
        O, I, R, S = w.shape
    w_fwd = torch.empty((O, R, S, I), dtype=BF16, device=w.device)
    w_dg = torch.empty((I, R, S, O), dtype=BF16, device=w.device) if need_dgrad else None
    lib.call("ecgmm_conv_weight_prep", _ptr(w), _ptr(w_fwd), _ptr(w_dg), O, I, R, S, _s())
    return w_fwd, w_dg


# ---------------------------------------------------------------- convolutions
def _conv_out(H, W, R, S, stride, pH, pW):
    return (H + 2 * pH - R) // stride + 1, (W + 2 * pW - S) // stride + 1


def conv2d_fwd(x: torch.Tensor, w_fwd: torch.Tensor, stride: int = 1) -> torch.Tensor:
    _chk(x, BF16, "x")
    _chk(w_fwd, BF16, "w_fwd")
    N, H, W, Cin = x.shape
    Cout, R, S, Cin2 = w_fwd.shape
    assert Cin == Cin2
    pH, pW = R // 2, S // 2
    Ho, Wo = _conv_out(H, W, R, S, stride, pH, pW)
    y = torch.empty((N, Ho, Wo, Cout), dtype=BF16, device=x.device)
    lib.call("ecgmm_conv2d_fwd", _ptr(x), _ptr(w_fwd), _ptr(y), N, H, W, Cin, Cout, R, S, stride, pH, pW, _s())
    return y


def conv2d_dgrad(dy: torch.Tensor, w_dgrad: torch.Tensor, in_hw, stride: int = 1, out: torch.Tensor = None,
                 accumulate: bool = False) -> torch.Tensor:
    _chk(dy, BF16, "dy")
    _chk(w_dgrad, BF16, "w_dgrad")
    N, Ho, Wo, Cout = dy.shape
    Cin, R, S, Cout2 = w_dgrad.shape
    assert Cout == Cout2
    H, W = in_hw
    pH, pW = R // 2, S // 2
    assert (Ho, Wo) == _conv_out(H, W, R, S, stride, pH, pW)
    if out is None:
        assert not accumulate
        out = torch.empty((N, H, W, Cin), dtype=BF16, device=dy.device)
    else:
        _chk(out, BF16, "out")
        assert tuple(out.shape) == (N, H, W, Cin)
    lib.call("ecgmm_conv2d_dgrad", _ptr(dy), _ptr(w_dgrad), _ptr(out), N, H, W, Cin, Cout, R, S, stride, pH, pW,
             int(accumulate), _s())
    return out


def conv2d_wgrad(x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor, R: int, S: int, stride: int = 1) -> None:
    """dw (fp32, [Cout, Cin, R, S] or [Cout, Cin, S]) += x^T * dy."""
    _chk(x, BF16, "x")
    _chk(dy, BF16, "dy")
    _chk(dw, torch.float32, "dw")
    N, H, W, Cin = x.shape
    Cout = dy.shape[3]
    assert dw.numel() == Cout * Cin * R * S
    lib.call("ecgmm_conv2d_wgrad", _ptr(x), _ptr(dy), _ptr(dw), N, H, W, Cin, Cout, R, S, stride, R // 2, S // 2,
             _s())


# ---------------------------------------------------------------- ResNet stem
def stem_s2d_dims(H: int, W: int):
    hs, ws = ctypes.c_int(0), ctypes.c_int(0)
    lib.load().ecgmm_stem_s2d_dims(H, W, ctypes.byref(hs), ctypes.byref(ws))
    return hs.value, ws.value


def stem_s2d(x: torch.Tensor) -> torch.Tensor:
    """NCHW image (fp32 or bf16, 3 channels) -> space-to-depth staging buffer [N][Hs][Ws][16] bf16."""
    if x.dtype not in (torch.float32, BF16):
        raise lib.EcgmmError(f"image must be fp32 or bf16, got {x.dtype}")
    _chk(x, x.dtype, "image")
    N, C, H, W = x.shape
    if C != 3:
        raise lib.EcgmmError(f"image must have 3 channels, got {C}")
    Hs, Ws = stem_s2d_dims(H, W)
    xs = torch.empty((N, Hs, Ws, 16), dtype=BF16, device=x.device)
    lib.call("ecgmm_stem_s2d", _ptr(x), int(x.dtype == BF16), _ptr(xs), N, H, W, _s())
    return xs


def stem_weight_prep(w: torch.Tensor) -> torch.Tensor:
    _chk(w, torch.float32, "w")
    assert tuple(w.shape) == (64, 3, 7, 7)
    ws = torch.empty((64, 256), dtype=BF16, device=w.device)
    lib.call("ecgmm_stem_weight_prep", _ptr(w), _ptr(ws), _s())
    return ws


def stem_conv_fwd(xs: torch.Tensor, w_s2d: torch.Tensor, H: int, W: int) -> torch.Tensor:
    _chk(xs, BF16, "xs")
    _chk(w_s2d, BF16, "w_s2d")
    N = xs.shape[0]
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((N, Ho, Wo, 64), dtype=BF16, device=xs.device)
    lib.call("ecgmm_stem_conv_fwd", _ptr(xs), _ptr(w_s2d), _ptr(y), N, H, W, _s())
    return y


def stem_conv_wgrad(xs: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor, H: int, W: int) -> None:
    _chk(xs, BF16, "xs")
    _chk(dy, BF16, "dy")
    _chk(dw, torch.float32, "dw")
    assert dw.numel() == 64 * 3 * 49
    lib.call("ecgmm_stem_conv_wgrad", _ptr(xs), _ptr(dy), _ptr(dw), xs.shape[0], H, W, _s())
