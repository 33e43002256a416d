"""Batched ECG signal preprocessing on the GPU, with the function names and arguments of the reference's
per-sample host code (dataset.py:76-95 = signal_model.py:203-224 = evaluation_signal.py:20-39):

    remove_baseline_drift(signal, window_size=200)
    lowpass_filter(signal, cutoff=0.05, fs=1.0, order=5)       # dataset.py;  evaluation_signal.py: (40, 250, 5)
    z_score_normalize(signal)
    preprocess_signal(raw_signal)                              # baseline removal -> low-pass

`signal` is a CUDA tensor [..., L] (float32 or float64); every row along the last dimension is one signal,
so a whole batch [B, L] or [B, 12, L] goes through ONE launch of ecgmm_signal_preprocess.  The arithmetic is
float64 like the reference; the result is float32 (what dataset.py:68 feeds the model).  No CPU fallback.
"""
from __future__ import annotations

import ctypes

import torch

from . import lib, ops


def butter_lowpass(order: int, wn: float):
    """(b, a, zi): scipy.signal.butter(order, wn, 'low') and lfilter_zi(b, a), designed by libecgmm on the host."""
    b = (ctypes.c_double * (order + 1))()
    a = (ctypes.c_double * (order + 1))()
    zi = (ctypes.c_double * max(order, 1))()
    lib.call("ecgmm_butter_lowpass", int(order), float(wn), b, a, zi)
    return list(b), list(a), list(zi)[:order]


def _run(signal: torch.Tensor, window: int, order: int, wn: float, zscore: bool, eps: float = 1e-8) -> torch.Tensor:
    if not isinstance(signal, torch.Tensor) or not signal.is_cuda:
        raise lib.EcgmmError("signal must be a CUDA tensor (no CPU fallback)")
    if signal.dtype not in (torch.float32, torch.float64):
        raise lib.EcgmmError(f"signal must be float32 or float64, got {signal.dtype}")
    if signal.dim() < 1:
        raise lib.EcgmmError("signal must have at least one dimension")
    x = signal.detach().contiguous()
    L = x.shape[-1]
    rows = x.numel() // L if L else 0
    y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    ws_bytes = ops._shape_query("ecgmm_signal_preprocess_workspace", rows, L, order)
    ws = torch.empty(max(ws_bytes, 8) // 8, dtype=torch.float64, device=x.device)
    lib.call("ecgmm_signal_preprocess", ops._ptr(x), int(x.dtype == torch.float64), ops._ptr(y), ops._ptr(ws),
             ws.numel() * 8, rows, L, int(window), int(order), float(wn), int(zscore), float(eps), ops._s())
    return y


def remove_baseline_drift(signal, window_size=200):
    return _run(signal, window_size, 0, 0.5, False)


def lowpass_filter(signal, cutoff=0.05, fs=1.0, order=5):
    return _run(signal, 0, order, cutoff / (0.5 * fs), False)


def z_score_normalize(signal):
    return _run(signal, 0, 0, 0.5, True)


def preprocess_signal(raw_signal, cutoff=0.05, fs=1.0, order=5, window_size=200, zscore=False):
    """dataset.py:91-95 for a whole batch: baseline removal, zero-phase Butterworth low-pass, optional z-score
    (the reference leaves it commented out), one launch."""
    return _run(raw_signal, window_size, order, cutoff / (0.5 * fs), zscore)
