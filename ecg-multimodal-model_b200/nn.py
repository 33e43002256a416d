"""Loss modules with the constructors of the ones the reference's loops instantiate, computed by
the fused libecgmm loss kernel (loss and d(loss)/d(logits) in one launch).

  CrossEntropyLoss  <- nn.CrossEntropyLoss()  train.py:31, train_kfold.py:41 (mean reduction, 2 logits)
  FocalLoss         <- signal_model.py:91-106 (alpha=1, gamma=2, mean reduction)
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import lib, ops

F32 = torch.float32


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, focal, alpha, gamma, ignore_index=-100, flag=None):
        if not logits.is_cuda:
            raise lib.EcgmmError("loss inputs must be CUDA tensors (no CPU fallback)")
        if logits.dim() != 2:
            raise lib.EcgmmError(f"logits must be [B,C], got {tuple(logits.shape)}")
        z = logits.detach().to(F32).contiguous()
        y = labels.detach().to(torch.int64).contiguous()
        B, C = z.shape
        out = torch.empty(1 + B * C, dtype=F32, device=z.device)
        loss, dz = out[0:1], out[1:].view(B, C)
        lib.call("ecgmm_ce_loss", ops._ptr(z), ops._ptr(y), ops._ptr(loss), ops._ptr(dz), B, C, int(focal),
                 float(alpha), float(gamma), 1.0, int(ignore_index), ops._ptr(flag), ops._s())
        ctx.save_for_backward(dz)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        # d(total)/d(logits) = g * dz : one SGEMM call with K = 1 (g is a device scalar; no sync)
        B, C = dz.shape
        out = ops.sgemm(dz.reshape(B * C, 1), g.detach().to(F32).reshape(1, 1).contiguous(), B * C, 1, 1)
        return out.view(B, C), None, None, None, None, None, None


class _LabelCheck:
    """Out-of-range labels without a device synchronisation: the kernel raises a device flag (and makes the loss NaN);
    the flag is copied to pinned host memory behind the kernel and looked at when the NEXT loss is computed, or by
    check() on demand (torch raises a device-side assert at once; a training loop here hears about it one step later,
    next to a NaN loss)."""

    def _flag(self, device):
        st = self.__dict__.get("_label_state")
        if st is None or st[0].device != device:
            dev = torch.zeros(1, dtype=torch.int32, device=device)
            host = torch.zeros(1, dtype=torch.int32).pin_memory()
            st = self.__dict__["_label_state"] = (dev, host)
        elif int(st[1][0]) != 0:
            st[1].zero_()
            st[0].zero_()
            raise lib.EcgmmError(f"{type(self).__name__}: a label of an earlier batch was outside [0, C) "
                                 "(and not ignore_index); that batch's loss was NaN")
        return st[0]

    def _publish(self):
        dev, host = self.__dict__["_label_state"]
        if dev.device.type == "cuda" and torch.cuda.is_current_stream_capturing():
            return  # a captured step keeps the flag on the device (check() reads it); its loss is NaN all the same
        host.copy_(dev, non_blocking=True)

    def check(self):
        """Synchronise and raise if any batch so far carried an out-of-range label."""
        st = self.__dict__.get("_label_state")
        if st is not None:
            torch.cuda.current_stream(st[0].device).synchronize()
            if int(st[1][0]) != 0 or int(st[0].item()) != 0:
                st[1].zero_()
                st[0].zero_()
                raise lib.EcgmmError(f"{type(self).__name__}: a label was outside [0, C) (and not ignore_index)")


class CrossEntropyLoss(_LabelCheck, nn.Module):
    """Mean softmax cross entropy over [B,C] logits and int64 labels; rows labelled ignore_index are left out of the
    mean and get a zero gradient, as in torch."""

    def __init__(self, weight=None, reduction="mean", label_smoothing=0.0, ignore_index=-100):
        super().__init__()
        if weight is not None or reduction != "mean" or label_smoothing != 0.0:
            raise lib.EcgmmError("only the reference configuration nn.CrossEntropyLoss() is implemented")
        self.ignore_index = int(ignore_index)

    def forward(self, logits, labels):
        out = _LossFn.apply(logits, labels, 0, 1.0, 0.0, self.ignore_index, self._flag(logits.device))
        self._publish()
        return out


class FocalLoss(_LabelCheck, nn.Module):
    """signal_model.py:91-106: mean(alpha * (1 - pt)^gamma * ce), pt = exp(-ce)."""

    def __init__(self, alpha=1.0, gamma=2.0, reduction="mean"):
        super().__init__()
        if reduction != "mean":
            raise lib.EcgmmError("only reduction='mean' is implemented (the reference default)")
        self.alpha, self.gamma = float(alpha), float(gamma)

    def forward(self, inputs, targets):
        # F.cross_entropy's default ignore_index (-100) applies inside the reference's FocalLoss too
        out = _LossFn.apply(inputs, targets, 1, self.alpha, self.gamma, -100, self._flag(inputs.device))
        self._publish()
        return out
