"""Deterministic weights / inputs shared by oracle/gen_golden.py and the parity tests.

Weights are procedural (seeded) because a full state_dict is 47 MB; the golden files store
per-tensor checksums so a drift of the generator is detected instead of silently accepted.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def make_oracle(seed=7, dims=(256, 256, 256), signal_channels=1, clinical_features=24):
    """Oracle model with seeded init, then non-trivial BatchNorm affine + running statistics so
    that eval-mode parity actually exercises them."""
    from oracle.model import ECGMultimodalModel

    torch.manual_seed(seed)
    m = ECGMultimodalModel(2, dims, clinical_features, signal_channels)
    perturb_norm_layers(m, seed + 1)
    return m


def perturb_norm_layers(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                mod.weight.copy_(1 + 0.2 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
                mod.running_mean.copy_(0.1 * torch.randn(mod.running_mean.shape, generator=g))
                mod.running_var.copy_(0.75 + 0.5 * torch.rand(mod.running_var.shape, generator=g))
            elif isinstance(mod, torch.nn.LayerNorm):
                mod.weight.copy_(1 + 0.2 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
        if hasattr(m, "attention_fusion"):
            m.attention_fusion.weights.copy_(torch.tensor([1.2, 0.9, 0.7]))


def make_inputs(seed, B, H, W, L, F=24, signal_channels=1):
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(B, 3, H, W, generator=g).clamp(-1, 1)
    if signal_channels == 1:
        ecg = torch.randn(B, L, generator=g)
    else:
        ecg = torch.randn(B, signal_channels, L, generator=g)
    clinical = torch.randn(B, F, generator=g)
    labels = torch.randint(0, 2, (B,), generator=g)
    return image, ecg, clinical, labels


def make_varied_inputs(seed, B, H, W, L, F=24):
    """Inputs whose statistics differ from sample to sample (per-sample image brightness / contrast / a sinusoidal
    pattern, signal amplitude and drift, wide clinical features), so that the logits of different rows differ by much
    more than the rounding noise -- used by the label-argmax test over many rows."""
    g = torch.Generator().manual_seed(seed)
    mu = torch.rand(B, 1, 1, 1, generator=g) * 1.2 - 0.6
    sig = torch.rand(B, 1, 1, 1, generator=g) * 0.8 + 0.2
    freq = torch.rand(B, 1, 1, 1, generator=g) * 0.2 + 0.01
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    image = (torch.randn(B, 3, H, W, generator=g) * sig + mu + 0.3 * torch.sin(freq * xs)).clamp(-1, 1)
    amp = torch.rand(B, 1, generator=g) * 2 + 0.2
    ecg = torch.randn(B, L, generator=g) * amp + torch.linspace(-1, 1, L).view(1, L) * (torch.rand(B, 1, generator=g) - 0.5)
    clinical = torch.randn(B, F, generator=g) * 2.0
    labels = torch.randint(0, 2, (B,), generator=g)
    return image, ecg, clinical, labels


def state_checksums(sd):
    """name -> (sum, abs-sum) in float64; cheap drift detector for procedural weights."""
    out = {}
    for k, v in sd.items():
        v64 = v.double()
        out[k] = (float(v64.sum()), float(v64.abs().sum()))
    return out


def set_dropout(m, p):
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = p
