"""GPU: exact modality Shapley values -- ecgmm.explain.modality_shapley against the
oracle's enumeration (first green on a B200 in round 2, gpurun_out/r02a_zz_tests.log)."""

import pytest
import torch

from ecgmm import explain, lib
from oracle import model as om
from parity_util import build_pair

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _case(S, V, D=768, seed=0):
    g = torch.Generator().manual_seed(seed)
    e = torch.randn(S, D, generator=g)
    bg = torch.randn(100, D, generator=g).mean(0)
    masks = (torch.rand(V, D, generator=g) < 0.5).to(torch.uint8)
    return e, bg, masks


def test_modality_shapley_matches_oracle():
    """exact 3-player Shapley values of the modalities (8 coalitions through the
    tensor-core head + one small SGEMM) against the oracle's enumeration; efficiency holds on the device numbers."""
    ora, dut = build_pair(seed=7)
    e, bg, _ = _case(6, 8, seed=21)
    phi_ref, f0_ref, f1_ref = om.modality_shapley(ora.fusion_classifier, e, bg)
    phi, f0, f1 = explain.modality_shapley(dut.fusion_classifier, e.to(DEV), bg.to(DEV))
    assert phi.shape == (6, 3) and f0.shape == (6,) and f1.shape == (6,)
    assert (phi.cpu() - phi_ref).abs().max().item() <= 1e-2
    assert (f0.cpu() - f0_ref).abs().max().item() <= 1e-2 and (f1.cpu() - f1_ref).abs().max().item() <= 1e-2
    assert (phi.sum(1) - (f1 - f0)).abs().max().item() <= 1e-5
    with pytest.raises(lib.EcgmmError):
        explain.modality_shapley(dut.fusion_classifier, e.to(DEV), bg.to(DEV), dims=(256, 256, 128))
