"""CPU: the Python glue of the attribution / serving rows with the kernels EMULATED in torch.

tests/test_control_flow_cpu.py checks call sequences with the launches stubbed out; here the entry points these rows
use are replaced by torch restatements of what include/ecgmm.h says each one computes (operating on the tensors whose
pointers the glue passes), so the glue's own arithmetic -- which operand goes where, which GEMM is transposed, how the
chunks are sliced -- is checked numerically against the oracle without a GPU.  The CUDA code itself is what the GPU
tests (tests/test_zz_attrib_serve_gpu.py) check."""
import pytest
import torch

import ecgmm
from ecgmm import explain, lib, ops, serve
from oracle import model as om
from parity_util import make_oracle


def _mat(t, rows, cols):
    return t.reshape(-1)[: rows * cols].view(rows, cols)


def emu_sgemm(A, B, C, bias, M, N, K, ta, tb, accumulate, relu, stream):
    a = _mat(A, K, M).t() if ta else _mat(A, M, K)
    b = _mat(B, N, K).t() if tb else _mat(B, K, N)
    r = a @ b
    if bias is not None:
        r = r + bias.reshape(-1)[:N]
    out = _mat(C, M, N)
    if accumulate:
        r = r + out
    out.copy_(torch.relu(r) if relu else r)


def emu_eg_points(e, bg, idx, alpha, points, S, K, D, NB, stream):
    b = _mat(bg, NB, D)[idx.reshape(-1)[: S * K].long()].view(S, K, D)
    x = _mat(e, S, D).unsqueeze(1)
    points.reshape(-1)[: S * K * D].view(S, K, D).copy_(b + alpha.reshape(-1)[: S * K].view(S, K, 1) * (x - b))


def emu_eg_gate(hidden, w2, gate, rows, HID, C, stream):
    h = _mat(hidden, rows, HID)
    gate.reshape(-1)[: C * rows * HID].view(C, rows, HID).copy_((h > 0).float().unsqueeze(0) * _mat(w2, C, HID).unsqueeze(1))


def emu_eg_reduce(e, bg, idx, grad, phi, S, K, D, C, NB, stream):
    b = _mat(bg, NB, D)[idx.reshape(-1)[: S * K].long()].view(S, K, D)
    diff = _mat(e, S, D).unsqueeze(1) - b
    g = grad.reshape(-1)[: C * S * K * D].view(C, S, K, D)
    phi.reshape(-1)[: S * D * C].view(S, D, C).copy_((diff.unsqueeze(0) * g).mean(2).permute(1, 2, 0))


def emu_modality_share(phi, share, S, C, D0, D1, D2, stream):
    share.reshape(-1)[: S * C * 3].view(S, C, 3).copy_(om.modality_share(phi.view(S, D0 + D1 + D2, C), (D0, D1, D2)))


def emu_gather_rows(table, idx, out, rows, D, NT, stream):
    _mat(out, rows, D).copy_(_mat(table, NT, D)[idx.reshape(-1)[:rows].long()])


def emu_layernorm_bwd(x, dy, gamma, mean, rstd, dx, dgamma, dbeta, rows, D, accumulate_dx, stream):
    xh = (_mat(x, rows, D) - mean.reshape(-1)[:rows, None]) * rstd.reshape(-1)[:rows, None]
    dxh = _mat(dy, rows, D) * gamma.reshape(1, D)
    r = rstd.reshape(-1)[:rows, None] * (dxh - dxh.mean(1, keepdim=True) - xh * (dxh * xh).mean(1, keepdim=True))
    if dx is not None:
        _mat(dx, rows, D).copy_(r + _mat(dx, rows, D) if accumulate_dx else r)
    assert dgamma is None and dbeta is None  # not needed by the rows under test


def emu_gradcam(act, g, cam, N, P, C, scale, stream):
    a = act.reshape(-1)[: N * P * C].view(N, P, C).float()
    _mat(cam, N, P).copy_(torch.relu(scale * (a * _mat(g, N, C).unsqueeze(1)).sum(-1)))


EMULATORS = {"ecgmm_sgemm": emu_sgemm, "ecgmm_eg_points": emu_eg_points, "ecgmm_eg_gate": emu_eg_gate,
             "ecgmm_eg_reduce": emu_eg_reduce, "ecgmm_modality_share": emu_modality_share,
             "ecgmm_gather_rows": emu_gather_rows, "ecgmm_layernorm_bwd": emu_layernorm_bwd,
             "ecgmm_gradcam": emu_gradcam}


@pytest.fixture
def emulated(monkeypatch):
    def call(name, *args):
        assert len(args) == len(lib.SIGNATURES[name]), name
        EMULATORS[name](*args)

    monkeypatch.setattr(lib, "call", call)
    monkeypatch.setattr(ops, "_ptr", lambda t: t)  # the emulators take the tensors themselves
    monkeypatch.setattr(ops, "_s", lambda: 0)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True), raising=False)


class Cfg:
    num_classes = 2
    device = "cpu"


def _pair():
    ora = make_oracle(seed=7)
    dut = ecgmm.ECGMultimodalModel(Cfg)
    dut.load_state_dict(ora.state_dict())
    return ora, dut


@pytest.mark.parametrize("S,K,NB,chunk", [(5, 9, 7, 0), (5, 9, 7, 2), (1, 1, 1, 0)])
def test_expected_gradients_glue(emulated, S, K, NB, chunk):
    ora, dut = _pair()
    g = torch.Generator().manual_seed(S * 31 + K)
    e, bg = torch.randn(S, 768, generator=g), torch.randn(NB, 768, generator=g)
    idx, alpha = explain.sampling_plan(S, K, NB, seed=3)
    ref = om.expected_gradients(ora.fusion_classifier, e, bg, idx, alpha)
    phi = explain.expected_gradients(dut.fusion_classifier, e, bg, idx, alpha, chunk_samples=chunk)
    assert phi.shape == ref.shape and torch.allclose(phi, ref, atol=1e-6)
    share = explain.modality_share(phi)
    assert torch.allclose(share, om.modality_share(ref), atol=1e-3)


def test_gradcam_tail_glue(emulated):
    """serve.gradcam_from_features (row gather -> LayerNorm backward -> fc^T SGEMM -> channel contraction) against the
    oracle's autograd Grad-CAM on the same layer4 activation."""
    ora, dut = _pair()
    ora.eval()
    g = torch.Generator().manual_seed(11)
    act = torch.relu(torch.randn(3, 2, 5, 512, generator=g)).to(torch.bfloat16)       # NHWC, as the device holds it
    classes = torch.tensor([1, 0, 1], dtype=torch.int32)
    a = act.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    feat = ora.image_encoder.fc(a.mean((2, 3)))
    logits = ora.image_classifier(ora.image_norm(feat))
    (grad,) = torch.autograd.grad(logits.gather(1, classes.view(-1, 1).long()).sum(), a)
    want = torch.relu((grad.mean((2, 3), keepdim=True) * a.detach()).sum(1))
    f = feat.detach()
    mean = f.mean(1)
    rstd = (f.var(1, unbiased=False) + dut.image_norm.eps).rsqrt()
    cam = serve.gradcam_from_features(dut, act, f.contiguous(), mean, rstd, classes)
    assert cam.shape == (3, 2, 5) and torch.allclose(cam, want, atol=1e-6 + 1e-4 * float(want.max()))
