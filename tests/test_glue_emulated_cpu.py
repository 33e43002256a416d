"""CPU: the Python glue of the attribution / serving rows with the kernels EMULATED in torch.

tests/test_control_flow_cpu.py checks call sequences with the launches stubbed out; here the entry points these rows
use are replaced by torch restatements of what include/ecgmm.h says each one computes (operating on the tensors whose
pointers the glue passes), so the glue's own arithmetic -- which operand goes where, which GEMM is transposed, how the
chunks are sliced -- is checked numerically against the oracle without a GPU.  The CUDA code itself is what the GPU
tests (tests/test_attrib_serve_gpu.py) check."""
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


def emu_modality_share(phi, share, S, C, D0, D1, D2, use_sum, stream):
    share.reshape(-1)[: S * C * 3].view(S, C, 3).copy_(
        om.modality_share(phi.view(S, D0 + D1 + D2, C), (D0, D1, D2), "sum" if use_sum else "mean"))


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
        with torch.no_grad():
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


# ---------------------------------------------------------------------------------------------- serving: block wiring
def emu_conv_weight_prep(w, w_fwd, w_dg, O, I, R, S, stream):
    w4 = w.reshape(O, I, R, S)
    if w_fwd is not None:
        w_fwd.copy_(w4.permute(0, 2, 3, 1).to(torch.bfloat16))
    if w_dg is not None:
        w_dg.copy_(w4.permute(1, 2, 3, 0).to(torch.bfloat16))


def _conv(x, w_fwd, stride, pH, pW):
    return torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w_fwd.float().permute(0, 3, 1, 2), None, stride,
                                      (pH, pW)).permute(0, 2, 3, 1)


def emu_conv2d_fwd(x, w, y, N, H, W, Cin, Cout, R, S, stride, pH, pW, stream):
    y.copy_(_conv(x, w, stride, pH, pW).to(torch.bfloat16))


def emu_conv2d_fwd_bn(x, w, y, scale, shift, res, relu, N, H, W, Cin, Cout, R, S, stride, pH, pW, stream):
    r = _conv(x, w, stride, pH, pW) * scale + shift
    if res is not None:
        r = r + res.float()
    y.copy_((torch.relu(r) if relu else r).to(torch.bfloat16))


def emu_bn_eval_coeffs(C, gamma, beta, conv_bias, mean, var, eps, scale, shift, stream):
    s = gamma / torch.sqrt(var + eps)
    scale.copy_(s)
    shift.copy_(beta - mean * s + (0 if conv_bias is None else conv_bias * s))


def emu_bn_apply(x, scale, shift, se, res, y, mask, N, P, C, relu, stream):
    assert se is None and mask is None
    r = x.float() * scale + shift
    if res is not None:
        r = r + res.float()
    y.copy_((torch.relu(r) if relu else r).to(torch.bfloat16))


BLOCK_EMULATORS = {"ecgmm_conv_weight_prep": emu_conv_weight_prep, "ecgmm_conv2d_fwd": emu_conv2d_fwd,
                   "ecgmm_conv2d_fwd_bn": emu_conv2d_fwd_bn, "ecgmm_bn_eval_coeffs": emu_bn_eval_coeffs,
                   "ecgmm_bn_apply": emu_bn_apply}


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("layer,block", [("layer1", 1), ("layer2", 0), ("layer4", 0)])
def test_serving_block_wiring(emulated, monkeypatch, fused, layer, block):
    """serve._block_eval (folded BatchNorm, identity or 1x1 projection, optional fused epilogue) against the oracle's
    torchvision BasicBlock in eval mode with non-trivial running statistics."""
    for k, v in BLOCK_EMULATORS.items():
        monkeypatch.setitem(EMULATORS, k, v)
    monkeypatch.setattr(serve, "FUSED_EPILOGUE", fused)
    ora = make_oracle(seed=7)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for m in ora.image_encoder.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.2)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.weight.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    dut = ecgmm.ECGMultimodalModel(Cfg)
    dut.load_state_dict(ora.state_dict())
    ora.eval()
    dut.eval()
    oblk, dblk = getattr(ora.image_encoder, layer)[block], getattr(dut.image_encoder, layer)[block]
    cin = oblk.conv1.in_channels
    x = torch.randn(2, 9, 14, cin, generator=g).to(torch.bfloat16)
    co = serve.fold_batchnorm(dut)
    out = serve._block_eval(dblk, x, co)
    with torch.no_grad():
        ref = oblk(x.float().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    assert out.shape == ref.shape and out.dtype == torch.bfloat16
    err = (out.float() - ref).abs().max().item()
    assert err <= 3e-2 * max(1.0, ref.abs().max().item()), err  # bf16 storage of the intermediate activations


# ---------------------------------------------------------------------------------------------- LIME surrogate glue
def emu_perturb_build(e, bg, masks, out, S, V, D, stream):
    z = _mat(masks, V, D).float()
    out.reshape(-1)[: S * V * D].view(S, V, D).copy_(
        (z.unsqueeze(0) * _mat(e, S, D).unsqueeze(1) + (1 - z).unsqueeze(0) * bg.reshape(1, 1, D)).to(torch.bfloat16))


def emu_head_tail(hidden, b1, w2, b2, out, rows, HID, C, cls, stream):
    h = torch.relu(_mat(hidden, rows, HID).float() + b1.reshape(1, HID))
    logits = h @ _mat(w2, C, HID).t() + b2.reshape(1, C)
    if cls < 0:
        _mat(out, rows, C).copy_(logits)
    else:
        out.reshape(-1)[:rows].copy_(torch.softmax(logits, 1)[:, cls])


def emu_f32_to_bf16(x, y, n, stream):
    y.reshape(-1)[:n].copy_(x.reshape(-1)[:n].to(torch.bfloat16))


def emu_perturb_pack_masks(masks, bits, V, D, stream):
    """Element 4k + b of each 32-element word on bit 8b + 7 - k; words laid out [D/64 chunks][V][2] (include/ecgmm.h)."""
    m = (_mat(masks, V, D) != 0).view(V, D // 32, 8, 4).long()  # [.., k, b]
    k, b = torch.arange(8).view(8, 1), torch.arange(4).view(1, 4)
    w = (m << (8 * b + 7 - k)).sum((-1, -2))  # [V, D/32]
    w = torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32)
    bits.reshape(-1)[: V * (D // 32)].view(D // 64, V, 2).copy_(w.view(V, D // 64, 2).permute(1, 0, 2))


def _unpack_masks(bits, V, D):
    w = bits.reshape(-1)[: V * (D // 32)].view(D // 64, V, 2).permute(1, 0, 2).reshape(V, D // 32).long() & 0xFFFFFFFF
    k, b = torch.arange(8).view(8, 1), torch.arange(4).view(1, 4)
    return ((w.view(V, D // 32, 1, 1) >> (8 * b + 7 - k)) & 1).view(V, D).bool()


def emu_perturb_head_fused(e, bg, bits, w1, b1, w2, b2, out, S, V, D, C, cls, stream):
    z = _unpack_masks(bits, V, D)
    x = torch.where(z.unsqueeze(0), _mat(e, S, D).unsqueeze(1), bg.reshape(1, 1, D))  # bf16: an exact selection
    h = torch.relu(x.float().view(S * V, D) @ _mat(w1, 128, D).float().t() + b1.reshape(1, 128))
    logits = h @ _mat(w2, C, 128).t() + b2.reshape(1, C)
    if cls < 0:
        _mat(out, S * V, C).copy_(logits)
    else:
        out.reshape(-1)[: S * V].copy_(torch.softmax(logits, 1)[:, cls])


def test_masked_regression_glue(emulated, monkeypatch):
    """explain.masked_regression: perturbation inference (bf16 operands of the first Linear, as on the device) and the
    [S, V] x [V, D+1] product with the REAL host-designed operator, against the fp32 oracle."""
    for k, v in {"ecgmm_perturb_build": emu_perturb_build, "ecgmm_conv2d_fwd": emu_conv2d_fwd,
                 "ecgmm_head_tail": emu_head_tail, "ecgmm_f32_to_bf16": emu_f32_to_bf16,
                 "ecgmm_perturb_pack_masks": emu_perturb_pack_masks,
                 "ecgmm_perturb_head_fused": emu_perturb_head_fused}.items():
        monkeypatch.setitem(EMULATORS, k, v)
    ora, dut = _pair()
    g = torch.Generator().manual_seed(8)
    S, V, D = 4, 200, 768
    e, bg = torch.randn(S, D, generator=g), torch.randn(D, generator=g)
    masks, w = explain.lime_plan(V, D, seed=2)
    # the operator design is a HOST function of libecgmm: call the real one, not an emulator
    import ctypes

    so = lib.load()
    R = torch.empty((D + 1, V), dtype=torch.float32)
    assert so.ecgmm_ridge_operator(ctypes.c_void_p(masks.data_ptr()), ctypes.cast(w.data_ptr(), ctypes.POINTER(ctypes.c_double)),
                                   V, D, 1.0, ctypes.c_void_p(R.data_ptr())) == 0
    coef, b = explain.masked_regression(dut.fusion_classifier, e, bg, masks, operator=R)
    cref, bref = om.masked_regression(ora.fusion_classifier, e, bg, masks, w, 1.0)
    # the responses differ from the fp32 oracle by the bf16 rounding of the variants / first-layer weights (~1e-3 on a
    # probability); the fit is a contraction of those differences
    assert coef.shape == (S, D) and b.shape == (S,)
    assert (coef - cref).abs().max().item() <= 2e-4 and (b - bref).abs().max().item() <= 5e-3
    sh = explain.modality_share(coef.unsqueeze(-1).contiguous(), reduce="sum")
    assert torch.allclose(sh, om.modality_share(coef.unsqueeze(-1), reduce="sum"), atol=1e-3)
