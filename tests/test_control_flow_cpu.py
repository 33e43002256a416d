"""CPU: the Python glue of the product path, executed end to end with the kernel launches stubbed out.

Every libecgmm entry point that launches a kernel goes through ecgmm.lib.call; here it is replaced by a recorder, and
tensors are made to claim `is_cuda`, so forward / backward / optimizer / explainers run their real control flow on
CPU memory (values are garbage, shapes and call sequences are real).  This catches what a GPU-less build cannot see
otherwise: wrong argument counts against the ctypes prototypes, missing gradients, stale caches, a code path that only
the GPU tests would reach."""
import ctypes

import pytest
import torch

import ecgmm
from ecgmm import explain, lib, ops, preprocess
from ecgmm import nn as enn
from ecgmm import optim as eoptim


class Cfg:
    num_classes = 2
    device = "cpu"


@pytest.fixture
def stub(monkeypatch):
    calls = []

    def fake_call(name, *args):
        argtypes = lib.SIGNATURES[name]
        assert len(args) == len(argtypes), f"{name}: {len(args)} arguments for {len(argtypes)} parameters"
        for a, t in zip(args, argtypes):  # what ctypes would accept
            if t is ctypes.c_void_p or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
                assert a is None or isinstance(a, (int, ctypes.c_void_p, ctypes.Array)), (name, type(a))
            elif t in (ctypes.c_int, ctypes.c_longlong, ctypes.c_ulonglong):
                assert isinstance(a, int) and not isinstance(a, bool) or isinstance(a, bool), (name, a)
            else:
                assert isinstance(a, (int, float)), (name, a)
        calls.append(name)

    monkeypatch.setattr(lib, "call", fake_call)
    monkeypatch.setattr(ops, "_s", lambda: 0)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True), raising=False)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self, raising=False)
    import types

    monkeypatch.setattr(torch.cuda, "cudart", lambda: types.SimpleNamespace(cudaHostRegister=lambda *a: 0))
    return calls


def _batch(B=2, H=64, W=160, L=600):
    g = torch.Generator().manual_seed(0)
    return (torch.randn(B, 3, H, W, generator=g), torch.randn(B, L, generator=g), torch.randn(B, 24, generator=g),
            torch.tensor([0, 1] * (B // 2)))


def test_train_step_flow(stub):
    m = ecgmm.ECGMultimodalModel(Cfg)
    m.overlap_branches = False  # the side-stream variant needs real CUDA streams
    m.train()
    crit, opt = enn.CrossEntropyLoss(), eoptim.Adam(m.parameters(), lr=1e-3)
    image, ecg, clin, labels = _batch()
    per_step = []
    for _ in range(2):
        n0 = len(stub)
        opt.zero_grad()
        out = m(image, ecg, clin)
        assert [tuple(o.shape) for o in out] == [(2, 2)] * 4 + [(), (3,)]
        (crit(out[3], labels) + 0.1 * out[4]).backward()
        opt.step()
        per_step.append(len(stub) - n0)
    assert per_step[0] > 300 and per_step[1] <= per_step[0]  # step 2 re-uses weight-shadow bookkeeping
    assert all(p.grad is not None for n, p in m.named_parameters()
               if not n.startswith(("image_classifier", "signal_classifier", "clinical_classifier")))
    assert all(int(opt.state[p]["step"]) == 2 for p in m.parameters() if p.grad is not None)
    names = set(stub)
    for needed in ("ecgmm_stem_s2d", "ecgmm_stem_conv_fwd_stats", "ecgmm_conv2d_fwd", "ecgmm_conv2d_fwd_stats",
                   "ecgmm_conv2d_dgrad", "ecgmm_conv2d_wgrad", "ecgmm_stem_conv_wgrad", "ecgmm_chan_stats",
                   "ecgmm_bn_finalize", "ecgmm_bn_apply", "ecgmm_bn_relu_maxpool", "ecgmm_bn_bwd_reduce",
                   "ecgmm_bn_bwd_finalize", "ecgmm_bn_bwd_apply", "ecgmm_se_fwd", "ecgmm_se_bwd", "ecgmm_ce_loss",
                   "ecgmm_adam_step", "ecgmm_fusion_gate_fwd", "ecgmm_var_loss_fwd", "ecgmm_dropout_fwd"):
        assert needed in names, needed


def test_eval_uint8_freeze_and_fusion_only_flows(stub):
    image, ecg, clin, labels = _batch()
    m = ecgmm.ECGMultimodalModel(Cfg)
    m.overlap_branches = False
    m.eval()
    with torch.no_grad():
        out = m((image * 100).to(torch.uint8), ecg, clin)  # raw pixels
    assert out[3].shape == (2, 2)
    # frozen statistics everywhere; only the three SE blocks of the signal encoder still need their per-sample sums
    assert "ecgmm_bn_eval_coeffs" in stub and "ecgmm_conv2d_fwd_stats" not in stub and stub.count("ecgmm_chan_stats") == 3
    # train.py:35-40 freeze mode: only the head trains
    m.train()
    for enc in (m.image_encoder, m.signal_encoder, m.clinical_encoder):
        for p in enc.parameters():
            p.requires_grad = False
    del stub[:]
    out = m(image, ecg, clin)
    (enn.CrossEntropyLoss()(out[3], labels) + 0.1 * out[4]).backward()
    assert "ecgmm_conv2d_wgrad" not in stub and "ecgmm_conv2d_dgrad" not in stub
    assert m.fusion_classifier.lin1.weight.grad is not None and m.image_encoder.conv1.weight.grad is None
    # single-tensor API of train_kfold.py:59-64 and the 512/128/32 layout of multimodal.py
    f = ecgmm.ECGMultimodalModel(Cfg, fusion_only=True, dims=(512, 128, 32))
    f.overlap_branches = False
    y = f(image, ecg, clin)
    assert isinstance(y, torch.Tensor) and y.shape == (2, 2)


def test_signal_model_and_helpers_flow(stub):
    net = ecgmm.ResNet1D_SE(12, 2).train()
    x = torch.randn(4, 12, 1000)
    opt = eoptim.Adam(net.parameters(), lr=1e-3)
    loss = enn.FocalLoss()(net(x), torch.tensor([0, 1, 1, 0]))
    loss.backward()
    opt.step()
    # L = 1000 (0 mod 4): the stem runs on the tensor cores through the regrouped signal
    assert "ecgmm_signal_s4d" in stub and "ecgmm_signal_stem_w4" in stub and "ecgmm_signal_stem_dw4_fold" in stub
    assert "ecgmm_signal_stem_fwd" not in stub
    assert all(p.grad is not None for p in net.parameters())
    del stub[:]
    enn.FocalLoss()(net(torch.randn(4, 12, 1001)), torch.tensor([0, 1, 1, 0])).backward()   # 1 mod 4: direct kernels
    assert "ecgmm_signal_stem_fwd" in stub and "ecgmm_signal_stem_wgrad" in stub and "ecgmm_signal_s4d" not in stub
    # preprocessing / explainers: argument marshalling of the newer entry points
    y = preprocess.preprocess_signal(torch.randn(3, 12, 500), zscore=True)
    assert y.shape == (3, 12, 500) and y.dtype == torch.float32 and stub[-1] == "ecgmm_signal_preprocess"
    m = ecgmm.ECGMultimodalModel(Cfg)
    e, bg = torch.randn(5, 768), torch.randn(768)
    p = explain.perturbation_inference(m.fusion_classifier, e, bg, (torch.rand(16, 768) < 0.5).to(torch.uint8))
    assert p.shape == (5, 16)
    assert stub[-4:] == ["ecgmm_f32_to_bf16", "ecgmm_f32_to_bf16", "ecgmm_perturb_pack_masks", "ecgmm_perturb_head_fused"]
    # widths the fused kernel does not cover (> 768) take the three-kernel path; 672 = 3 x 224 (G3) is zero-padded to 704
    from ecgmm.model import MLPHead

    p = explain.perturbation_inference(MLPHead(672, 128, 2), torch.randn(5, 672), torch.randn(672),
                                       (torch.rand(16, 672) < 0.5).to(torch.uint8))
    assert p.shape == (5, 16) and stub[-1] == "ecgmm_perturb_head_fused"
    p = explain.perturbation_inference(MLPHead(1024, 128, 2), torch.randn(5, 1024), torch.randn(1024),
                                       (torch.rand(16, 1024) < 0.5).to(torch.uint8))
    assert p.shape == (5, 16)
    assert stub[-3:] == ["ecgmm_perturb_build", "ecgmm_conv2d_fwd", "ecgmm_head_tail"]  # the uncovered-shape path
    phi, f0, f1 = explain.modality_shapley(m.fusion_classifier, e, bg)
    assert phi.shape == (5, 3) and f0.shape == (5,) and f1.shape == (5,) and stub[-1] == "ecgmm_sgemm"


def test_graph_mode_marshalling(stub, monkeypatch):
    """What ecgmm.graph switches on while it captures a step: Adam reads lr / step count from device words through
    pre-pinned chunk tables, dropout adds the device seed offset."""
    m = ecgmm.ECGMultimodalModel(Cfg)
    m.overlap_branches = False
    m.train()
    crit, opt = enn.CrossEntropyLoss(), eoptim.Adam(m.parameters(), lr=1e-3)
    image, ecg, clin, labels = _batch()
    state = torch.zeros(2, dtype=torch.int64)
    lr_dev = torch.tensor([1e-3])
    opt.reserve_tables()
    monkeypatch.setattr(ops, "GRAPH_STATE", state)
    opt._graph_mode = (state, lr_dev)
    seen = []
    real = lib.call

    def spy(name, *args):
        if name in ("ecgmm_dropout_fwd", "ecgmm_adam_step_dev"):
            seen.append((name, args))
        real(name, *args)

    monkeypatch.setattr(lib, "call", spy)
    out = m(image, ecg, clin)
    (crit(out[3], labels) + 0.1 * out[4]).backward()
    opt.step()
    opt._graph_mode = None
    drops = [a for n, a in seen if n == "ecgmm_dropout_fwd"]
    assert drops and all(a[7] == state.data_ptr() + 8 for a in drops)  # seed_dev -> second word of the state
    adams = [a for n, a in seen if n == "ecgmm_adam_step_dev"]
    assert len(adams) == 1 and adams[0][2] == lr_dev.data_ptr() and adams[0][7] == state.data_ptr()
    assert "ecgmm_adam_step" not in stub
    host, dev = opt._reserved[0]
    assert adams[0][1] <= host.shape[0] and int(adams[0][0].value) == dev.data_ptr()


def test_attribution_and_serving_flows(stub):
    """argument marshalling and call sequences of expected gradients, the modality
    shares, the image endpoint and its Grad-CAM (eager; the graph variant needs real CUDA streams)."""
    from ecgmm import serve

    m = ecgmm.ECGMultimodalModel(Cfg)
    e, bg = torch.randn(5, 768), torch.randn(12, 768)
    idx, alpha = explain.sampling_plan(5, 7, 12, seed=1)
    assert idx.shape == (5, 7) and idx.dtype == torch.int32 and float(alpha.min()) >= 0 and float(alpha.max()) < 1
    del stub[:]
    phi = explain.expected_gradients(m.fusion_classifier, e, bg, idx, alpha)
    assert phi.shape == (5, 768, 2)
    assert stub == ["ecgmm_eg_points", "ecgmm_sgemm", "ecgmm_eg_gate", "ecgmm_sgemm", "ecgmm_eg_reduce"]
    del stub[:]
    phi = explain.expected_gradients(ecgmm.FusionClassifierWrapper(m.fusion_classifier), e, bg, idx, alpha,
                                     chunk_samples=2)
    assert stub.count("ecgmm_eg_reduce") == 3  # 2 + 2 + 1 samples
    sh = explain.modality_share(phi)
    assert sh.shape == (5, 2, 3) and stub[-1] == "ecgmm_modality_share"
    with pytest.raises(lib.EcgmmError):
        explain.expected_gradients(m.fusion_classifier, e, bg, idx + 12, alpha)  # background row out of range
    with pytest.raises(lib.EcgmmError):
        explain.expected_gradients(m.fusion_classifier, e, bg, idx[:, :3], alpha)
    with pytest.raises(lib.EcgmmError):
        explain.expected_gradients(m.fusion_classifier, e, bg, idx.float(), alpha)
    with pytest.raises(lib.EcgmmError):
        explain.modality_share(phi, dims=(256, 256, 128))
    # LIME-style surrogate: the plan's operator is designed by a host function of the library (not a launch); every
    # sample then costs the perturbation path plus one SGEMM
    masks, w = explain.lime_plan(40, 768, seed=3)
    del stub[:]
    coef, icpt = explain.masked_regression(m.fusion_classifier, e, bg[0], masks, w, alpha=1.0)
    assert coef.shape == (5, 768) and icpt.shape == (5,)
    assert stub == ["ecgmm_ridge_operator", "ecgmm_f32_to_bf16", "ecgmm_f32_to_bf16", "ecgmm_perturb_pack_masks",
                    "ecgmm_perturb_head_fused", "ecgmm_sgemm"]
    explain.modality_share(coef.unsqueeze(-1).contiguous(), reduce="sum")
    with pytest.raises(lib.EcgmmError):
        explain.masked_regression(m.fusion_classifier, e, bg[0], masks)  # neither weights nor operator
    with pytest.raises(lib.EcgmmError):
        explain.modality_share(phi, reduce="max")

    image = (torch.rand(2, 3, 64, 160) * 255).to(torch.uint8)
    ep = serve.ImageEndpoint(m, example_image=image, graph=False)
    with pytest.raises(lib.EcgmmError):
        ep(image)  # the model is still in train mode
    m.eval()
    del stub[:]
    probs, classes = ep(image)
    assert probs.shape == (2, 2) and classes.shape == (2,) and classes.dtype == torch.int32
    assert stub[-1] == "ecgmm_softmax_rows" and "ecgmm_chan_stats" not in stub and "ecgmm_conv2d_fwd_stats" not in stub
    assert stub.count("ecgmm_bn_eval_coeffs") == 20  # folded once per endpoint ...
    del stub[:]
    ep(image)
    n_classify = len(stub)
    assert "ecgmm_bn_eval_coeffs" not in stub and "ecgmm_conv_weight_prep" not in stub  # ... not per request
    ep_cam = serve.ImageEndpoint(m, graph=False, class_index=1)
    ep_cam.gradcam(image)
    del stub[:]
    probs, classes, cam = ep_cam.gradcam(image)
    assert cam.shape == (2, 2, 5) and cam.dtype == torch.float32
    assert stub[-4:] == ["ecgmm_gather_rows", "ecgmm_layernorm_bwd", "ecgmm_sgemm", "ecgmm_gradcam"]
    assert len(stub) == n_classify + 4  # Grad-CAM costs four small launches on top of a classification
    with torch.no_grad():
        m.image_encoder.bn1.running_mean.add_(1.0)  # new statistics are picked up: folded again
    ep_cam.gradcam(image)
    assert stub.count("ecgmm_bn_eval_coeffs") == 20
    with pytest.raises(lib.EcgmmError):
        serve.ImageEndpoint(m, graph=False, class_index=5).gradcam(image)


def test_fused_serving_flow(stub, monkeypatch):
    """ECGMM_SERVE_FUSED=1: every Conv -> BatchNorm (-> += identity) (-> ReLU) of the 8 residual blocks is one
    ecgmm_conv2d_fwd_bn call: 19 convolutions, no scale/shift pass."""
    from ecgmm import serve

    monkeypatch.setattr(serve, "FUSED_EPILOGUE", True)
    m = ecgmm.ECGMultimodalModel(Cfg).eval()
    image = (torch.rand(2, 3, 64, 160) * 255).to(torch.uint8)
    ep = serve.ImageEndpoint(m, graph=False)
    ep(image)
    del stub[:]
    probs, classes = ep(image)
    assert probs.shape == (2, 2)
    assert stub.count("ecgmm_conv2d_fwd_bn") == 19 and "ecgmm_bn_apply" not in stub and "ecgmm_conv2d_fwd" not in stub
    assert len(stub) == 27  # s2d, stem conv, bn+relu+maxpool, 19 convs, avgpool, fc, LayerNorm, classifier, softmax
