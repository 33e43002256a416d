"""GPU: expected gradients / modality shares and the image endpoint with Grad-CAM (rank 4)
against the oracle (first green on a B200 in round 2, gpurun_out/r02a_zz_tests.log; part of the default GPU suite)."""

import os

import pytest
import torch

from ecgmm import explain, lib, serve
from oracle import model as om
from parity_util import build_pair

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _eg_case(S, K, NB, seed):
    g = torch.Generator().manual_seed(seed)
    e = torch.randn(S, 768, generator=g)
    bg = torch.randn(NB, 768, generator=g)
    idx, alpha = explain.sampling_plan(S, K, NB, seed=seed + 1)
    return e, bg, idx, alpha


@pytest.mark.parametrize("S,K,NB", [(4, 50, 20), (1, 1, 1), (7, 33, 100)])
def test_expected_gradients_match_oracle(S, K, NB):
    """fp32 end to end: |phi - phi_ref| <= 2e-4 (phi ~ 5e-2; one hidden unit whose pre-activation changes sign
    between the two summation orders moves an entry by ~3e-5)."""
    ora, dut = build_pair(seed=7)
    e, bg, idx, alpha = _eg_case(S, K, NB, seed=100 + S)
    ref = om.expected_gradients(ora.fusion_classifier, e, bg, idx, alpha)
    phi = explain.expected_gradients(dut.fusion_classifier, e.to(DEV), bg.to(DEV), idx, alpha)
    assert phi.shape == ref.shape == (S, 768, 2) and phi.dtype == torch.float32
    assert (phi.cpu() - ref).abs().max().item() <= 2e-4
    # plan on the device, wrapper module, chunked evaluation: same numbers
    phi2 = explain.expected_gradients(om_wrapper(dut), e.to(DEV), bg.to(DEV), idx.to(DEV), alpha.to(DEV), chunk_samples=3)
    assert (phi2 - phi).abs().max().item() <= 1e-6
    share = explain.modality_share(phi)
    assert share.shape == (S, 2, 3)
    assert (share.cpu() - om.modality_share(phi.cpu())).abs().max().item() <= 1e-3
    assert (share.sum(-1) - 100.0).abs().max().item() <= 1e-3


def om_wrapper(dut):
    import ecgmm

    return ecgmm.FusionClassifierWrapper(dut.fusion_classifier)


def test_modality_share_edge_cases():
    phi = torch.zeros(3, 672, 2, device=DEV)
    assert float(explain.modality_share(phi, dims=(512, 128, 32)).abs().max()) == 0.0  # the all-zero guard
    phi[:, 512:640, 1] = -2.0  # only the signal slice of class 1 carries attribution
    sh = explain.modality_share(phi, dims=(512, 128, 32))
    assert torch.equal(sh[:, 1].cpu(), torch.tensor([[0.0, 100.0, 0.0]] * 3)) and float(sh[:, 0].abs().max()) == 0.0
    with pytest.raises(lib.EcgmmError):
        explain.modality_share(phi, dims=(256, 256, 256))


def _images(N, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    u8 = torch.randint(0, 256, (N, 3, H, W), generator=g, dtype=torch.uint8)
    return u8, (u8.float() / 255.0 - 0.5) / 0.5  # ToTensor + Normalize(0.5, 0.5), dataset.py:119-123


def _tail_oracle(ora, act_nchw, classes):
    """The oracle's Grad-CAM tail (autograd) on a GIVEN layer4 activation."""
    act = act_nchw.clone().requires_grad_(True)
    feat = ora.image_encoder.fc(act.mean((2, 3)))
    logits = ora.image_classifier(ora.image_norm(feat))
    (g,) = torch.autograd.grad(logits.gather(1, classes.view(-1, 1).long()).sum(), act)
    cam = torch.relu((g.mean((2, 3), keepdim=True) * act.detach()).sum(1))
    return torch.softmax(logits.detach(), 1), cam


@pytest.mark.parametrize("N,H,W", [(3, 64, 160), (2, 224, 224)])
def test_image_endpoint_and_gradcam(N, H, W):
    ora, dut = build_pair(seed=7)
    ora.eval()
    dut.eval()
    u8, img = _images(N, H, W, seed=N * 10 + 1)
    p_ref, cam_ref, _ = om.image_endpoint(ora, img, class_index=1)
    ep = serve.ImageEndpoint(dut, graph=False, class_index=1)
    probs, classes, cam = ep.gradcam(u8.to(DEV))
    assert classes.tolist() == [1] * N and cam.shape == cam_ref.shape and float(cam.min()) >= 0.0
    # (1) the whole chain against the fp32 oracle: bf16 activations through 20 convolutions
    assert (probs.cpu() - p_ref).abs().max().item() <= 4e-2
    num, den = (cam.cpu() - cam_ref).norm().item(), cam_ref.norm().item()
    assert den > 0 and num / den <= 0.15, (num, den)
    # (2) the tail (pool, fc, LayerNorm, classifier, the closed-form gradient, the channel contraction) on the
    # DEVICE's own layer4 activation: fp32 arithmetic on both sides
    act, *_ = serve.image_features(dut, u8.to(DEV))
    act_nchw = act.float().permute(0, 3, 1, 2).contiguous().cpu()
    p_tail, cam_tail = _tail_oracle(ora, act_nchw, torch.ones(N, dtype=torch.int64))
    assert (probs.cpu() - p_tail).abs().max().item() <= 1e-5
    assert (cam.cpu() - cam_tail).abs().max().item() <= 1e-6 + 1e-4 * cam_tail.max().item()
    # (3) argmax classes; eval mode has no batch statistics, so a sample's result does not depend on its neighbours
    ep2 = serve.ImageEndpoint(dut, graph=False)
    probs2, classes2 = ep2(u8.to(DEV))
    assert (probs2 - probs).abs().max().item() <= 1e-6
    assert classes2.tolist() == probs2.argmax(1).tolist()
    probs3, _ = ep2(u8.flip(0).contiguous().to(DEV))
    assert (probs3.flip(0) - probs2).abs().max().item() <= 1e-5
    dut.train()
    with pytest.raises(lib.EcgmmError):
        ep2(u8.to(DEV))


def test_image_endpoint_cuda_graph_replay():
    """One launch per request: the graphed endpoint returns what the eager one does, for new contents of the input
    buffer too, and notices new weights."""
    ora, dut = build_pair(seed=7)
    dut.eval()
    u8a, _ = _images(2, 64, 160, seed=5)
    u8b, _ = _images(2, 64, 160, seed=6)
    eager = serve.ImageEndpoint(dut, graph=False)
    ga = serve.ImageEndpoint(dut, example_image=u8a.to(DEV), graph=True)
    for u8 in (u8a, u8b, u8a):
        pe, ce, came = [t.clone() for t in eager.gradcam(u8.to(DEV))]
        pg, cg, camg = ga.gradcam(u8.to(DEV))
        torch.cuda.synchronize()
        assert (pg - pe).abs().max().item() <= 1e-6 and cg.tolist() == ce.tolist()
        assert (camg - came).abs().max().item() <= 1e-6 + 1e-5 * came.max().item()
        pg2, _ = ga(u8.to(DEV))
        assert (pg2 - pe).abs().max().item() <= 1e-6
    n0 = lib.launch_count()
    ga(u8b.to(DEV))
    assert lib.launch_count() == n0  # a replay issues no launch of its own from the host side of the library
    with torch.no_grad():
        dut.image_classifier.bias.add_(torch.tensor([0.5, -0.5], device=DEV))
    pe, _ = eager(u8b.to(DEV))
    pg, _ = ga(u8b.to(DEV))  # re-captured
    assert (pg - pe).abs().max().item() <= 1e-6
    with pytest.raises(lib.EcgmmError):
        ga(u8a[:1].to(DEV))


# ---------------------------------------------------------------------------------------------------------------------
# Fused folded-BatchNorm epilogue (ecgmm_conv2d_fwd_bn, the serving default): separate template instantiations of the
# tcgen05 kernels (the training instantiations' SASS is unchanged).
@pytest.mark.parametrize("case", [
    # N, H, W, Cin, Cout, R, stride, residual, relu          kernel
    (2, 16, 160, 64, 64, 3, 1, True, True),      # halo kernel, residual through the TMA prefetch
    (3, 5, 150, 64, 64, 3, 1, False, True),      # halo kernel, ragged width
    (2, 16, 40, 64, 128, 3, 2, False, True),     # generic N=128, stride 2
    (2, 16, 40, 64, 128, 1, 2, False, False),    # 1x1 projection, no activation
    (2, 8, 20, 128, 128, 3, 1, True, True),      # generic N=128 with residual
    (2, 4, 10, 256, 256, 3, 1, True, True),      # generic N=256
    (1, 2, 5, 512, 512, 3, 1, True, False),      # two N tiles
], ids=lambda c: "x".join(str(v) for v in c))
def test_conv_bn_epilogue(case):
    from ecgmm import ops

    N, H, W, Cin, Cout, R, stride, with_res, relu = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(N, H, W, Cin, generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, R, R, generator=g) * (2.0 / (Cin * R * R)) ** 0.5)
    scale = torch.rand(Cout, generator=g) + 0.5
    shift = torch.randn(Cout, generator=g) * 0.5
    Ho, Wo = (H + 2 * (R // 2) - R) // stride + 1, (W + 2 * (R // 2) - R) // stride + 1
    res = torch.randn(N, Ho, Wo, Cout, generator=g).to(torch.bfloat16) if with_res else None
    w_fwd = w.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)                  # [O][R][S][I]
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w_fwd.float().permute(0, 3, 1, 2), None, stride,
                                     R // 2).permute(0, 2, 3, 1)
    ref = ref * scale + shift
    if with_res:
        ref = ref + res.float()
    if relu:
        ref = torch.relu(ref)
    st = ops.BNStats(None, None, scale.to(DEV), shift.to(DEV))
    y = ops.conv2d_fwd_bn(x.to(DEV), w_fwd.to(DEV), st, stride, None if res is None else res.to(DEV), relu)
    assert tuple(y.shape) == (N, Ho, Wo, Cout)
    err = (y.float().cpu() - ref).abs()
    assert float((err / (1.0 + ref.abs())).max()) <= 1e-2  # one bf16 rounding of the result


def test_fused_endpoint_matches_unfused(monkeypatch):
    ora, dut = build_pair(seed=7)
    dut.eval()
    u8, _ = _images(3, 64, 160, seed=4)
    monkeypatch.setattr(serve, "FUSED_EPILOGUE", False)
    plain = [t.clone() for t in serve.ImageEndpoint(dut, graph=False, class_index=1).gradcam(u8.to(DEV))]
    n0 = lib.launch_count()
    serve.ImageEndpoint(dut, graph=False, class_index=1).gradcam(u8.to(DEV))
    n_plain = lib.launch_count() - n0
    monkeypatch.setattr(serve, "FUSED_EPILOGUE", True)
    ep = serve.ImageEndpoint(dut, graph=False, class_index=1)
    fused = ep.gradcam(u8.to(DEV))
    n0 = lib.launch_count()
    ep.gradcam(u8.to(DEV))
    assert lib.launch_count() - n0 < n_plain - 15  # no scale/shift passes (the first request also folds: 20 launches)
    assert (fused[0] - plain[0]).abs().max().item() <= 2e-2
    num, den = (fused[2] - plain[2]).norm().item(), plain[2].norm().item()
    assert num / den <= 0.1


def test_graphed_eval_step_matches_eager_forward():
    """The inference step of the fusion model (train.py:183-200) as one CUDA graph: same outputs as the eager eval
    forward, for new batch contents too; re-captured after the weights change; refuses train mode."""
    from ecgmm import graph as eg

    ora, dut = build_pair(seed=7)
    dut.eval()
    g = torch.Generator().manual_seed(3)

    def batch():
        return (torch.randn(4, 3, 64, 160, generator=g).clamp_(-1, 1).to(DEV), torch.randn(4, 600, generator=g).to(DEV),
                torch.randn(4, 24, generator=g).to(DEV))

    b1, b2 = batch(), batch()
    infer = eg.GraphedEvalStep(dut, b1)
    for b in (b1, b2, b1):
        with torch.no_grad():
            want = [t.clone() for t in dut(*b)]
        got = infer(*b)
        torch.cuda.synchronize()
        assert len(got) == 6
        for a, w in zip(got, want):
            assert (a - w).abs().max().item() <= 1e-6
    n0 = lib.launch_count()
    infer(*b2)
    assert lib.launch_count() == n0
    with torch.no_grad():
        dut.fusion_classifier.lin2.bias.add_(1.0)
        want = dut(*b2)[3].clone()
    assert (infer(*b2)[3] - want).abs().max().item() <= 1e-6
    dut.train()
    with torch.no_grad():
        dut.fusion_classifier.lin2.bias.add_(1.0)  # forces a re-capture, which must refuse the mode
    with pytest.raises(lib.EcgmmError):
        infer(*b2)


def test_attribution_and_endpoint_match_reference_golden():
    """Against tests/golden/attrib_g2.pt: attributions and Grad-CAM evaluated on the real reference model's own modules
    (oracle/gen_golden_attrib.py)."""
    import os

    from golden_util import GOLDEN_DIR

    gold = torch.load(os.path.join(GOLDEN_DIR, "attrib_g2.pt"))
    ora, dut = build_pair(seed=7)
    dut.eval()
    phi = explain.expected_gradients(dut.fusion_classifier, gold["e"].to(DEV), gold["bg"].to(DEV), gold["idx"],
                                     gold["alpha"])
    assert (phi.cpu() - gold["phi"]).abs().max().item() <= 2e-4
    assert (explain.modality_share(phi).cpu() - gold["share"]).abs().max().item() <= 0.5  # percent
    ep = serve.ImageEndpoint(dut, graph=False, class_index=gold["class_index"])
    probs, classes, cam = ep.gradcam(gold["u8"].to(DEV))
    assert (probs.cpu() - gold["probs"]).abs().max().item() <= 4e-2
    num, den = (cam.cpu() - gold["cam"]).norm().item(), gold["cam"].norm().item()
    assert num / den <= 0.15, (num, den)


def test_masked_regression_matches_oracle():
    """LIME-style local surrogate (sklearn-Ridge-equivalent fit on binary keep-masks): V model evaluations per sample on
    the tensor-core perturbation path + one SGEMM with the host-designed operator, against the fp32 oracle."""
    ora, dut = build_pair(seed=7)
    g = torch.Generator().manual_seed(12)
    S, V, D = 5, 300, 768
    e, bg = torch.randn(S, D, generator=g), torch.randn(D, generator=g)
    masks, w = explain.lime_plan(V, D, seed=4)
    cref, bref = om.masked_regression(ora.fusion_classifier, e, bg, masks, w, 1.0)
    coef, b = explain.masked_regression(dut.fusion_classifier, e.to(DEV), bg.to(DEV), masks, w, alpha=1.0)
    assert coef.shape == (S, D) and b.shape == (S,)
    assert (coef.cpu() - cref).abs().max().item() <= 2e-4 and (b.cpu() - bref).abs().max().item() <= 5e-3
    R = explain.regression_operator(masks, w, 1.0, torch.device(DEV))  # a ready operator, masks on the device
    coef2, b2 = explain.masked_regression(dut.fusion_classifier, e.to(DEV), bg.to(DEV), masks.to(DEV), operator=R)
    assert (coef2 - coef).abs().max().item() <= 1e-6 and (b2 - b).abs().max().item() <= 1e-6
    sh = explain.modality_share(coef.unsqueeze(-1).contiguous(), reduce="sum")
    assert (sh.cpu() - om.modality_share(coef.cpu().unsqueeze(-1), reduce="sum")).abs().max().item() <= 1e-2
