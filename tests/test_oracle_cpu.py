"""CPU: the oracle (oracle/model.py) against the committed golden vectors, which
oracle/gen_golden.py produced while asserting bit-identity with the real reference module."""
import os

import pytest
import torch

from golden_util import GOLDEN_DIR, make_inputs, make_oracle, set_dropout, state_checksums
from oracle import model as om

torch.set_num_threads(os.cpu_count() or 4)


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(GOLDEN_DIR, "fusion_g2.pt"), map_location="cpu", weights_only=False)


def test_ptbxl_known_answer():
    """The reference's only real checkpoint (best_ptbxl.pth) through the oracle's ResNet1D_SE."""
    kat = torch.load(os.path.join(GOLDEN_DIR, "ptbxl_kat.pt"), map_location="cpu", weights_only=False)
    sd = torch.load(os.path.join(GOLDEN_DIR, "best_ptbxl.pth"), map_location="cpu")
    net = om.ResNet1D_SE(1, 2)
    net.load_state_dict(sd, strict=True)
    net.eval()
    x = torch.randn(4, 1, 2476, generator=torch.Generator().manual_seed(kat["input_seed"]))
    with torch.no_grad():
        out = net(x)
    assert torch.allclose(out, kat["logits"], atol=1e-5)
    expect = torch.tensor([[3.9425, 0.3780], [4.0321, 0.0800], [3.9976, 0.2021], [3.8703, 0.1938]])
    assert torch.allclose(out, expect, atol=1e-3)
    assert out.argmax(1).tolist() == [0, 0, 0, 0]


def test_procedural_weights_match_golden_checksums(golden):
    sd = make_oracle(seed=7).state_dict()
    assert len(sd) == 229
    cs = state_checksums(sd)
    for k, (s, a) in golden["checksums"].items():
        assert abs(cs[k][0] - s) <= 1e-6 * max(1.0, abs(a)), k
        assert abs(cs[k][1] - a) <= 1e-6 * max(1.0, abs(a)), k


def test_oracle_matches_golden_small(golden):
    case = golden["cases"]["small"]
    B, H, W, L = case["shape"]
    m = make_oracle(seed=7)
    inputs = make_inputs(case["input_seed"], B, H, W, L)
    m.eval()
    with torch.no_grad():
        out = m(*inputs[:3])
    for a, b in zip(out, case["eval"]["outputs"]):
        assert torch.allclose(a, b, atol=1e-5, rtol=1e-5)
    set_dropout(m, 0.0)
    m.train()
    out = m(*inputs[:3])
    loss = om.fusion_loss(out, inputs[3])
    loss.backward()
    tp = case["train_p0"]
    assert abs(float(loss) - float(tp["loss"])) < 1e-5
    for k, g in tp["grads"].items():
        got = dict(m.named_parameters())[k].grad
        assert torch.allclose(got, g, atol=1e-6 + 1e-4 * float(g.abs().max())), k
    for k, v in tp["bn_after"].items():
        assert torch.allclose(m.state_dict()[k].float(), v.float(), atol=1e-5), k


def test_focal_loss_and_zscore_definitions():
    z = torch.tensor([[2.0, -1.0], [0.3, 0.1]])
    y = torch.tensor([0, 1])
    ce = torch.nn.functional.cross_entropy(z, y, reduction="none")
    want = ((1 - torch.exp(-ce)) ** 2 * ce).mean()
    assert torch.allclose(om.FocalLoss()(z, y), want)
    x = torch.tensor([[1.0, 2.0, 3.0, 6.0]])
    zs = om.z_score(x)
    assert abs(float(zs.mean())) < 1e-6 and abs(float(zs.var(unbiased=False)) - 1) < 1e-5


def test_signal12_golden(golden):
    s12 = golden["signal12"]
    torch.manual_seed(s12["init_seed"])
    net = om.ResNet1D_SE(12, 2)
    x = torch.randn(4, 12, 5000, generator=torch.Generator().manual_seed(s12["input_seed"]))
    net.train()
    set_dropout(net, 0.0)
    lo = net(x)
    assert torch.allclose(lo, s12["logits"], atol=1e-5)
    assert abs(float(om.FocalLoss()(lo, s12["labels"])) - float(s12["focal_loss"])) < 1e-6


def test_focal_loss_matches_reference_golden():
    """tests/golden/focal_kat.pt: written by oracle/gen_golden_focal.py while asserting bit-identity of the oracle's
    FocalLoss with signal_model.FocalLoss (signal_model.py:91-106), value and gradient."""
    kat = torch.load(os.path.join(GOLDEN_DIR, "focal_kat.pt"), map_location="cpu", weights_only=False)
    assert len(kat["cases"]) == 3
    for c in kat["cases"]:
        z = c["logits"].clone().requires_grad_(True)
        loss = om.FocalLoss()(z, c["labels"])
        loss.backward()
        assert torch.equal(loss.detach(), c["loss"]) and torch.equal(z.grad, c["grad"])


def test_perturbation_inference_matches_reference_golden():
    """tests/golden/perturb_g2.pt: the reference's own fusion_classifier (through its FusionClassifierWrapper),
    evaluated row by row on the masked variants by oracle/gen_golden_perturb.py; the oracle's batched evaluation
    agrees to the last bits of an fp32 GEMM."""
    kat = torch.load(os.path.join(GOLDEN_DIR, "perturb_g2.pt"), map_location="cpu", weights_only=False)
    ora = make_oracle(seed=kat["weights_seed"])
    logits = om.perturbation_inference(ora.fusion_classifier, kat["e"], kat["background"], kat["masks"], -1)
    prob = om.perturbation_inference(ora.fusion_classifier, kat["e"], kat["background"], kat["masks"], 1)
    assert logits.shape == kat["logits"].shape == (3, 64, 2)
    assert float((logits - kat["logits"]).abs().max()) < 1e-6
    assert float((prob - kat["prob1"]).abs().max()) < 1e-6
    assert ora.fusion_classifier.training  # the helper restores the module's mode


def test_modality_shapley_spec_properties():
    """The oracle's enumeration: efficiency (sum phi = f(all) - f(none)), a modality whose slice equals the
    background gets exactly 0, and the closed-form weight matrix of the product path gives the same numbers."""
    import ecgmm  # noqa: F401
    from ecgmm import explain

    ora = make_oracle(seed=7)
    g = torch.Generator().manual_seed(3)
    e = torch.randn(5, 768, generator=g)
    bg = torch.randn(768, generator=g)
    e[:, 256:512] = bg[256:512]  # the signal slice carries no information beyond the background
    phi, f0, f1 = om.modality_shapley(ora.fusion_classifier, e, bg)
    assert phi.shape == (5, 3)
    assert torch.allclose(phi.sum(1), f1 - f0, atol=1e-6)
    assert float(phi[:, 1].abs().max()) < 1e-7
    masks = explain.coalition_masks((256, 256, 256))
    assert masks.shape == (8, 768) and masks[0].sum() == 0 and masks[7].sum() == 768 and masks[2, 256:512].all()
    f = om.perturbation_inference(ora.fusion_classifier, e, bg, masks, 1)
    assert torch.allclose(f @ explain.shapley_matrix(), phi, atol=1e-6)
    w = explain.shapley_matrix()
    assert torch.allclose(w.sum(0), torch.zeros(3), atol=1e-7) and torch.allclose(w[7], torch.full((3,), 1 / 3))


def _eg_case(S=3, K=40, NB=12, seed=5):
    g = torch.Generator().manual_seed(seed)
    e = torch.randn(S, 768, generator=g)
    bg = torch.randn(NB, 768, generator=g)
    idx = torch.randint(0, NB, (S, K), generator=g, dtype=torch.int32)
    alpha = torch.rand(S, K, generator=g)
    return e, bg, idx, alpha


def test_expected_gradients_spec_properties():
    """SURVEY.md section 8f rank 3 (self-contained spec; `shap` is absent): the autograd oracle against (i) the closed
    form the CUDA path evaluates -- W1^T([W1 x + b1 > 0] * w2[c]) around two GEMMs --, (ii) the exact answer for a
    head that is linear on the path, (iii) completeness: on a fine alpha grid against ONE background the attributions
    sum to logit(e) - logit(background)."""
    ora = make_oracle(seed=7)
    head = ora.fusion_classifier
    e, bg, idx, alpha = _eg_case()
    S, K, D = 3, 40, 768
    phi = om.expected_gradients(head, e, bg, idx, alpha)
    assert phi.shape == (S, D, 2)
    # (i) the device algorithm, step by step in torch
    W1, b1, W2 = head[0].weight.detach(), head[0].bias.detach(), head[3].weight.detach()
    b = bg[idx.long()]
    diff = e[:, None] - b
    pts = (b + alpha[..., None] * diff).reshape(S * K, D)
    hidden = torch.relu(pts @ W1.T + b1)
    gate = (hidden > 0).float()[None] * W2[:, None, :]            # [C, rows, HID]
    grad = (gate.reshape(-1, W1.shape[0]) @ W1).view(2, S, K, D)  # one [C*rows, HID] x [HID, D] product
    phi_dev = (diff[None] * grad).mean(2).permute(1, 2, 0)
    assert torch.allclose(phi, phi_dev, atol=1e-7)
    # (ii) a bias so large that every hidden unit is active: the head is affine, phi = (e - mean_k bg_jk) * (W2 W1)^T
    import copy

    lin = copy.deepcopy(head)
    with torch.no_grad():
        lin[0].bias.fill_(1e3)
    phi_lin = om.expected_gradients(lin, e, bg, idx, alpha)
    expect = (e - b.mean(1))[:, :, None] * (W2 @ W1).T[None]
    assert torch.allclose(phi_lin, expect, atol=1e-5)
    # (iii) completeness of the path integral
    Kf = 2000
    idx1 = torch.zeros(1, Kf, dtype=torch.int32)
    alpha1 = ((torch.arange(Kf) + 0.5) / Kf).view(1, Kf)
    phi1 = om.expected_gradients(head, e[:1], bg[:1], idx1, alpha1)
    head.eval()
    with torch.no_grad():
        delta = head(e[:1]) - head(bg[:1])
    assert torch.allclose(phi1.sum(1), delta, atol=2e-4)


def test_modality_share_spec():
    """shap_fusion_modal_balance.py:177-200: mean |phi| per modality slice as a percentage of the three."""
    g = torch.Generator().manual_seed(2)
    phi = torch.randn(4, 768, 2, generator=g)
    phi[:, 256:512] *= 3.0
    sh = om.modality_share(phi)
    assert sh.shape == (4, 2, 3) and torch.allclose(sh.sum(-1), torch.full((4, 2), 100.0), atol=1e-4)
    a = phi[1, :, 0].abs()
    ref = torch.stack([a[:256].mean(), a[256:512].mean(), a[512:].mean()])
    assert torch.allclose(sh[1, 0], ref / ref.sum() * 100, atol=1e-4) and float(sh[:, :, 1].min()) > 50
    assert float(om.modality_share(torch.zeros(2, 768, 2)).abs().max()) == 0.0
    sh2 = om.modality_share(phi[:, :672], dims=(512, 128, 32))
    assert torch.allclose(sh2.sum(-1), torch.full((4, 2), 100.0), atol=1e-4)


def test_image_endpoint_gradcam_spec():
    """the oracle's autograd Grad-CAM against the closed form the CUDA path evaluates
    (classifier row -> LayerNorm backward -> fc^T -> 1/(h w)), and the image-only chain against the full forward."""
    ora = make_oracle(seed=7)
    ora.eval()
    g = torch.Generator().manual_seed(9)
    image = torch.randn(3, 3, 64, 160, generator=g).clamp_(-1, 1)
    probs, cam, cls = om.image_endpoint(ora, image, class_index=1)
    with torch.no_grad():
        full = ora(image, torch.randn(3, 600, generator=g), torch.randn(3, 24, generator=g))
        enc = ora.image_encoder
        x = enc.maxpool(enc.relu(enc.bn1(enc.conv1(image))))
        act = enc.layer4(enc.layer3(enc.layer2(enc.layer1(x))))
        pooled = act.mean((2, 3))
        feat = enc.fc(pooled)
    assert torch.allclose(probs, torch.softmax(full[0], 1), atol=1e-6) and cam.shape == (3,) + tuple(act.shape[2:])
    assert float(cam.min()) >= 0.0 and float(cam.max()) > 0.0 and cls.tolist() == [1, 1, 1]
    ln, wc = ora.image_norm, ora.image_classifier.weight.detach()
    mu = feat.mean(1, keepdim=True)
    rstd = (feat.var(1, unbiased=False, keepdim=True) + ln.eps).rsqrt()
    xhat = (feat - mu) * rstd
    dxh = wc[1][None] * ln.weight.detach()[None]                                 # d logit_1 / d xhat
    dfeat = rstd * (dxh - dxh.mean(1, keepdim=True) - xhat * (dxh * xhat).mean(1, keepdim=True))
    dpooled = dfeat @ enc.fc.weight.detach()
    cam2 = torch.relu((dpooled[:, :, None, None] * act).sum(1) / (act.shape[2] * act.shape[3]))
    assert torch.allclose(cam, cam2, atol=1e-6 + 1e-4 * float(cam.max()))
    p2, cam3, cls3 = om.image_endpoint(ora, image)  # argmax classes
    assert cls3.tolist() == probs.argmax(1).tolist() and not ora.training


def test_attribution_oracles_match_reference_golden():
    """tests/golden/attrib_g2.pt (oracle/gen_golden_attrib.py): expected gradients and Grad-CAM evaluated on the REAL
    reference model's own fusion_classifier / image branch with torch.autograd -- the oracle reproduces them."""
    import os

    from golden_util import GOLDEN_DIR

    gold = torch.load(os.path.join(GOLDEN_DIR, "attrib_g2.pt"))
    ora = make_oracle(seed=7)
    phi = om.expected_gradients(ora.fusion_classifier, gold["e"], gold["bg"], gold["idx"], gold["alpha"])
    assert torch.allclose(phi, gold["phi"], atol=1e-7)
    assert torch.allclose(om.modality_share(phi), gold["share"], atol=1e-4)
    image = (gold["u8"].float() / 255.0 - 0.5) / 0.5
    probs, cam, _ = om.image_endpoint(ora, image, class_index=gold["class_index"])
    assert torch.allclose(probs, gold["probs"], atol=1e-6) and torch.allclose(cam, gold["cam"], atol=1e-7)


def test_masked_regression_is_sklearn_ridge_and_host_operator_matches():
    """The LIME regressor (lime_fusion_modal_balance.py:158-160 -> sklearn Ridge(alpha=1, fit_intercept=True,
    sample_weight=kernel weights)): the oracle's closed form against sklearn itself, and libecgmm's host-side operator
    design (ecgmm_ridge_operator, the product path's plan step: runs on the CPU by design) against the oracle."""
    import numpy as np
    from sklearn.linear_model import Ridge

    import ecgmm  # noqa: F401
    from ecgmm import explain, lib

    ora = make_oracle(seed=7)
    g = torch.Generator().manual_seed(4)
    S, V, D = 3, 260, 768
    e, bg = torch.randn(S, D, generator=g), torch.randn(D, generator=g)
    masks, w = explain.lime_plan(V, D, seed=9)
    assert masks.shape == (V, D) and masks[0].all() and w.dtype == torch.float64 and float(w[0]) == 1.0
    coef, b = om.masked_regression(ora.fusion_classifier, e, bg, masks, w, alpha=1.0)
    f = om.perturbation_inference(ora.fusion_classifier, e, bg, masks, 1)
    for s in range(S):
        r = Ridge(alpha=1.0, fit_intercept=True).fit(masks.numpy().astype(np.float64), f[s].double().numpy(),
                                                     sample_weight=w.numpy())
        assert np.abs(r.coef_ - coef[s].numpy()).max() < 1e-7 and abs(r.intercept_ - float(b[s])) < 1e-6
    R = explain.regression_operator(masks, w, 1.0)
    assert R.shape == (D + 1, V) and R.dtype == torch.float32
    fit = f @ R.T
    assert torch.allclose(fit[:, :D], coef, atol=1e-6) and torch.allclose(fit[:, D], b, atol=1e-6)
    # KernelSHAP-style weights (a few huge ones) and another alpha go through the same operator
    w2 = torch.rand(V, generator=g, dtype=torch.float64)
    w2[:2] = 1e4
    c2, b2 = om.masked_regression(ora.fusion_classifier, e, bg, masks, w2, alpha=0.1)
    fit2 = f @ explain.regression_operator(masks, w2, 0.1).T
    assert torch.allclose(fit2[:, :D], c2, atol=1e-5) and torch.allclose(fit2[:, D], b2, atol=1e-5)
    # summed |coefficients| per modality (lime_fusion_modal_balance.py:163-175)
    sh = om.modality_share(coef.unsqueeze(-1), reduce="sum")
    assert sh.shape == (S, 1, 3) and torch.allclose(sh.sum(-1), torch.full((S, 1), 100.0), atol=1e-3)
    with pytest.raises(lib.EcgmmError):  # alpha = 0 and a constant column (row 0 is all ones, make column 0 all ones)
        m = masks.clone()
        m[:, 0] = 1
        explain.regression_operator(m, w, 0.0)
    with pytest.raises(lib.EcgmmError):
        explain.regression_operator(masks, -w, 1.0)
    with pytest.raises(lib.EcgmmError):
        explain.regression_operator(masks, w[:-1], 1.0)
