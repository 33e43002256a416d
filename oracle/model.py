"""ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU fp32 restatement of the reference hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module, and only as the checker or as the timed CPU baseline -- never as the
product path (the product is ecg-multimodal-model_b200/ + libecgmm.so and has no CPU fallback).

Pinned: oracle/gen_golden.py imports the real reference from /root/reference, copies one
state_dict into both models and asserts bit-identical outputs, losses and gradients (eval and
train mode, same torch RNG state) before it writes tests/golden/*.  The known-answer logits of
the reference's only real checkpoint (best_ptbxl.pth) are checked too.
FocalLoss is pinned against signal_model.FocalLoss by oracle/gen_golden_focal.py (bit-identical value and gradient),
perturbation_inference against the reference model's own fusion_classifier by oracle/gen_golden_perturb.py.

Each class cites the reference lines it restates:
  AttentionFusion      multimodal_paper_modal_balance.py:31-46   (= multimodal.py:12-27)
  SEBlock              multimodal_paper_modal_balance.py:49-64   (= signal_model.py:12-27)
  BasicBlock1D         multimodal_paper_modal_balance.py:67-93   (= signal_model.py:30-56)
  ResNet1D_SE          multimodal_paper_modal_balance.py:96-125  (= signal_model.py:59-88)
  ResNet18 / BasicBlock2D   torchvision 0.26 models/resnet.py:59-104,166-284 (third party,
                       reached from multimodal_paper_modal_balance.py:210,221)
  ECGMultimodalModel   multimodal_paper_modal_balance.py:197-354 (G2); dims option covers
                       multimodal.py:333-469 (G3) minus its TabNet clinical encoder
  FocalLoss            signal_model.py:91-106
  fusion_train_step    train.py:60-86 (zero_grad, forward, CE + 0.1*var_loss, backward, Adam step)
  z_score              signal_model.py:203-206
  perturbation_inference   configs[3]: fusion_classifier.py:5-11 driven as in
                       shap_fusion_modal_balance.py:126-159 / lime_fusion_modal_balance.py:118-160
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class AttentionFusion(nn.Module):
    def __init__(self, dims):
        super().__init__()
        self.weights = nn.Parameter(torch.ones(3))
        self.norm = nn.LayerNorm(sum(dims))

    def forward(self, img_feat, signal_feat, clinical_feat):
        soft_weights = torch.softmax(self.weights, dim=0)
        fused = torch.cat(
            [soft_weights[0] * img_feat, soft_weights[1] * signal_feat, soft_weights[2] * clinical_feat], dim=1
        )
        return self.norm(fused), soft_weights


class SEBlock(nn.Module):
    def __init__(self, channels, reduction=16):
        super().__init__()
        self.pool = nn.AdaptiveAvgPool1d(1)
        self.fc = nn.Sequential(
            nn.Linear(channels, channels // reduction),
            nn.ReLU(),
            nn.Linear(channels // reduction, channels),
            nn.Sigmoid(),
        )

    def forward(self, x):
        b, c, _ = x.size()
        y = self.pool(x).view(b, c)
        y = self.fc(y).view(b, c, 1)
        return x * y


class BasicBlock1D(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1):
        super().__init__()
        padding = kernel_size // 2
        self.conv1 = nn.Conv1d(in_channels, out_channels, kernel_size, stride=stride, padding=padding)
        self.bn1 = nn.BatchNorm1d(out_channels)
        self.relu = nn.ReLU()
        self.conv2 = nn.Conv1d(out_channels, out_channels, kernel_size, padding=padding)
        self.bn2 = nn.BatchNorm1d(out_channels)
        self.se = SEBlock(out_channels)
        self.downsample = None
        if in_channels != out_channels or stride != 1:
            self.downsample = nn.Sequential(
                nn.Conv1d(in_channels, out_channels, kernel_size=1, stride=stride), nn.BatchNorm1d(out_channels)
            )

    def forward(self, x):
        identity = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        out = self.se(out)
        if self.downsample is not None:
            identity = self.downsample(x)
        out = out + identity
        return self.relu(out)


class ResNet1D_SE(nn.Module):
    def __init__(self, input_channels=1, num_classes=2, base_filters=64):
        super().__init__()
        self.initial = nn.Sequential(
            nn.Conv1d(input_channels, base_filters, kernel_size=7, stride=2, padding=3),
            nn.BatchNorm1d(base_filters),
            nn.ReLU(),
            nn.MaxPool1d(kernel_size=3, stride=2, padding=1),
        )
        self.layer1 = BasicBlock1D(base_filters, base_filters)
        self.layer2 = BasicBlock1D(base_filters, base_filters * 2, stride=2)
        self.layer3 = BasicBlock1D(base_filters * 2, base_filters * 4, stride=2)
        self.global_pool = nn.AdaptiveAvgPool1d(1)
        self.classifier = nn.Sequential(
            nn.Flatten(), nn.Linear(base_filters * 4, 64), nn.ReLU(), nn.Dropout(0.3), nn.Linear(64, num_classes)
        )

    def forward(self, x):
        x = self.initial(x)
        x = self.layer1(x)
        x = self.layer2(x)
        x = self.layer3(x)
        x = self.global_pool(x)
        return self.classifier(x)


class BasicBlock2D(nn.Module):
    """torchvision BasicBlock (resnet.py:59-104): conv3x3-BN-ReLU-conv3x3-BN (+downsample) -ReLU."""

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = None
        if stride != 1 or inplanes != planes:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))

    def forward(self, x):
        identity = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        out = out + identity
        return self.relu(out)


class ResNet18(nn.Module):
    """torchvision resnet18 (resnet.py:166-284) with the same attribute / state_dict names."""

    def __init__(self, num_classes=1000):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = nn.Sequential(BasicBlock2D(64, 64), BasicBlock2D(64, 64))
        self.layer2 = nn.Sequential(BasicBlock2D(64, 128, 2), BasicBlock2D(128, 128))
        self.layer3 = nn.Sequential(BasicBlock2D(128, 256, 2), BasicBlock2D(256, 256))
        self.layer4 = nn.Sequential(BasicBlock2D(256, 512, 2), BasicBlock2D(512, 512))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512, num_classes)
        for m in self.modules():  # resnet.py:207-212
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        x = torch.flatten(self.avgpool(x), 1)
        return self.fc(x)


class ECGMultimodalModel(nn.Module):
    """G2 fusion model.  `dims` = (image_dim, signal_dim, clinical_dim): (256,256,256) for G2
    (multimodal_paper_modal_balance.py:202-206), (512,128,32) for G3 (multimodal.py:338-341)."""

    def __init__(self, num_classes=2, dims=(256, 256, 256), clinical_features=24, signal_channels=1):
        super().__init__()
        self.image_dim, self.signal_dim, self.clinical_dim = dims
        self.modal_dim = self.image_dim
        self.image_encoder = ResNet18()
        self.image_encoder.fc = nn.Linear(512, self.image_dim)
        self.image_norm = nn.LayerNorm(self.image_dim)
        self.signal_encoder = ResNet1D_SE(input_channels=signal_channels, num_classes=self.signal_dim)
        self.signal_norm = nn.LayerNorm(self.signal_dim)
        self.clinical_encoder = nn.Sequential(
            nn.Linear(clinical_features, 64), nn.BatchNorm1d(64), nn.ReLU(), nn.Dropout(0.3),
            nn.Linear(64, self.clinical_dim),
        )
        self.clinical_norm = nn.LayerNorm(self.clinical_dim)
        self.image_classifier = nn.Linear(self.image_dim, num_classes)
        self.signal_classifier = nn.Linear(self.signal_dim, num_classes)
        self.clinical_classifier = nn.Linear(self.clinical_dim, num_classes)
        self.attention_fusion = AttentionFusion(dims=list(dims))
        self.fusion_classifier = nn.Sequential(
            nn.Linear(sum(dims), 128), nn.ReLU(), nn.Dropout(0.3), nn.Linear(128, num_classes)
        )

    def forward(self, image, ecg_signal, clinical):
        img_feat = self.image_norm(self.image_encoder(image))
        if ecg_signal.dim() == 2:
            ecg_signal = ecg_signal.unsqueeze(1)
        signal_feat = self.signal_norm(self.signal_encoder(ecg_signal))
        clinical_feat = self.clinical_norm(self.clinical_encoder(clinical))
        img_logits = self.image_classifier(img_feat)
        signal_logits = self.signal_classifier(signal_feat)
        clinical_logits = self.clinical_classifier(clinical_feat)
        fused, soft_weights = self.attention_fusion(img_feat, signal_feat, clinical_feat)
        fusion_logits = self.fusion_classifier(fused)
        var_img = torch.var(img_feat, dim=1).mean()
        var_signal = torch.var(signal_feat, dim=1).mean()
        var_clinical = torch.var(clinical_feat, dim=1).mean()
        var_loss = (
            torch.abs(var_img - var_signal) + torch.abs(var_img - var_clinical) + torch.abs(var_signal - var_clinical)
        )
        return img_logits, signal_logits, clinical_logits, fusion_logits, var_loss, soft_weights


class FocalLoss(nn.Module):
    def __init__(self, alpha=1.0, gamma=2.0):
        super().__init__()
        self.alpha, self.gamma = alpha, gamma

    def forward(self, inputs, targets):
        ce = F.cross_entropy(inputs, targets, reduction="none")
        pt = torch.exp(-ce)
        return (self.alpha * (1 - pt) ** self.gamma * ce).mean()


def z_score(x: torch.Tensor, dim=-1) -> torch.Tensor:
    """(x - mean) / (std_population + 1e-8) along `dim` (signal_model.py:203-206, numpy semantics)."""
    mean = x.mean(dim=dim, keepdim=True)
    std = x.var(dim=dim, unbiased=False, keepdim=True).sqrt()
    return (x - mean) / (std + 1e-8)


def fusion_loss(outputs, labels, var_weight=0.1):
    """train.py:72,78 -- CrossEntropyLoss(fusion_logits) + 0.1 * var_loss."""
    return F.cross_entropy(outputs[3], labels) + var_weight * outputs[4]


def branch_fusion_loss(outputs, labels):
    """train_exhausted.py:70-75 -- the sum of four cross entropies: image, signal and clinical branch heads + fusion."""
    return (F.cross_entropy(outputs[0], labels) + F.cross_entropy(outputs[1], labels) +
            F.cross_entropy(outputs[2], labels) + F.cross_entropy(outputs[3], labels))


def fusion_train_step(model, optimizer, image, ecg_signal, clinical, labels):
    """One iteration of train.py:60-86.  Returns (total_loss, outputs)."""
    optimizer.zero_grad()
    outputs = model(image, ecg_signal, clinical)
    loss = fusion_loss(outputs, labels)
    loss.backward()
    optimizer.step()
    return loss.detach(), outputs


def freeze_encoders(model):
    """train.py:35-40."""
    for enc in (model.image_encoder, model.signal_encoder, model.clinical_encoder):
        for p in enc.parameters():
            p.requires_grad = False


def perturbation_inference(fusion_classifier, e, background, masks, class_index=1):
    """configs[3] / SURVEY.md section 8d cfg4 in plain fp32 torch: every masked variant of every sample's fused
    embedding, z*e + (1-z)*background, through fusion_classifier in eval mode (what the explainers of
    shap_fusion_modal_balance.py:135,159 / lime_fusion_modal_balance.py:118-160 evaluate row by row), reduced to
    softmax(logits)[..., class_index] (class_index < 0: logits).  e [S,D], background [D], masks [V,D] in {0,1}."""
    was_training = fusion_classifier.training
    fusion_classifier.eval()
    with torch.no_grad():
        z = masks.to(e.dtype)
        variants = z.unsqueeze(0) * e.unsqueeze(1) + (1.0 - z).unsqueeze(0) * background.view(1, 1, -1)
        logits = fusion_classifier(variants.reshape(-1, e.shape[1])).view(e.shape[0], masks.shape[0], -1)
    fusion_classifier.train(was_training)
    return logits if class_index < 0 else F.softmax(logits, dim=-1)[..., class_index]


def modality_shapley(fusion_classifier, e, background, dims=(256, 256, 256), class_index=1):
    """Exact 3-player Shapley values of the modalities under background replacement (the self-contained spec of
    SURVEY.md section 8f rank 3; the reference reaches a per-modality importance through the unpinned `shap` / `lime`
    packages, shap_fusion_modal_balance.py:177-200).  Plain enumeration: phi_i = sum_S w(|S|) (f(S+i) - f(S))."""
    from itertools import combinations
    from math import factorial

    D = sum(dims)
    offs = [0, dims[0], dims[0] + dims[1], D]

    def f(players):
        z = torch.zeros(1, D, dtype=torch.uint8)
        for k in players:
            z[0, offs[k]:offs[k + 1]] = 1
        return perturbation_inference(fusion_classifier, e, background, z, class_index)[:, 0]

    phi = torch.zeros(e.shape[0], 3)
    for i in range(3):
        others = [k for k in range(3) if k != i]
        for r in range(3):
            for S in combinations(others, r):
                w = factorial(len(S)) * factorial(2 - len(S)) / factorial(3)
                phi[:, i] += w * (f(S + (i,)) - f(S))
    return phi, f(()), f((0, 1, 2))


def expected_gradients(fusion_classifier, e, background, idx, alpha):
    """Expected-gradients attribution of the fusion head's LOGITS, by autograd.

    The reference obtains `shap_values [S, D, C]` from shap.GradientExplainer(FusionClassifierWrapper, bg_embeddings)
    (shap_fusion_modal_balance.py:135,159); `shap` is unpinned and not installed, so this is the published estimator
    with an explicit sampling plan instead of the package's internal RNG: for sample s and draw k, a background row
    idx[s, k] and an interpolation weight alpha[s, k] in [0, 1):
        phi[s, d, c] = mean_k (e[s, d] - bg[idx[s,k], d]) * d logit_c / d x_d (bg[idx] + alpha (e[s] - bg[idx]))
    e [S, D], background [NB, D], idx [S, K] integer, alpha [S, K].  Eval mode (dropout = identity).
    PARITY UNPINNED against `shap` itself (absent, unpinned by the reference).  Pinned: the same estimator evaluated
    with torch.autograd on the REAL reference model's fusion_classifier gives bit-identical attributions
    (oracle/gen_golden_attrib.py -> tests/golden/attrib_g2.pt), plus the estimator's own properties."""
    was_training = fusion_classifier.training
    fusion_classifier.eval()
    S, D = e.shape
    K = idx.shape[1]
    b = background[idx.long()]                                    # [S, K, D]
    diff = e.unsqueeze(1) - b
    pts = (b + alpha.unsqueeze(-1).to(e.dtype) * diff).detach().reshape(S * K, D).requires_grad_(True)
    logits = fusion_classifier(pts)
    C = logits.shape[1]
    phi = torch.zeros(S, D, C, dtype=e.dtype)
    for c in range(C):
        (g,) = torch.autograd.grad(logits[:, c].sum(), pts, retain_graph=c + 1 < C)
        phi[:, :, c] = (diff * g.view(S, K, D)).mean(1)
    fusion_classifier.train(was_training)
    return phi


def masked_regression(fusion_classifier, e, background, masks, weights, alpha=1.0, class_index=1):
    """The regressor of lime_fusion_modal_balance.py:158-160 -- lime's default model_regressor, sklearn
    Ridge(alpha=1, fit_intercept=True) with the kernel weights as sample_weight -- fitted to the class probability on the
    binary keep-masks of perturbation_inference (lime's own sampler is unpinned and absent; the plan is explicit).
    Closed form in float64: centre with the weighted means, solve the normal equations.  Returns (coef [S, D],
    intercept [S]) float32.  Pinned against sklearn.linear_model.Ridge in tests/test_oracle_cpu.py."""
    f = perturbation_inference(fusion_classifier, e, background, masks, class_index).double()          # [S, V]
    Z, pi = masks.double(), weights.double()
    sw = pi.sum()
    zbar, fbar = (pi[:, None] * Z).sum(0) / sw, (f * pi).sum(1) / sw
    Zc = Z - zbar
    G = Zc.T @ (pi[:, None] * Zc) + alpha * torch.eye(Z.shape[1], dtype=torch.float64)
    w = torch.linalg.solve(G, Zc.T @ (pi[:, None] * (f - fbar[:, None]).T))                            # [D, S]
    return w.T.float().contiguous(), (fbar - w.T @ zbar).float()


def modality_share(phi, dims=(256, 256, 256), reduce="mean"):
    """shap_fusion_modal_balance.py:177-200: per sample and class, the mean |attribution| of the image / signal /
    clinical slices of the fused embedding as a percentage of their sum.  phi [S, D, C] -> [S, C, 3]
    (all-zero attributions give 0, the guard of lime_fusion_modal_balance.py:171-173)."""
    a = phi.abs()
    o1, o2 = dims[0], dims[0] + dims[1]
    red = (lambda t: t.mean(1)) if reduce == "mean" else (lambda t: t.sum(1))  # lime_fusion_modal_balance.py:163-175 sums
    m = torch.stack([red(a[:, :o1]), red(a[:, o1:o2]), red(a[:, o2:])], dim=-1)  # [S, C, 3]
    total = m.sum(-1, keepdim=True)
    return torch.where(total > 0, m / total.clamp_min(1e-38) * 100.0, torch.zeros_like(m))


def image_endpoint(model, image, class_index=None):
    """the image-only chain of multimodal_paper_modal_balance.py:325-327,337 in eval mode
    (image_encoder -> image_norm -> image_classifier), softmax, and the Grad-CAM map of the last ResNet stage:
        cam[n, y, x] = relu( sum_k alpha[n, k] A[n, k, y, x] ),  alpha[n, k] = mean_{y,x} d logit_c / d A[n, k, y, x]
    with A = layer4's output and c = class_index (None: each sample's argmax).  Returns (probs [N, C], cam [N, h, w],
    classes [N]).  The forward chain is the reference's (bit-identical modules); the Grad-CAM generator is not in the
    reference repository (only its output images, gpt/*.png): that half is the published definition (parity with the
    missing script unpinned), checked against autograd on the REAL reference model's image branch
    (oracle/gen_golden_attrib.py -> tests/golden/attrib_g2.pt)."""
    was_training = model.training
    model.eval()
    enc = model.image_encoder
    with torch.no_grad():
        x = enc.maxpool(enc.relu(enc.bn1(enc.conv1(image))))
        x = enc.layer3(enc.layer2(enc.layer1(x)))
    act = enc.layer4(x).detach().requires_grad_(True)
    feat = enc.fc(torch.flatten(enc.avgpool(act), 1))
    logits = model.image_classifier(model.image_norm(feat))
    probs = F.softmax(logits.detach(), dim=1)
    cls = probs.argmax(1) if class_index is None else torch.full((image.shape[0],), int(class_index))
    (g,) = torch.autograd.grad(logits.gather(1, cls.view(-1, 1)).sum(), act)
    alpha = g.mean(dim=(2, 3), keepdim=True)
    cam = F.relu((alpha * act.detach()).sum(1))
    model.train(was_training)
    return probs, cam, cls
