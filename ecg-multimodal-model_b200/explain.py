"""Batched perturbation inference of the fusion head (BASELINE.json configs[3]).

The reference's explainers (shap_fusion_modal_balance.py:100-160, lime_fusion_modal_balance.py:101-160) push
thousands of perturbed copies of one fused embedding through `fusion_classifier`, a few at a time through
torch.nn on the host's schedule.  Here all V masked variants of all S samples go through ONE tensor-core GEMM:

    variant[s][v] = z[v] * e[s] + (1 - z[v]) * background          z[v] in {0,1}^D
    prob[s][v]    = softmax(fusion_classifier(variant[s][v]))[class_index]

`fusion_classifier` is the model's MLPHead (Linear(D,128) -> ReLU -> Dropout -> Linear(128,C)); inference
runs in eval semantics (dropout = identity) whatever the module's mode, like the explainers (model.eval()).
Operands of the first Linear are bf16 (fp32 accumulation); bias, ReLU, the second Linear and the softmax are
fp32.  Sharding over GPUs: samples are independent -- give each rank its slice of `e` (parallel.shard_batch);
there is no collective on this path.
"""
from __future__ import annotations

import torch

from . import lib, ops

BF16 = torch.bfloat16
F32 = torch.float32
# ECGMM_PERTURB_FUSED=0 keeps the three-kernel path (variants and hidden layer through HBM) for A/B measurements; shapes
# the fused kernel does not cover (D % 64 != 0, D > 768, hidden != 128) take it regardless.
FUSED = __import__("os").environ.get("ECGMM_PERTURB_FUSED", "1") != "0"


def _head_weights(head):
    """(w1 bf16 [HID][1][1][D], b1, w2, b2) of an MLPHead, the bf16 copy cached on the module per weight version."""
    head = getattr(head, "fusion_classifier", head)  # FusionClassifierWrapper (shap_fusion_modal_balance.py:100-108)
    lin1, lin2 = head.lin1, head.lin2
    w = lin1.weight
    key = (w.data_ptr(), w._version)
    if head.__dict__.get("_perturb_key") != key:
        w1 = w.detach().to(BF16).contiguous().view(w.shape[0], 1, 1, w.shape[1])
        head.__dict__["_perturb_w1"] = w1
        head.__dict__["_perturb_key"] = key
    return head.__dict__["_perturb_w1"], lin1.bias.detach(), lin2.weight.detach(), lin2.bias.detach()


def masked_variants(e: torch.Tensor, background: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
    """e [S,D] fp32, background [D] fp32, masks [V,D] uint8/bool -> variants [S,V,D] bf16."""
    for t, name in ((e, "e"), (background, "background"), (masks, "masks")):
        if not t.is_cuda:
            raise lib.EcgmmError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if e.dim() != 2 or background.dim() != 1 or masks.dim() != 2 or not (e.shape[1] == background.shape[0] == masks.shape[1]):
        raise lib.EcgmmError(f"shapes must be e [S,D], background [D], masks [V,D]; got {tuple(e.shape)}, "
                             f"{tuple(background.shape)}, {tuple(masks.shape)}")
    S, D = e.shape
    V = masks.shape[0]
    if masks.dtype == torch.bool:
        masks = masks.view(torch.uint8)
    if masks.dtype != torch.uint8:
        raise lib.EcgmmError(f"masks must be uint8 or bool, got {masks.dtype}")
    e = e.detach().to(F32).contiguous()
    background = background.detach().to(F32).contiguous()
    masks = masks.contiguous()
    out = torch.empty((S, V, D), dtype=BF16, device=e.device)
    lib.call("ecgmm_perturb_build", ops._ptr(e), ops._ptr(background), ops._ptr(masks), ops._ptr(out), S, V, D,
             ops._s())
    return out


def head_inference(head, x_bf16: torch.Tensor, class_index: int = 1, weights=None) -> torch.Tensor:
    """x [rows, D] bf16 -> softmax(fusion_classifier(x))[:, class_index] fp32 (class_index < 0: the logits).
    weights: (w1, b1, w2, b2) as _head_weights returns them, when the caller has already prepared (padded) them."""
    ops._chk(x_bf16, BF16, "x")
    rows, D = x_bf16.shape
    w1, b1, w2, b2 = weights if weights is not None else _head_weights(head)
    HID, C = w1.shape[0], w2.shape[0]
    if w1.shape[3] != D:
        raise lib.EcgmmError(f"embedding width {D} does not match fusion_classifier[0] ({w1.shape[3]})")
    hidden = ops.conv2d_fwd(x_bf16.view(1, 1, rows, D), w1, 1).view(rows, HID)
    out = torch.empty((rows,) if class_index >= 0 else (rows, C), dtype=F32, device=x_bf16.device)
    lib.call("ecgmm_head_tail", ops._ptr(hidden), ops._ptr(b1), ops._ptr(w2), ops._ptr(b2), ops._ptr(out), rows, HID,
             C, int(class_index), ops._s())
    return out


def _to_bf16(x: torch.Tensor) -> torch.Tensor:
    x = x.detach().to(F32).contiguous()
    y = torch.empty(x.shape, dtype=BF16, device=x.device)
    lib.call("ecgmm_f32_to_bf16", ops._ptr(x), ops._ptr(y), x.numel(), ops._s())
    return y


def _perturbation_inference_fused(w1, b1, w2, b2, e, background, masks, class_index):
    """One kernel for the whole path (csrc/perturb_fused.cu): the variants only ever exist in shared memory."""
    for t, name in ((e, "e"), (background, "background"), (masks, "masks")):
        if not t.is_cuda:
            raise lib.EcgmmError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if e.dim() != 2 or background.dim() != 1 or masks.dim() != 2 or not (e.shape[1] == background.shape[0] == masks.shape[1]):
        raise lib.EcgmmError(f"shapes must be e [S,D], background [D], masks [V,D]; got {tuple(e.shape)}, "
                             f"{tuple(background.shape)}, {tuple(masks.shape)}")
    if w1.shape[3] != e.shape[1]:
        raise lib.EcgmmError(f"embedding width {e.shape[1]} does not match fusion_classifier[0] ({w1.shape[3]})")
    if masks.dtype == torch.bool:
        masks = masks.view(torch.uint8)
    if masks.dtype != torch.uint8:
        raise lib.EcgmmError(f"masks must be uint8 or bool, got {masks.dtype}")
    S, D = e.shape
    V, C = masks.shape[0], w2.shape[0]
    if class_index >= C:
        raise lib.EcgmmError(f"class_index {class_index} out of range for {C} classes")
    eb, bb, masks = _to_bf16(e), _to_bf16(background), masks.contiguous()
    bits = torch.empty((V, D // 32), dtype=torch.int32, device=e.device)
    lib.call("ecgmm_perturb_pack_masks", ops._ptr(masks), ops._ptr(bits), V, D, ops._s())
    out = torch.empty((S, V) if class_index >= 0 else (S, V, C), dtype=F32, device=e.device)
    lib.call("ecgmm_perturb_head_fused", ops._ptr(eb), ops._ptr(bb), ops._ptr(bits), ops._ptr(w1),
             ops._ptr(b1.to(F32).contiguous()), ops._ptr(w2.to(F32).contiguous()), ops._ptr(b2.to(F32).contiguous()),
             ops._ptr(out), S, V, D, C, int(class_index), ops._s())
    return out


def _pad_width(head, e, background, masks):
    """Embedding widths that are not a multiple of 64 (G3's 3 x 224 = 672): zero columns change nothing --
    0 * w = 0 whichever of e / background the mask selects -- so pad every operand to the next multiple of 64."""
    D = e.shape[1]
    pad = (-D) % 64
    w1, b1, w2, b2 = _head_weights(head)
    if w1.shape[3] != D or background.shape[0] != D or masks.shape[1] != D:
        raise lib.EcgmmError(f"embedding width: e {D}, background {background.shape[0]}, masks {masks.shape[1]}, "
                             f"fusion_classifier[0] {w1.shape[3]}")
    if pad == 0:
        return w1, b1, w2, b2, e, background, masks
    if masks.dtype == torch.bool:
        masks = masks.view(torch.uint8)
    P = torch.nn.functional.pad
    return (P(w1.view(w1.shape[0], D), (0, pad)).view(w1.shape[0], 1, 1, D + pad), b1, w2, b2, P(e.detach(), (0, pad)),
            P(background.detach(), (0, pad)), P(masks, (0, pad)))


def perturbation_inference(fusion_classifier, e, background, masks, class_index: int = 1, chunk_samples: int = 0):
    """prob [S, V] (or logits [S, V, C] when class_index < 0) for all masked variants of all samples.

    chunk_samples bounds the bf16 variant buffer (S*V*D*2 bytes; 6.3 MB per sample at V=4096, D=768):
    0 = as many samples per launch as fit in ~4 GB."""
    if e.dim() != 2 or masks.dim() != 2 or background.dim() != 1:
        raise lib.EcgmmError(f"shapes must be e [S,D], background [D], masks [V,D]; got {tuple(e.shape)}, "
                             f"{tuple(background.shape)}, {tuple(masks.shape)}")
    S, V = e.shape[0], masks.shape[0]
    w1, b1, w2, b2, e, background, masks = _pad_width(fusion_classifier, e, background, masks)
    D = e.shape[1]
    if FUSED and ops._shape_query("ecgmm_perturb_head_fused_supported", D, w1.shape[0], w2.shape[0]):
        return _perturbation_inference_fused(w1, b1, w2, b2, e, background, masks, class_index)
    if chunk_samples <= 0:
        chunk_samples = max(1, (4 << 30) // max(1, V * D * 2))
    chunk_samples = min(chunk_samples, 65535)  # grid.y limit of the variant-build kernel
    outs = []
    for s0 in range(0, S, chunk_samples):
        es = e[s0:s0 + chunk_samples]
        x = masked_variants(es, background, masks)
        out = head_inference(fusion_classifier, x.view(-1, D), class_index, weights=(w1, b1, w2, b2))
        outs.append(out.view(es.shape[0], V) if class_index >= 0 else out.view(es.shape[0], V, -1))
    return outs[0] if len(outs) == 1 else torch.cat(outs, 0)


# ---------------------------------------------------------------------------------------------- modality attribution
# SURVEY.md section 8f rank 3.  The reference's explainers end in a per-MODALITY importance (the |SHAP| mass of the
# image / signal / clinical slices of the fused embedding, shap_fusion_modal_balance.py:177-200) obtained from
# third-party samplers (`shap`, `lime`: unpinned, not installed).  With three modalities the Shapley values of
#     f(S) = softmax(fusion_classifier(z_S * e + (1 - z_S) * background))[class],   S subset of {image, signal, clinical}
# need no sampling: all 2^3 coalitions are evaluated exactly (8 of the masked variants above) and
#     phi_i = sum_{S not containing i}  |S|! (2 - |S|)! / 3!  * ( f(S + i) - f(S) ).
# This is the self-contained spec implemented here and in oracle.model.modality_shapley.
def coalition_masks(dims, device=None) -> torch.Tensor:
    """uint8 [8, sum(dims)]: row c keeps modality m (its slice of the fused embedding) iff bit m of c is set."""
    D = int(sum(dims))
    m = torch.zeros((8, D), dtype=torch.uint8)
    off = 0
    for k, d in enumerate(dims):
        for c in range(8):
            if (c >> k) & 1:
                m[c, off:off + d] = 1
        off += d
    return m if device is None else m.to(device)


def shapley_matrix() -> torch.Tensor:
    """[8, 3] fp32: phi = f_coalitions @ this (exact Shapley weights for 3 players, coalition c = bit set)."""
    w = torch.zeros(8, 3)
    fact = [1.0, 1.0, 2.0, 6.0]
    for i in range(3):
        for c in range(8):
            if (c >> i) & 1:
                continue
            size = bin(c).count("1")
            wt = fact[size] * fact[2 - size] / fact[3]
            w[c | (1 << i), i] += wt
            w[c, i] -= wt
    return w


def modality_shapley(fusion_classifier, e, background, dims=(256, 256, 256), class_index: int = 1):
    """Exact Shapley values of the three modalities for every sample.

    e [S, sum(dims)] fused embeddings, background [sum(dims)].  Returns (phi [S, 3], f_none [S], f_all [S]) fp32 on
    e's device; phi.sum(1) == f_all - f_none (efficiency).  One perturbation_inference call with the 8 coalition masks
    plus one [S,8] x [8,3] product on the library's SGEMM."""
    if class_index < 0:
        raise lib.EcgmmError("modality_shapley needs a class index (the value function is a probability)")
    if e.shape[1] != sum(dims):
        raise lib.EcgmmError(f"embedding width {e.shape[1]} != sum of modality widths {sum(dims)}")
    masks = coalition_masks(dims, e.device)
    f = perturbation_inference(fusion_classifier, e, background, masks, class_index).contiguous()  # [S, 8]
    wmat = shapley_matrix().to(e.device).contiguous()
    phi = ops.sgemm(f, wmat, f.shape[0], 3, 8)
    return phi, f[:, 0].contiguous(), f[:, 7].contiguous()


# ---------------------------------------------------------------------------------------------- expected gradients
# SURVEY.md section 8f rank 3, the 768-dimensional attribution.  The reference gets `shap_values [S, D, C]` from
# shap.GradientExplainer(FusionClassifierWrapper(model.fusion_classifier), bg_embeddings).shap_values(fused)
# (shap_fusion_modal_balance.py:135,159): expected gradients of the LOGITS.  `shap` is unpinned and absent, so the
# estimator is restated with an explicit sampling plan (the caller's RNG, not the package's):
#     phi[s, d, c] = mean_k (e[s, d] - bg[j_sk, d]) * d logit_c / d x_d (bg[j_sk] + a_sk (e[s] - bg[j_sk]))
# For Linear(D,HID) -> ReLU -> Linear(HID,C) the gradient at x is W1^T ([W1 x + b1 > 0] * w2[c]): two fp32 SGEMMs
# around three bandwidth-bound kernels (csrc/attrib.cu).  fp32 throughout, unlike perturbation_inference: the SIGN of
# a hidden pre-activation switches a whole column of the gradient on or off.
def sampling_plan(S: int, K: int, n_background: int, seed: int = 0):
    """(idx [S,K] int32, alpha [S,K] fp32) on the host: uniform background rows and interpolation weights, the draws
    shap.GradientExplainer makes internally (nsamples = K, default 200)."""
    g = torch.Generator().manual_seed(int(seed))
    idx = torch.randint(0, n_background, (S, K), generator=g, dtype=torch.int32)
    alpha = torch.rand(S, K, generator=g)
    return idx, alpha


def expected_gradients(fusion_classifier, e, background, idx, alpha, chunk_samples: int = 0):
    """phi [S, D, C] fp32: expected gradients of fusion_classifier's logits (eval semantics).

    e [S, D] and background [NB, D] fp32 CUDA tensors; idx [S, K] (integer, values in [0, NB)) and alpha [S, K] may
    live on the host (they are validated there) or on the device.  chunk_samples bounds the [C, S*K, D] gradient
    buffer (0: ~2 GB)."""
    for t, name in ((e, "e"), (background, "background")):
        if not t.is_cuda:
            raise lib.EcgmmError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if e.dim() != 2 or background.dim() != 2 or e.shape[1] != background.shape[1]:
        raise lib.EcgmmError(f"shapes must be e [S,D], background [NB,D]; got {tuple(e.shape)}, {tuple(background.shape)}")
    S, D = e.shape
    NB = background.shape[0]
    if idx.dim() != 2 or idx.shape[0] != S or tuple(alpha.shape) != tuple(idx.shape) or idx.shape[1] < 1:
        raise lib.EcgmmError(f"idx and alpha must both be [S,K] with K >= 1; got {tuple(idx.shape)}, {tuple(alpha.shape)}")
    if idx.dtype.is_floating_point or idx.dtype == torch.bool:
        raise lib.EcgmmError(f"idx must be an integer tensor, got {idx.dtype}")
    if idx.device.type == "cpu" and idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= NB):
        raise lib.EcgmmError(f"idx values must lie in [0, {NB})")
    K = idx.shape[1]
    head = getattr(fusion_classifier, "fusion_classifier", fusion_classifier)  # FusionClassifierWrapper or the head
    w1, b1 = head.lin1.weight.detach().contiguous(), head.lin1.bias.detach().contiguous()
    w2 = head.lin2.weight.detach().contiguous()
    HID, C = w1.shape[0], w2.shape[0]
    if w1.shape[1] != D:
        raise lib.EcgmmError(f"embedding width {D} does not match fusion_classifier[0] ({w1.shape[1]})")
    dev = e.device
    e = e.detach().to(F32).contiguous()
    background = background.detach().to(F32).contiguous()
    idx = idx.to(device=dev, dtype=torch.int32).contiguous()
    alpha = alpha.to(device=dev, dtype=F32).contiguous()
    if chunk_samples <= 0:
        chunk_samples = max(1, (2 << 30) // max(1, C * K * D * 4))
    chunk_samples = min(chunk_samples, 65535)
    phi = torch.empty((S, D, C), dtype=F32, device=dev)
    for s0 in range(0, S, chunk_samples):
        n = min(chunk_samples, S - s0)
        rows = n * K
        es, ix, al = e[s0:s0 + n], idx[s0:s0 + n], alpha[s0:s0 + n]
        pts = torch.empty((rows, D), dtype=F32, device=dev)
        lib.call("ecgmm_eg_points", ops._ptr(es), ops._ptr(background), ops._ptr(ix), ops._ptr(al), ops._ptr(pts), n, K,
                 D, NB, ops._s())
        hidden = ops.linear_fwd(pts, w1, b1, relu=True)                        # relu(h) > 0  <=>  h > 0
        gate = torch.empty((C, rows, HID), dtype=F32, device=dev)
        lib.call("ecgmm_eg_gate", ops._ptr(hidden), ops._ptr(w2), ops._ptr(gate), rows, HID, C, ops._s())
        grad = ops.sgemm(gate, w1, C * rows, D, HID)                           # [C*rows, HID] x [HID, D]
        lib.call("ecgmm_eg_reduce", ops._ptr(es), ops._ptr(background), ops._ptr(ix), ops._ptr(grad),
                 ops._ptr(phi[s0:s0 + n]), n, K, D, C, NB, ops._s())
    return phi


def modality_share(phi, dims=(256, 256, 256), reduce: str = "mean"):
    """phi [S, D, C] -> [S, C, 3] percentages of the image / signal / clinical slices' |attribution| on the device
    (0 where all three are 0).  reduce="mean": mean |phi| per slice (shap_fusion_modal_balance.py:189-200);
    reduce="sum": summed |phi| per slice (lime_fusion_modal_balance.py:163-175).  Equal for equally wide modalities."""
    if reduce not in ("mean", "sum"):
        raise lib.EcgmmError(f"reduce must be 'mean' or 'sum', got {reduce!r}")
    ops._chk(phi, F32, "phi")
    if phi.dim() != 3 or len(dims) != 3 or phi.shape[1] != sum(dims) or min(dims) < 1:
        raise lib.EcgmmError(f"phi must be [S, {sum(dims)}, C] for modality widths {tuple(dims)}; got {tuple(phi.shape)}")
    S, _, C = phi.shape
    out = torch.empty((S, C, 3), dtype=F32, device=phi.device)
    lib.call("ecgmm_modality_share", ops._ptr(phi), ops._ptr(out), S, C, int(dims[0]), int(dims[1]), int(dims[2]),
             int(reduce == "sum"), ops._s())
    return out


# ---------------------------------------------------------------------------------------------- local surrogate (LIME)
# lime_fusion_modal_balance.py:118-123,158-160: LimeTabularExplainer.explain_instance(fused[b], predict_fn,
# num_features=768, num_samples=1000) perturbs the fused embedding, weighs the perturbed rows with an exponential kernel
# of their distance to the instance and fits lime's default regressor -- sklearn Ridge(alpha=1, fit_intercept=True,
# sample_weight=kernel weights) -- to the class-1 probability; the |coefficients| are then summed per modality (:163-175).
# `lime` is unpinned and absent; its sampler (quartile discretisation of training-set statistics) is replaced by the
# binary keep-masks of the perturbation path with an explicit plan, its REGRESSOR is reproduced exactly (pinned against
# sklearn's Ridge in tests/test_oracle_cpu.py).  The fit is linear in the responses, so its operator is designed once
# per plan on the host (ecgmm_ridge_operator, float64, like the filter taps of preprocess.butter_lowpass) and every
# sample's coefficients are one row of a device GEMM.  Shapley-kernel weights in `weights` give KernelSHAP instead.
def lime_plan(V: int, D: int, seed: int = 0, keep_prob: float = 0.5, kernel_width: float = None):
    """(masks [V, D] uint8, weights [V] float64) on the host.  Row 0 is the instance itself (all kept), like lime's first
    sample; weights = sqrt(exp(-d^2 / kernel_width^2)) with d = Euclidean distance of the binary row to the instance and
    kernel_width = 0.75 sqrt(D) (lime_tabular's defaults)."""
    g = torch.Generator().manual_seed(int(seed))
    masks = (torch.rand(V, D, generator=g) < keep_prob).to(torch.uint8)
    masks[0] = 1
    kw = 0.75 * (D ** 0.5) if kernel_width is None else float(kernel_width)
    d2 = (1 - masks.to(torch.float64)).sum(1)
    return masks, torch.sqrt(torch.exp(-d2 / (kw * kw)))


def regression_operator(masks, weights, alpha: float = 1.0, device=None):
    """R [(D+1), V] fp32: (coefficients, intercept) = R @ responses for the weighted ridge fit on the binary masks.
    masks [V, D] uint8 / bool and weights [V] are HOST tensors (the sampling plan); the design runs in libecgmm on the
    host in float64 and the result is moved to `device` (a CUDA device) when given."""
    if masks.device.type != "cpu" or weights.device.type != "cpu":
        raise lib.EcgmmError("masks and weights are the host-side sampling plan: pass CPU tensors")
    if masks.dim() != 2 or weights.dim() != 1 or weights.shape[0] != masks.shape[0]:
        raise lib.EcgmmError(f"shapes must be masks [V,D], weights [V]; got {tuple(masks.shape)}, {tuple(weights.shape)}")
    if masks.dtype == torch.bool:
        masks = masks.to(torch.uint8)
    if masks.dtype != torch.uint8:
        raise lib.EcgmmError(f"masks must be uint8 or bool, got {masks.dtype}")
    masks = masks.contiguous()
    weights = weights.to(torch.float64).contiguous()
    V, D = masks.shape
    R = torch.empty((D + 1, V), dtype=F32)
    lib.call("ecgmm_ridge_operator", masks.data_ptr(), weights.data_ptr(), V, D, float(alpha), R.data_ptr())
    return R if device is None else R.to(device)


def masked_regression(fusion_classifier, e, background, masks, weights=None, alpha: float = 1.0, class_index: int = 1,
                      operator=None):
    """Local linear surrogate of softmax(fusion_classifier(.))[class_index] around every sample.

    e [S, D], background [D] CUDA tensors; masks [V, D] (host or device; the plan) and weights [V] (host) as returned by
    lime_plan, or a ready `operator` = regression_operator(...) on e's device.  Returns (coef [S, D], intercept [S]):
    all V x S model evaluations through perturbation_inference, then ONE [S, V] x [V, D+1] product."""
    if class_index < 0:
        raise lib.EcgmmError("masked_regression explains a class probability: class_index >= 0")
    S, D = e.shape
    if operator is None:
        if weights is None:
            raise lib.EcgmmError("pass weights (with host-side masks) or a ready operator")
        operator = regression_operator(masks.cpu(), weights, alpha, e.device)
    V = masks.shape[0]
    if tuple(operator.shape) != (D + 1, V):
        raise lib.EcgmmError(f"operator must be [{D + 1}, {V}], got {tuple(operator.shape)}")
    ops._chk(operator, F32, "operator")
    f = perturbation_inference(fusion_classifier, e, background, masks.to(e.device), class_index).contiguous()  # [S, V]
    coef = ops.sgemm(f, operator, S, D + 1, V, transB=True)
    return coef[:, :D].contiguous(), coef[:, D].contiguous()
