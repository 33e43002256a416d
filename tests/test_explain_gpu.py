"""GPU: batched perturbation inference of the fusion head (configs[3]) on the tcgen05 GEMM path against the
fp32 oracle.  Tolerance: the first Linear runs with bf16 operands (fp32 accumulation), everything after it in fp32:
|p - p_ref| <= 1e-2 on probabilities, logits within OUT_TOL.  Widths that are a multiple of 64 (<= 768) run the
one-kernel path (csrc/perturb_fused.cu), the others the three-kernel path; both are checked, and the fused kernel
additionally against the exact fp32 evaluation of its own bf16-rounded operands."""
import pytest
import torch

from ecgmm import explain, lib
from oracle import model as om
from parity_util import OUT_TOL, build_pair, relmax

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _case(S, V, D=768, seed=0):
    g = torch.Generator().manual_seed(seed)
    e = torch.randn(S, D, generator=g)
    bg = torch.randn(100, D, generator=g).mean(0)  # mean of 100 background embeddings
    masks = (torch.rand(V, D, generator=g) < 0.5).to(torch.uint8)
    return e, bg, masks


@pytest.mark.parametrize("S,V", [(3, 256), (1, 4096), (5, 130)])
def test_perturbation_inference_matches_oracle(S, V):
    ora, dut = build_pair(seed=7)
    e, bg, masks = _case(S, V, seed=S * 1000 + V)
    ref_p = om.perturbation_inference(ora.fusion_classifier, e, bg, masks, 1)
    ref_l = om.perturbation_inference(ora.fusion_classifier, e, bg, masks, -1)
    p = explain.perturbation_inference(dut.fusion_classifier, e.to(DEV), bg.to(DEV), masks.to(DEV), 1)
    l = explain.perturbation_inference(dut.fusion_classifier, e.to(DEV), bg.to(DEV), masks.to(DEV), -1)
    assert p.shape == (S, V) and l.shape == (S, V, 2)
    assert (p.cpu() - ref_p).abs().max().item() <= 1e-2
    assert relmax(l, ref_l) <= OUT_TOL
    margin = (ref_l[..., 1] - ref_l[..., 0]).abs() > 2 * OUT_TOL * max(1.0, ref_l.abs().max().item())
    assert torch.equal(l.cpu().argmax(-1)[margin], ref_l.argmax(-1)[margin])


@pytest.mark.parametrize("S,V,D,C", [(1, 1, 64, 2), (3, 127, 256, 5), (2, 129, 768, 2), (300, 130, 768, 2),
                                     (1, 4096, 768, 8), (40, 1000, 768, 2), (5, 128, 192, 3), (7, 700, 192, 2),
                                     (3, 300, 64, 2)])
@pytest.mark.parametrize("pair", ["1", "0"], ids=["cta_pairs", "single_cta"])
def test_fused_kernel_against_exact_evaluation_of_its_operands(S, V, D, C, pair, monkeypatch):
    """The fused kernel selects bf16 bit patterns (exact) and accumulates bf16 x bf16 products in fp32: against torch
    fp32 on the SAME rounded operands only the summation order differs.  (300, 130): 600 tiles, several per CTA, so
    the ring stages and both TMEM accumulators wrap; V = 1 / 127 / 129 / 1000: ragged last tiles.  Both kernels: CTA
    pairs (cta_group::2, 256 variants per tile, half of W1 per CTA, 8-stage ring) and the single-CTA one."""
    from ecgmm.model import MLPHead

    monkeypatch.setenv("ECGMM_PERTURB_PAIR", pair)  # V <= 128 runs the single-CTA kernel either way
    torch.manual_seed(S + V + D + C)
    head = MLPHead(D, 128, C).to(DEV)
    e, bg, masks = _case(S, V, D=D, seed=S + 7 * V)
    e, bg, masks = e.to(DEV), bg.to(DEV), masks.to(DEV)
    assert lib.load().ecgmm_perturb_head_fused_supported(D, 128, C) == 1
    l = explain.perturbation_inference(head, e, bg, masks, -1)
    p = explain.perturbation_inference(head, e, bg, masks, C - 1)
    eb, bb = e.to(torch.bfloat16).float(), bg.to(torch.bfloat16).float()
    w1 = head.lin1.weight.detach().to(torch.bfloat16).float()
    for s0 in range(0, S, 64):
        x = torch.where(masks.bool().unsqueeze(0), eb[s0:s0 + 64].unsqueeze(1), bb.view(1, 1, -1))
        h = torch.relu(x @ w1.t() + head.lin1.bias.detach())
        ref = h @ head.lin2.weight.detach().t() + head.lin2.bias.detach()
        assert relmax(l[s0:s0 + 64], ref) <= 2e-5, (s0, relmax(l[s0:s0 + 64], ref))
        assert (p[s0:s0 + 64] - torch.softmax(ref, -1)[..., C - 1]).abs().max().item() <= 1e-5


def test_fused_and_three_kernel_paths_agree(monkeypatch):
    """Same inputs through both paths: they differ only by the bf16 rounding of the hidden activation that the
    three-kernel path stores; a width the fused kernel does not cover (672 = 3 x 224, G3) takes the latter."""
    _, dut = build_pair(seed=7)
    e, bg, masks = _case(9, 700, seed=21)
    a = explain.perturbation_inference(dut.fusion_classifier, e.to(DEV), bg.to(DEV), masks.to(DEV), -1)
    monkeypatch.setattr(explain, "FUSED", False)
    b = explain.perturbation_inference(dut.fusion_classifier, e.to(DEV), bg.to(DEV), masks.to(DEV), -1)
    assert relmax(a, b) <= OUT_TOL
    monkeypatch.undo()
    from ecgmm.model import MLPHead

    # a width that is not a multiple of 64 (672 = 3 x 224, G3) is zero-padded by the glue, on either path
    assert lib.load().ecgmm_perturb_head_fused_supported(672, 128, 2) == 0
    torch.manual_seed(1)
    head = MLPHead(672, 128, 2).to(DEV)
    e, bg, masks = _case(4, 200, D=672, seed=22)
    x = torch.where(masks.bool().unsqueeze(0), e.unsqueeze(1), bg.view(1, 1, -1)).to(DEV)
    with torch.no_grad():
        ref = torch.relu(x @ head.lin1.weight.t() + head.lin1.bias) @ head.lin2.weight.t() + head.lin2.bias
    for fused in (True, False):
        monkeypatch.setattr(explain, "FUSED", fused)
        l = explain.perturbation_inference(head, e.to(DEV), bg.to(DEV), masks.bool().to(DEV), -1)
        assert l.shape == (4, 200, 2) and relmax(l, ref) <= OUT_TOL


def test_variants_are_exact_selections():
    e, bg, masks = _case(4, 64, D=768, seed=11)
    x = explain.masked_variants(e.to(DEV), bg.to(DEV), masks.to(DEV)).cpu()
    want = torch.where(masks.bool().unsqueeze(0), e.unsqueeze(1), bg.view(1, 1, -1)).to(torch.bfloat16)
    assert torch.equal(x, want)
    xb = explain.masked_variants(e.to(DEV), bg.to(DEV), masks.bool().to(DEV)).cpu()
    assert torch.equal(xb, want)


def test_endpoints_and_chunking():
    """All-ones masks reproduce the model's own fusion logits for e; all-zeros give the background's; chunked
    evaluation equals the single-launch one bit for bit."""
    _, dut = build_pair(seed=7)
    dut.eval()
    e, bg, _ = _case(6, 8, seed=3)
    masks = torch.zeros(2, 768, dtype=torch.uint8)
    masks[1] = 1
    l = explain.perturbation_inference(dut.fusion_classifier, e.to(DEV), bg.to(DEV), masks.to(DEV), -1)
    with torch.no_grad():
        own_e = dut.fusion_classifier(e.to(DEV))
        own_b = dut.fusion_classifier(bg.to(DEV).view(1, -1))
    assert relmax(l[:, 1], own_e) <= OUT_TOL
    assert relmax(l[:, 0], own_b.expand(6, -1)) <= OUT_TOL
    e2, bg2, m2 = _case(7, 96, seed=5)
    a = explain.perturbation_inference(dut.fusion_classifier, e2.to(DEV), bg2.to(DEV), m2.to(DEV), 1)
    b = explain.perturbation_inference(dut.fusion_classifier, e2.to(DEV), bg2.to(DEV), m2.to(DEV), 1, chunk_samples=2)
    assert torch.equal(a, b)


def test_argument_errors():
    _, dut = build_pair(seed=7)
    e, bg, masks = _case(2, 8)
    with pytest.raises(lib.EcgmmError):
        explain.perturbation_inference(dut.fusion_classifier, e, bg, masks)  # CPU tensors
    with pytest.raises(lib.EcgmmError):
        explain.masked_variants(e.to(DEV), bg[:100].to(DEV), masks.to(DEV))
    with pytest.raises(lib.EcgmmError):
        explain.masked_variants(e.to(DEV), bg.to(DEV), masks.to(DEV).float())


def test_perturbation_inference_matches_reference_golden():
    """Against tests/golden/perturb_g2.pt (the REFERENCE module's fusion_classifier on the same variants, see
    oracle/gen_golden_perturb.py), same tolerances as against the live oracle above."""
    import os

    from golden_util import GOLDEN_DIR

    kat = torch.load(os.path.join(GOLDEN_DIR, "perturb_g2.pt"), map_location="cpu", weights_only=False)
    _, dut = build_pair(seed=kat["weights_seed"])
    e, bg, masks = kat["e"].to(DEV), kat["background"].to(DEV), kat["masks"].to(DEV)
    p = explain.perturbation_inference(dut.fusion_classifier, e, bg, masks, 1)
    l = explain.perturbation_inference(dut.fusion_classifier, e, bg, masks, -1)
    assert (p.cpu() - kat["prob1"]).abs().max().item() <= 1e-2
    assert relmax(l, kat["logits"]) <= OUT_TOL
