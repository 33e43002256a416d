"""Pins oracle.model.perturbation_inference (configs[3]) against the REAL reference module and writes
tests/golden/perturb_g2.pt.

Run in the build container only (needs /root/reference):   python oracle/gen_golden_perturb.py

The reference model (multimodal_paper_modal_balance.ECGMultimodalModel, imported as in oracle/gen_golden.py) gets the
procedural weights of tests/golden_util.make_oracle(seed=7); its own `fusion_classifier`, wrapped by the reference's
fusion_classifier.FusionClassifierWrapper (fusion_classifier.py:5-11, what shap_fusion_modal_balance.py:126 explains),
is evaluated row by row on the masked variants z*e + (1-z)*background and must be BIT-IDENTICAL to the oracle's
batched evaluation.  Fused embeddings come from the reference's own attention_fusion on seeded encoder features.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from gen_golden import import_reference  # noqa: E402
from golden_util import GOLDEN_DIR, make_oracle  # noqa: E402
from oracle import model as om  # noqa: E402


def main():
    torch.set_num_threads(8)
    ref, ref_m = import_reference()
    sys.path.insert(0, "/root/reference")
    from fusion_classifier import FusionClassifierWrapper  # the reference's wrapper

    ora = make_oracle(seed=7)
    ref_m.load_state_dict({k: v.clone() for k, v in ora.state_dict().items()}, strict=True)
    ref_m.eval()
    ora.eval()
    g = torch.Generator().manual_seed(42)
    S, V, D = 3, 64, 768
    feats = [torch.randn(S, 256, generator=g) for _ in range(3)]
    with torch.no_grad():
        e_ref, _ = ref_m.attention_fusion(*feats)
        e_ora, _ = ora.attention_fusion(*feats)
    assert torch.equal(e_ref, e_ora)
    bg = torch.randn(100, D, generator=g).mean(0)
    masks = (torch.rand(V, D, generator=g) < 0.5).to(torch.uint8)
    wrapper = FusionClassifierWrapper(ref_m.fusion_classifier).eval()
    rows = []
    with torch.no_grad():
        for s in range(S):
            for v in range(V):
                z = masks[v].float()
                rows.append(wrapper((z * e_ref[s] + (1 - z) * bg).unsqueeze(0))[0])
    logits_ref = torch.stack(rows).view(S, V, -1)
    logits_ora = om.perturbation_inference(ora.fusion_classifier, e_ora, bg, masks, -1)
    prob_ora = om.perturbation_inference(ora.fusion_classifier, e_ora, bg, masks, 1)
    worst = float((logits_ref - logits_ora).abs().max())
    assert worst < 1e-6, worst  # row-wise vs batched GEMM may differ in the last bit
    exact = bool(torch.equal(logits_ref, logits_ora))
    assert torch.allclose(torch.softmax(logits_ref, -1)[..., 1], prob_ora, atol=1e-7)
    path = os.path.join(GOLDEN_DIR, "perturb_g2.pt")
    torch.save({"weights_seed": 7, "e": e_ora, "background": bg, "masks": masks, "logits": logits_ref,
                "prob1": torch.softmax(logits_ref, -1)[..., 1]}, path)
    print(f"oracle perturbation_inference vs the reference's fusion_classifier row by row: "
          f"{'bit-identical' if exact else 'max diff %.2e (batched vs row-wise GEMM)' % float((logits_ref - logits_ora).abs().max())}; "
          f"wrote {path} {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
