"""Pins the attribution / serving oracles against the REAL reference module and writes
tests/golden/attrib_g2.pt.

Run in the build container only (needs /root/reference):   python oracle/gen_golden_attrib.py

`shap`, `lime` and the Grad-CAM generator are not available (unpinned third-party packages / a script that is not in the
reference repository), so their estimators stay self-specified -- but everything they are estimators OF is the
reference's own code, and that is what is pinned here:
  * expected gradients: the interpolation points of the explicit sampling plan go through the reference model's own
    fusion_classifier behind the reference's FusionClassifierWrapper (fusion_classifier.py:5-11), the gradients come
    from torch.autograd on that module; oracle.model.expected_gradients must reproduce the attribution bit for bit;
  * image endpoint: probabilities and the Grad-CAM map are computed on the reference model's own image_encoder /
    image_norm / image_classifier (a forward hook on layer4 + autograd); oracle.model.image_endpoint must match.
"""
import os
import sys

import torch
import torch.nn.functional as F

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
    g = torch.Generator().manual_seed(43)
    S, K, NB, D = 3, 16, 10, 768
    feats = [torch.randn(S, 256, generator=g) for _ in range(3)]
    with torch.no_grad():
        e, _ = ref_m.attention_fusion(*feats)
    bg = torch.randn(NB, D, generator=g)
    idx = torch.randint(0, NB, (S, K), generator=g, dtype=torch.int32)
    alpha = torch.rand(S, K, generator=g)
    # ---- expected gradients on the reference's own module
    wrapper = FusionClassifierWrapper(ref_m.fusion_classifier).eval()
    b = bg[idx.long()]
    diff = e.unsqueeze(1) - b
    pts = (b + alpha.unsqueeze(-1) * diff).detach().reshape(S * K, D).requires_grad_(True)
    logits = wrapper(pts)
    phi_ref = torch.zeros(S, D, logits.shape[1])
    for c in range(logits.shape[1]):
        (gr,) = torch.autograd.grad(logits[:, c].sum(), pts, retain_graph=True)
        phi_ref[:, :, c] = (diff * gr.view(S, K, D)).mean(1)
    phi = om.expected_gradients(ora.fusion_classifier, e, bg, idx, alpha)
    assert torch.equal(phi, phi_ref), float((phi - phi_ref).abs().max())
    share = om.modality_share(phi)
    # ---- image endpoint + Grad-CAM on the reference's own image branch
    u8 = torch.randint(0, 256, (2, 3, 64, 160), generator=g, dtype=torch.uint8)
    image = (u8.float() / 255.0 - 0.5) / 0.5
    kept = {}
    h = ref_m.image_encoder.layer4.register_forward_hook(lambda m, i, o: kept.__setitem__("act", o))
    feat = ref_m.image_norm(ref_m.image_encoder(image))
    lg = ref_m.image_classifier(feat)
    h.remove()
    act = kept["act"]
    (ga,) = torch.autograd.grad(lg[:, 1].sum(), act)
    cam_ref = F.relu((ga.mean(dim=(2, 3), keepdim=True) * act.detach()).sum(1))
    probs_ref = F.softmax(lg.detach(), 1)
    probs, cam, cls = om.image_endpoint(ora, image, class_index=1)
    assert torch.equal(probs, probs_ref) and torch.allclose(cam, cam_ref, atol=1e-8), float((cam - cam_ref).abs().max())
    path = os.path.join(GOLDEN_DIR, "attrib_g2.pt")
    torch.save({"e": e, "bg": bg, "idx": idx, "alpha": alpha, "phi": phi, "share": share, "u8": u8, "probs": probs,
                "cam": cam, "class_index": 1,
                "note": "expected gradients / Grad-CAM computed on the reference model's own modules (seed-7 weights)"},
               path)
    print(f"wrote {path}: phi max {float(phi.abs().max()):.4f}, cam max {float(cam.max()):.4f}; oracle == reference "
          f"module (bit-identical attributions and probabilities)")


if __name__ == "__main__":
    main()
