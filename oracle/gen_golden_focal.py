"""Pins oracle.model.FocalLoss / z_score against the reference's signal_model.py and writes tests/golden/focal_kat.pt.

Run in the build container only (needs /root/reference):   python oracle/gen_golden_focal.py

signal_model.py imports matplotlib and seaborn at module level (neither is installed here and neither touches the
arithmetic), so they are replaced by empty stand-in modules for the import; sklearn, torch and numpy are the real
ones.  FocalLoss (signal_model.py:91-106) must be BIT-IDENTICAL to the oracle's in value and gradient, for 2 and 5
classes; the reference's ResNet1D_SE class of that file must accept the oracle's state_dict and give identical logits
(it is the textual twin of the one in multimodal_paper_modal_balance.py that oracle/gen_golden.py already pins).
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")

from golden_util import GOLDEN_DIR  # noqa: E402
from oracle import model as om  # noqa: E402


def import_signal_model():
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import signal_model  # noqa

    return signal_model


def main():
    sm = import_signal_model()
    out = {"cases": []}
    for C, B, seed in ((2, 16, 1), (5, 9, 2), (2, 1, 3)):
        g = torch.Generator().manual_seed(seed)
        logits = (torch.randn(B, C, generator=g) * 3).requires_grad_(True)
        labels = torch.randint(0, C, (B,), generator=g)
        res = []
        for cls in (sm.FocalLoss, om.FocalLoss):
            z = logits.detach().clone().requires_grad_(True)
            loss = cls()(z, labels)
            loss.backward()
            res.append((loss.detach(), z.grad.clone()))
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]), (C, B)
        out["cases"].append({"logits": logits.detach(), "labels": labels, "loss": res[1][0], "grad": res[1][1]})
    # the 12-lead model class of signal_model.py against the oracle's
    torch.manual_seed(3)
    o_net = om.ResNet1D_SE(12, 2).eval()
    r_net = sm.ResNet1D_SE(12, 2).eval()
    r_net.load_state_dict(o_net.state_dict(), strict=True)
    x = torch.randn(2, 12, 1000, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        assert torch.equal(r_net(x), o_net(x))
    path = os.path.join(GOLDEN_DIR, "focal_kat.pt")
    torch.save(out, path)
    print("oracle FocalLoss == signal_model.FocalLoss (value + gradient, 3 cases); ResNet1D_SE twin identical; wrote",
          path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
