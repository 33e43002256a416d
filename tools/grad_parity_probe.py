"""Calibration run for the gradient tolerances of tests/parity_util.py (needs a B200): for several shapes prints, per
parameter tensor, the relative L2 error and cosine of the CUDA path against (a) the fp32 oracle, (b) the oracle with
bf16 storage emulated at the product's rounding points (storage_matched_oracle), next to the error of (b) against (a).
Output of the last run: profiles/r02_grad_parity_probe.txt."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ecgmm  # noqa: E402,F401
from parity_util import run_fusion_parity  # noqa: E402

CASES = [dict(B=4, H=64, W=160, L=600), dict(B=6, H=224, W=224, L=2476, dims=(512, 128, 32)),
         dict(B=4, H=64, W=160, L=600, loss="branches"), dict(B=8, H=96, W=320, L=1200, seed=11)]
if len(sys.argv) > 1 and sys.argv[1] == "native":
    CASES.append(dict(B=2, H=250, W=2500, L=2476))
for c in CASES:
    rep = run_fusion_parity(train=True, adam=False, **c)
    m = rep["matched"]
    rels = sorted(v[0] for v in m.values())
    coss = sorted(v[1] for v in m.values())
    n = len(rels)
    print(f"== {c}: ok={rep['ok']} tensors={n}")
    print(f"   vs storage-matched oracle: rel median {rels[n // 2]:.4f} p90 {rels[int(n * 0.9)]:.4f} max {rels[-1]:.4f};"
          f" cos min {coss[0]:.5f} p10 {coss[n // 10]:.5f}")
    r32 = sorted(v[2] for v in m.values())
    rem = sorted(v[3] for v in m.values())
    print(f"   vs fp32 oracle: rel median {r32[n // 2]:.4f} max {r32[-1]:.4f}; emulated-vs-fp32 median {rem[n // 2]:.4f} max {rem[-1]:.4f}")
    for k, v in sorted(m.items(), key=lambda kv: -kv[1][0])[:8]:
        print(f"   {k:55s} matched rel {v[0]:.4f} cos {v[1]:.5f} | fp32 rel {v[2]:.4f} (emulated {v[3]:.4f})")
    for f in rep["failures"][:5]:
        print("   FAIL", f)
