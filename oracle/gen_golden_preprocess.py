"""Pins oracle/preprocess.py against the REAL reference preprocessing and writes tests/golden/preprocess.npz.

Run in the build container only (needs /root/reference):   python oracle/gen_golden_preprocess.py

The reference methods are imported from /root/reference/dataset.py (ECGMultimodalDataset.preprocess_signal,
.remove_baseline_drift, .lowpass_filter, .z_score_normalize; dataset.py:76-95).  For every seeded input the
oracle's library form must be BIT-IDENTICAL to them and the restated numpy-loop form within 1e-10; the inputs
(float32, so they are exactly representable on the GPU path) and float64 outputs become the fixture.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import preprocess as op  # noqa: E402


def synth(rs, L):
    t = np.arange(L)
    drift = 0.8 * np.sin(2 * np.pi * t / 900.0 + rs.rand()) + 0.002 * t * rs.randn()
    beats = np.exp(-0.5 * ((t % 200 - 60) / 4.0) ** 2) * (1.0 + 0.1 * rs.randn())
    return (drift + beats + 0.05 * rs.randn(L)).astype(np.float32)


def main():
    import dataset  # the reference module

    D = dataset.ECGMultimodalDataset
    obj = D.__new__(D)
    rs = np.random.RandomState(42)
    out = {}
    for name, L in (("l2476", 2476), ("l5000", 5000), ("l200", 200), ("l333", 333)):
        x = np.stack([synth(rs, L) for _ in range(3)])
        ref = np.stack([D.preprocess_signal(obj, r.astype(np.float64)) for r in x])
        lib_form = np.stack([op.preprocess_signal(r) for r in x])
        assert np.array_equal(ref, lib_form), name
        if L <= 2476:
            rest = np.stack([op.restated_preprocess_signal(r) for r in x])
            assert np.abs(rest - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max()), name
        out[f"{name}_x"] = x
        out[f"{name}_y"] = ref
    # the individual steps, and the evaluation_signal.py filter setting (cutoff 40 Hz at fs 250, :25-29)
    x = np.stack([synth(rs, 1000) for _ in range(2)])
    out["steps_x"] = x
    out["steps_baseline"] = np.stack([D.remove_baseline_drift(obj, r.astype(np.float64)) for r in x])
    out["steps_lowpass"] = np.stack([D.lowpass_filter(obj, r.astype(np.float64)) for r in x])
    out["steps_lowpass_40_250"] = np.stack([D.lowpass_filter(obj, r.astype(np.float64), cutoff=40, fs=250, order=5) for r in x])
    out["steps_zscore"] = np.stack([D.z_score_normalize(obj, r.astype(np.float64)) for r in x])
    for k in ("baseline", "lowpass", "zscore"):
        fn = {"baseline": op.remove_baseline_drift, "lowpass": op.lowpass_filter, "zscore": op.z_score_normalize}[k]
        assert np.array_equal(out[f"steps_{k}"], np.stack([fn(r.astype(np.float64)) for r in x])), k
    path = os.path.join(ROOT, "tests", "golden", "preprocess.npz")
    np.savez_compressed(path, **out)
    print("oracle/preprocess.py bit-identical to the reference on", len(out) // 2, "cases; wrote", path,
          os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
