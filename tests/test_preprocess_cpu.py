"""CPU: the preprocessing oracle (oracle/preprocess.py) against tests/golden/preprocess.npz (written by
oracle/gen_golden_preprocess.py while asserting bit-identity with /root/reference/dataset.py:76-95), and the
library's HOST-side Butterworth design (no device involved) against the restated design and scipy's."""
import os

import numpy as np
import pytest

from golden_util import GOLDEN_DIR
from oracle import preprocess as op


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "preprocess.npz"))


@pytest.mark.parametrize("case", ["l2476", "l5000", "l200", "l333"])
def test_oracle_matches_golden(golden, case):
    x, y = golden[f"{case}_x"], golden[f"{case}_y"]
    got = np.stack([op.preprocess_signal(r) for r in x])
    assert np.abs(got - y).max() <= 1e-12 * max(1.0, np.abs(y).max())


def test_restated_algorithm_matches_golden(golden):
    """Pure-numpy restatement (moving sum, Butterworth design, lfilter_zi, odd extension, DF2T both ways)."""
    for case in ("l200", "l333"):
        x, y = golden[f"{case}_x"], golden[f"{case}_y"]
        got = np.stack([op.restated_preprocess_signal(r) for r in x])
        assert np.abs(got - y).max() <= 1e-10 * max(1.0, np.abs(y).max())


def test_individual_steps_match_golden(golden):
    x = golden["steps_x"].astype(np.float64)
    for key, fn in (("baseline", op.remove_baseline_drift), ("lowpass", op.lowpass_filter),
                    ("zscore", op.z_score_normalize)):
        got = np.stack([fn(r) for r in x])
        assert np.abs(got - golden[f"steps_{key}"]).max() <= 1e-12, key
    got = np.stack([op.lowpass_filter(r, cutoff=40, fs=250, order=5) for r in x])
    assert np.abs(got - golden["steps_lowpass_40_250"]).max() <= 1e-12


@pytest.mark.parametrize("order,wn", [(5, 0.1), (5, 0.32), (2, 0.5), (1, 0.9), (4, 0.02)])
def test_native_filter_design(order, wn):
    """ecgmm_butter_lowpass runs on the host: scipy.signal.butter / lfilter_zi to ~1 ulp of the taps."""
    from scipy.signal import butter, lfilter_zi

    from ecgmm import preprocess as pp

    b, a, zi = pp.butter_lowpass(order, wn)
    rb, ra = butter(order, wn)
    assert np.abs(np.array(b) - rb).max() <= 1e-13 * np.abs(rb).max()
    assert np.abs(np.array(a) - ra).max() <= 1e-13 * np.abs(ra).max()
    rz = lfilter_zi(rb, ra)
    assert np.abs(np.array(zi) - rz).max() <= 1e-9 * np.abs(rz).max()
    b2, a2 = op.restated_butter_lowpass(order, wn)
    assert np.abs(np.array(b) - b2).max() <= 1e-13 * np.abs(b2).max() and np.abs(np.array(a) - a2).max() <= 1e-13 * np.abs(a2).max()


def test_filter_design_rejects_bad_arguments():
    from ecgmm import lib
    from ecgmm import preprocess as pp

    with pytest.raises(lib.EcgmmError):
        pp.butter_lowpass(9, 0.1)
    with pytest.raises(lib.EcgmmError):
        pp.butter_lowpass(5, 1.5)


def test_preprocess_has_no_cpu_fallback():
    import torch

    from ecgmm import lib
    from ecgmm import preprocess as pp

    with pytest.raises(lib.EcgmmError):
        pp.preprocess_signal(torch.zeros(2, 300))
