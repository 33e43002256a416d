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


# ---------------------------------------------------------------------------------------------------------------------
# Schedule emulation of the time-parallel kernel (csrc/signal_prep.cu, signal_preprocess_block_kernel):
# 256 blocks of T (odd) samples, zero-state runs, Hillis-Steele scan of s -> M s + f inside 32-block warps with the
# powers M^1, M^2, M^4, M^8, M^16, a serial pass over the 8 warps with M^32, one matrix-vector product per block, rerun.
def _df2t(b, a, x, z):
    z = z.copy()
    y = np.empty(len(x))
    order = len(a) - 1
    for t, xv in enumerate(x):
        yv = b[0] * xv + z[0]
        for k in range(order - 1):
            z[k] = b[k + 1] * xv + z[k + 1] - a[k + 1] * yv
        z[order - 1] = b[order] * xv - a[order] * yv
        y[t] = yv
    return y, z


def _block_pass(seq, b, a, z0, NT=256, WARP=32):
    order, n = len(a) - 1, len(seq)
    T = -(-n // NT) | 1
    nb = -(-n // T)
    A = np.zeros((order, order))
    for k in range(order):
        A[k, 0] = -a[k + 1]
        if k + 1 < order:
            A[k, k + 1] = 1.0
    M = np.linalg.matrix_power(A, T)
    P = [np.eye(order)]
    for _ in range(WARP):
        P.append(M @ P[-1])
    f = np.zeros((NT, order))
    for j in range(nb):
        _, f[j] = _df2t(b, a, seq[j * T:(j + 1) * T], np.zeros(order))
    nw = NT // WARP
    inc = f.reshape(nw, WARP, order).copy()
    d = 1
    while d < WARP:
        new = inc.copy()
        for lane in range(d, WARP):
            new[:, lane] = inc[:, lane] + inc[:, lane - d] @ P[d].T
        inc, d = new, d * 2
    lp = np.zeros_like(inc)
    lp[:, 1:] = inc[:, :-1]
    ws = np.zeros((nw + 1, order))
    ws[0] = z0
    for w in range(nw):
        ws[w + 1] = P[WARP] @ ws[w] + inc[w, WARP - 1]
    y = np.empty(n)
    for j in range(nb):
        w, lane = divmod(j, WARP)
        y[j * T:(j + 1) * T], _ = _df2t(b, a, seq[j * T:(j + 1) * T], P[lane] @ ws[w] + lp[w, lane])
    return y, max(np.abs(p).max() for p in P)


@pytest.mark.parametrize("L,cutoff,fs", [(5000, 0.05, 1.0), (2476, 0.05, 1.0), (600, 40.0, 250.0), (40, 0.05, 1.0),
                                         (19, 0.2, 1.0)])
def test_block_parallel_schedule_equals_filtfilt(L, cutoff, fs):
    rng = np.random.default_rng(L)
    x = np.cumsum(rng.standard_normal(L)) + 3 * np.sin(np.arange(L) / 50)
    b, a = op.restated_butter_lowpass(5, cutoff / (0.5 * fs))
    zi = op.restated_lfilter_zi(b, a)
    edge = 18
    ext = np.concatenate([2 * x[0] - x[edge:0:-1], x, 2 * x[-1] - x[-2:-edge - 2:-1]])
    y, g1 = _block_pass(ext, b, a, zi * ext[0])
    y2, _ = _block_pass(y[::-1], b, a, zi * y[-1])
    got, ref = y2[::-1][edge:edge + L], op.lowpass_filter(x, cutoff, fs, 5)
    assert g1 < 1e3  # the launcher's conditioning guard admits the reference's two designs
    assert np.abs(got - ref).max() <= 1e-8 * max(1.0, np.abs(ref).max())


def test_block_parallel_schedule_needs_the_conditioning_guard():
    """A narrow-band design: the powers of the companion matrix explode and the block result leaves the serial
    recurrence -- which is why the launcher keeps the serial kernel for it."""
    rng = np.random.default_rng(1)
    x = np.cumsum(rng.standard_normal(5000))
    b, a = op.restated_butter_lowpass(5, 0.02)
    zi = op.restated_lfilter_zi(b, a)
    ext = np.concatenate([2 * x[0] - x[18:0:-1], x, 2 * x[-1] - x[-2:-20:-1]])
    y, growth = _block_pass(ext, b, a, zi * ext[0])
    y_ser, _ = _df2t(b, a, ext, zi * ext[0])
    assert growth > 1e3 and np.abs(y - y_ser).max() > 1e-6 * np.abs(y_ser).max()
