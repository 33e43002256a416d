"""CPU restatement of the reference's per-sample ECG signal preprocessing.  TEST INFRASTRUCTURE ONLY
(imported by tests/ and the generator below it; the product path is ecgmm.preprocess -> libecgmm).

Follows /root/reference/dataset.py:76-92 (identical copies: signal_model.py:203-224,
evaluation_signal.py:20-39, dataset_kfold.py):

    remove_baseline_drift   x - np.convolve(x, ones(200)/200, mode='same')             dataset.py:81-83
    lowpass_filter          butter(order=5, cutoff/(0.5*fs)) -> filtfilt(b, a, x)       dataset.py:85-89
    z_score_normalize       (x - mean) / (std_population + 1e-8)                       dataset.py:76-79
    preprocess_signal       baseline removal -> low-pass (z-score commented out)        dataset.py:91-95

The arithmetic lives in numpy / scipy.signal (third party, unpinned by the reference's README; here
scipy 1.18.1, numpy 2.3): this file calls the SAME library functions, and `restated_*` below spell the
published algorithms out in plain numpy so that the CUDA kernel has a line-by-line statement to follow.
Pinned by oracle/gen_golden_preprocess.py: both forms are compared with the reference's own methods
(imported from /root/reference/dataset.py) and the fixtures land in tests/golden/preprocess.npz.
All arithmetic is float64, as in the reference (pandas/numpy float64 rows); the result is cast to float32
where the reference builds its tensor (dataset.py:68).
"""
import numpy as np


# ----------------------------------------------------------------------------- library form
def z_score_normalize(signal):
    mean = np.mean(signal)
    std = np.std(signal)
    return (signal - mean) / (std + 1e-8)


def remove_baseline_drift(signal, window_size=200):
    baseline = np.convolve(signal, np.ones(window_size) / window_size, mode="same")
    return signal - baseline


def lowpass_filter(signal, cutoff=0.05, fs=1.0, order=5):
    from scipy.signal import butter, filtfilt

    nyq = 0.5 * fs
    b, a = butter(order, cutoff / nyq, btype="low", analog=False)
    return filtfilt(b, a, signal)


def preprocess_signal(raw_signal, cutoff=0.05, fs=1.0, order=5, window_size=200, zscore=False):
    signal = remove_baseline_drift(np.asarray(raw_signal, dtype=np.float64), window_size)
    signal = lowpass_filter(signal, cutoff, fs, order)
    if zscore:
        signal = z_score_normalize(signal)
    return signal.copy()


def preprocess_batch(x, **kw):
    """[..., L] float array -> float32 array of the same shape, every row through preprocess_signal."""
    x = np.asarray(x)
    flat = x.reshape(-1, x.shape[-1])
    out = np.stack([preprocess_signal(r, **kw) for r in flat]).astype(np.float32)
    return out.reshape(x.shape)


# ----------------------------------------------------------------------------- restated algorithms
def restated_butter_lowpass(order, wn):
    """scipy.signal.butter(order, wn, 'low') spelled out: analog prototype poles on the unit circle,
    frequency pre-warp (fs = 2), low-pass scaling, bilinear transform, polynomial expansion."""
    m = np.arange(-order + 1, order, 2)
    p = -np.exp(1j * np.pi * m / (2 * order))           # buttap
    warped = 4.0 * np.tan(np.pi * wn / 2.0)              # 2*fs*tan(pi*wn/fs), fs = 2
    p = warped * p                                        # lp2lp_zpk
    k = warped ** order
    fs2 = 4.0
    pz = (fs2 + p) / (fs2 - p)                            # bilinear_zpk
    kz = k * np.real(1.0 / np.prod(fs2 - p))
    b = kz * np.poly(-np.ones(order)).real               # zeros at z = -1
    a = np.poly(pz).real
    return b, a


def restated_lfilter_zi(b, a):
    """scipy.signal.lfilter_zi: steady-state DF2T state for a unit step (a[0] == 1)."""
    n = len(a)
    col0 = np.ones(n - 1)
    col0[0] = 1.0 + a[1]
    col0[1:] = a[2:]
    B = b[1:] - a[1:] * b[0]
    zi = np.zeros(n - 1)
    zi[0] = B.sum() / col0.sum()
    asum, csum = 1.0, 0.0
    for k in range(1, n - 1):
        asum += a[k]
        csum += b[k] - a[k] * b[0]
        zi[k] = asum * zi[0] - csum
    return zi


def _lfilter_df2t(b, a, x, z):
    n = len(a) - 1
    y = np.empty_like(x)
    z = z.copy()
    for i in range(len(x)):
        xi = x[i]
        yi = b[0] * xi + z[0]
        for k in range(n - 1):
            z[k] = b[k + 1] * xi + z[k + 1] - a[k + 1] * yi
        z[n - 1] = b[n] * xi - a[n] * yi
        y[i] = yi
    return y


def restated_preprocess_signal(raw, cutoff=0.05, fs=1.0, order=5, window_size=200, zscore=False):
    """Pure-numpy loops (slow: small cases only)."""
    x = np.asarray(raw, dtype=np.float64)
    L = len(x)
    lo = window_size // 2                                 # 'same' window of an even-length box: [i-lo, i+lo-1]
    hi = window_size - lo - 1
    cs = np.concatenate([[0.0], np.cumsum(x)])
    idx = np.arange(L)
    s = cs[np.minimum(L, idx + hi + 1)] - cs[np.maximum(0, idx - lo)]
    y = x - s / window_size
    b, a = restated_butter_lowpass(order, cutoff / (0.5 * fs))
    zi = restated_lfilter_zi(b, a)
    edge = 3 * max(len(a), len(b))                        # filtfilt default padlen, odd extension
    ext = np.concatenate([2 * y[0] - y[edge:0:-1], y, 2 * y[-1] - y[-2:-edge - 2:-1]])
    f = _lfilter_df2t(b, a, ext, zi * ext[0])
    r = _lfilter_df2t(b, a, f[::-1], zi * f[-1])[::-1]
    out = r[edge:-edge]
    if zscore:
        out = (out - out.mean()) / (out.std() + 1e-8)
    return out
