"""GPU: ecgmm.preprocess (one launch of ecgmm_signal_preprocess per batch) against the committed golden vectors
of the reference's per-sample numpy/scipy preprocessing (dataset.py:76-95) and against the CPU oracle on seeded
inputs.  Tolerance: the kernel computes in float64 like the reference and returns float32, so results must agree
to float32 rounding: |a - b| <= 2e-6 * max(1, max|b|)."""
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR
from oracle import preprocess as op

pytestmark = pytest.mark.gpu
TOL = 2e-6


def close(got, want):
    want = np.asarray(want, dtype=np.float64)
    got = got.detach().cpu().numpy().astype(np.float64)
    return np.abs(got - want).max() <= TOL * max(1.0, np.abs(want).max())


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "preprocess.npz"))


@pytest.mark.parametrize("case", ["l2476", "l5000", "l200", "l333"])
def test_preprocess_matches_reference_golden(golden, case):
    from ecgmm import preprocess as pp

    x = torch.from_numpy(golden[f"{case}_x"]).cuda()
    assert close(pp.preprocess_signal(x), golden[f"{case}_y"])
    assert close(pp.preprocess_signal(x.double()), golden[f"{case}_y"])  # float64 rows, as pandas hands them over


def test_individual_steps_match_reference_golden(golden):
    from ecgmm import preprocess as pp

    x = torch.from_numpy(golden["steps_x"]).cuda()
    assert close(pp.remove_baseline_drift(x), golden["steps_baseline"])
    assert close(pp.lowpass_filter(x), golden["steps_lowpass"])
    assert close(pp.lowpass_filter(x, cutoff=40, fs=250, order=5), golden["steps_lowpass_40_250"])
    assert close(pp.z_score_normalize(x), golden["steps_zscore"])


@pytest.mark.parametrize("shape", [(256, 12, 1000), (70, 2476), (1, 19 + 200), (33, 1, 512)])
def test_batched_layouts_match_oracle(shape):
    """[B, L], [B, 12, L] (configs[1] layout) and ragged row counts (not a multiple of the 32 signals per CTA)."""
    from ecgmm import preprocess as pp

    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(*shape, generator=g).cumsum(-1) * 0.05 + torch.randn(*shape, generator=g)).float()
    y = pp.preprocess_signal(x.cuda(), zscore=True)
    assert y.shape == x.shape and y.dtype == torch.float32
    flat = x.reshape(-1, shape[-1]).numpy()
    pick = sorted({0, flat.shape[0] // 2, flat.shape[0] - 1, min(31, flat.shape[0] - 1), min(32, flat.shape[0] - 1)})
    want = np.stack([op.preprocess_signal(flat[i], zscore=True) for i in pick])
    assert close(y.reshape(-1, shape[-1])[pick], want)


def test_linearity_and_constant_signal():
    """Size-independent properties at full configs[1] size (256 x 12 x 5000): the pipeline without z-score is
    linear, and a constant signal is removed by the baseline stage except near the zero-padded edges."""
    from ecgmm import preprocess as pp

    g = torch.Generator().manual_seed(5)
    a = torch.randn(256, 12, 5000, generator=g).cuda()
    b = torch.randn(256, 12, 5000, generator=g).cuda()
    ya, yb, yab = pp.preprocess_signal(a), pp.preprocess_signal(b), pp.preprocess_signal(2.0 * a - 3.0 * b)
    ref = 2.0 * ya.double() - 3.0 * yb.double()
    assert (yab.double() - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())
    c = pp.remove_baseline_drift(torch.full((3, 1000), 2.5, device="cuda"))
    assert c[:, 100:900].abs().max().item() <= 1e-6


def test_shape_errors():
    from ecgmm import lib
    from ecgmm import preprocess as pp

    with pytest.raises(lib.EcgmmError):
        pp.preprocess_signal(torch.zeros(2, 150, device="cuda"))  # shorter than the 200-sample window
    with pytest.raises(lib.EcgmmError):
        pp.lowpass_filter(torch.zeros(2, 18, device="cuda"))  # filtfilt needs len(x) > padlen = 18
    with pytest.raises(lib.EcgmmError):
        pp.preprocess_signal(torch.zeros(2, 300, device="cuda", dtype=torch.float16))


def test_block_parallel_kernel(golden, monkeypatch):
    """The time-parallel kernel (the default; one CTA per signal, block-wise zero-state runs + a scan of the block
    start states) against the same golden vectors and oracle, and against the serial kernel (ECGMM_PREP_BLOCK=0)."""
    from ecgmm import preprocess as pp

    monkeypatch.setenv("ECGMM_PREP_BLOCK", "0")
    serial = {c: pp.preprocess_signal(torch.from_numpy(golden[f"{c}_x"]).cuda()) for c in ("l2476", "l5000")}
    monkeypatch.setenv("ECGMM_PREP_BLOCK", "1")
    for case in ("l2476", "l5000", "l200", "l333"):
        x = torch.from_numpy(golden[f"{case}_x"]).cuda()
        assert close(pp.preprocess_signal(x), golden[f"{case}_y"]), case
        assert close(pp.preprocess_signal(x.double()), golden[f"{case}_y"]), case
    for c, y in serial.items():
        assert close(pp.preprocess_signal(torch.from_numpy(golden[f"{c}_x"]).cuda()), y.cpu().numpy())
    x = torch.from_numpy(golden["steps_x"]).cuda()
    assert close(pp.remove_baseline_drift(x), golden["steps_baseline"])
    assert close(pp.lowpass_filter(x), golden["steps_lowpass"])
    assert close(pp.lowpass_filter(x, cutoff=40, fs=250, order=5), golden["steps_lowpass_40_250"])
    assert close(pp.z_score_normalize(x), golden["steps_zscore"])
    for shape in [(256, 12, 1000), (70, 2476), (1, 19 + 200), (33, 1, 512)]:
        g = torch.Generator().manual_seed(sum(shape))
        xx = (torch.randn(*shape, generator=g).cumsum(-1) * 0.05 + torch.randn(*shape, generator=g)).float()
        y = pp.preprocess_signal(xx.cuda(), zscore=True)
        flat = xx.reshape(-1, shape[-1]).numpy()
        pick = sorted({0, flat.shape[0] // 2, flat.shape[0] - 1})
        want = np.stack([op.preprocess_signal(flat[i], zscore=True) for i in pick])
        assert close(y.reshape(-1, shape[-1])[pick], want), shape
    # a badly conditioned design (narrow band) must fall back to the serial kernel: same numbers as without the switch
    y_blk = pp.lowpass_filter(x, cutoff=0.005, fs=1.0, order=5)
    monkeypatch.setenv("ECGMM_PREP_BLOCK", "0")
    assert torch.equal(y_blk, pp.lowpass_filter(x, cutoff=0.005, fs=1.0, order=5))
