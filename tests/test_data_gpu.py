"""GPU: ecgmm.data.Prefetcher -- the batches it yields are the loader's batches, on the
device, for pinned / pageable / ragged / uint8-convertible inputs, and a loop over it trains like the plain loop."""
import pytest
import torch

import ecgmm
from ecgmm import data as edata, lib
from golden_util import make_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _loader(n, B, last=None, pin=False):
    out = []
    for i in range(n):
        b = last if (last is not None and i == n - 1) else B
        image, ecg, clin, labels = make_inputs(500 + i, b, 32, 48, 200)
        u8 = (image * 127.5 + 127.5).round().clamp(0, 255).to(torch.uint8)
        image = (u8.float() / 255.0 - 0.5) / 0.5          # what ToTensor + Normalize(0.5, 0.5) yields
        t = [image, ecg, clin, labels]
        out.append(tuple(x.pin_memory() for x in t) if pin else tuple(t))
    return out


@pytest.mark.parametrize("pin", [False, True])
@pytest.mark.parametrize("depth", [1, 2, 3])
def test_prefetcher_yields_the_loader_batches(pin, depth):
    lib.require_device()
    batches = _loader(5, 4, last=3, pin=pin)
    pf = edata.Prefetcher(batches, DEV, depth=depth)
    seen = 0
    for got, want in zip(pf, batches):
        # consume on the current stream BEFORE asking for the next batch, like a training loop does
        clones = [g.clone() for g in got]
        for g, w in zip(clones, want):
            assert g.is_cuda and g.dtype == w.dtype and torch.equal(g.cpu(), w)
        seen += 1
    assert seen == len(batches) and pf.batches == len(batches)
    assert pf.h2d_bytes == sum(t.numel() * t.element_size() for b in batches for t in b)


def test_prefetcher_uint8_images_are_exact_and_smaller():
    batches = _loader(3, 4)
    pf = edata.Prefetcher(batches, DEV, images_as_uint8=True)
    for got, want in zip(pf, batches):
        assert got[0].dtype == torch.uint8
        back = (got[0].float().cpu() / 255.0 - 0.5) / 0.5
        assert torch.equal(back, want[0])
    fp32_bytes = sum(t.numel() * t.element_size() for b in batches for t in b)
    assert pf.h2d_bytes < 0.3 * fp32_bytes
    with pytest.raises(lib.EcgmmError):  # not 8-bit pixels in disguise
        bad = [(torch.rand(2, 3, 8, 8) * 2 - 1, torch.zeros(2, 10))]
        next(iter(edata.Prefetcher(bad, DEV, images_as_uint8=True)))


def test_training_loop_over_prefetcher_matches_plain_loop():
    """train.py:60-86 with the loader wrapped: same losses as with blocking .to(device) copies (uint8 images are
    bit-identical to the normalised tensors)."""
    from ecgmm import nn as enn, optim as eoptim
    from parity_util import build_pair

    losses = []
    for wrapped in (False, True):
        torch.manual_seed(0)
        _, dut = build_pair(seed=7, dropout=0.0)
        dut.train()
        opt = eoptim.Adam(dut.parameters(), lr=1e-3)
        crit = enn.CrossEntropyLoss()
        batches = _loader(4, 4)
        it = edata.Prefetcher(batches, DEV, images_as_uint8=True) if wrapped else \
            ([t.to(DEV) for t in b] for b in batches)
        cur = []
        for image, ecg, clin, labels in it:
            opt.zero_grad()
            out = dut(image, ecg, clin)
            loss = crit(out[3], labels) + 0.1 * out[4]
            loss.backward()
            opt.step()
            cur.append(float(loss.item()))
        losses.append(cur)
    assert losses[0] == losses[1], losses
