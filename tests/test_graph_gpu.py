"""GPU: ecgmm.graph.GraphedTrainStep (one CUDA graph per training step) against the eager loop of train.py:60-86 on
the same weights and batches.  Every kernel is the same; only the stride-2 weight gradients use fp32 atomics, so the
first steps agree to ~1e-6 and the difference then grows with the (ill-conditioned, see parity_util) training
dynamics: losses are compared at 2e-3, weights at 2e-3 relative L2 after 4 steps."""
import pytest
import torch

import ecgmm
from ecgmm import graph as egraph
from ecgmm import lib
from ecgmm import nn as enn
from ecgmm import optim as eoptim
from golden_util import make_inputs, set_dropout
from parity_util import build_pair

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _pair(dropout):
    _, a = build_pair(seed=7, dropout=dropout)
    _, b = build_pair(seed=7, dropout=dropout)
    a.train()
    b.train()
    return a, b


def _batches(n, B=4, H=64, W=160, L=600):
    return [[t.to(DEV) for t in make_inputs(100 + i, B, H, W, L)] for i in range(n)]


def _eager(model, opt, crit, batch):
    opt.zero_grad()
    out = model(*batch[:3])
    loss = crit(out[3], batch[3]) + 0.1 * out[4]
    loss.backward()
    opt.step()
    return float(loss)


def test_graphed_steps_match_eager_steps():
    a, b = _pair(0.0)
    crit = enn.CrossEntropyLoss()
    oa, ob = eoptim.Adam(a.parameters(), lr=2e-4), eoptim.Adam(b.parameters(), lr=2e-4)
    batches = _batches(4)
    sd0 = {k: v.detach().clone() for k, v in b.state_dict().items()}
    step = egraph.GraphedTrainStep(b, crit, ob, batches[0])
    # construction warms up on the example batch and must put everything back
    for k, v in b.state_dict().items():
        assert torch.equal(v, sd0[k]), k
    n0 = lib.launch_count()
    losses_a, losses_b = [], []
    for i, batch in enumerate(batches):
        if i == 2:  # train.py:158-161 style learning-rate change between steps
            for o in (oa, ob):
                for g in o.param_groups:
                    g["lr"] /= 10
        losses_a.append(_eager(a, oa, crit, batch))
        losses_b.append(float(step(*batch)))
    assert lib.launch_count() - n0 > 300 * len(batches) - 400  # the eager model's launches; the graph adds none
    assert abs(losses_a[0] - losses_b[0]) <= 1e-5 * max(1.0, abs(losses_a[0])), (losses_a, losses_b)
    for x, y in zip(losses_a, losses_b):
        assert abs(x - y) <= 2e-3 * max(1.0, abs(x)), (losses_a, losses_b)
    for (k, p), q in zip(a.named_parameters(), b.parameters()):
        assert float((p - q).norm()) <= 2e-3 * float(p.norm()) + 1e-6, k
    for (k, u), v in zip(a.named_buffers(), b.buffers()):
        assert float((u.double() - v.double()).norm()) <= 2e-3 * float(u.double().norm()) + 1e-6, k
    # host-side optimizer bookkeeping follows the replays (checkpoints, state_dict)
    assert all(int(ob.state[p]["step"]) == len(batches) for p in b.parameters())
    assert int(b.image_encoder.bn1.num_batches_tracked) == len(batches)
    # an eager evaluation pass afterwards sees the trained weights (weight shadows are refreshed)
    a.eval()
    b.eval()
    with torch.no_grad():
        ya, yb = a(*batches[0][:3])[3], b(*batches[0][:3])[3]
    assert float((ya - yb).abs().max()) <= 2e-2 * max(1.0, float(ya.abs().max()))


def test_second_graphed_step_after_the_first_is_gone():
    """Regression: a replaced Adam chunk table (pinned host memory) used to be freed INSIDE the capture of the next
    graphed step; torch's pinned allocator then recorded its free-time event on a capturing stream whenever torch's
    stream pool had handed the earlier warm-up stream out again as the capture stream, and the next Tensor.item()
    died with cudaErrorInvalidValue (cuEventQuery of a captured event).  Build, drop and rebuild several times."""
    import gc

    for i in range(3):
        _, b = _pair(0.0)
        crit = enn.CrossEntropyLoss()
        ob = eoptim.Adam(b.parameters(), lr=2e-4)
        batch = _batches(1)[0]
        step = egraph.GraphedTrainStep(b, crit, ob, batch)
        losses = [float(step(*batch)) for _ in range(2)]
        assert losses[1] < losses[0]
        del step, ob, b
        gc.collect()
        for _ in range(40):  # walk torch's stream pool so that different pool streams meet the capture
            torch.cuda.Stream()


def test_dropout_masks_change_between_replays():
    _, b = _pair(0.3)
    crit = enn.CrossEntropyLoss()
    ob = eoptim.Adam(b.parameters(), lr=0.0)  # lr 0: the weights stay put, only the dropout masks can differ
    batch = _batches(1)[0]
    step = egraph.GraphedTrainStep(b, crit, ob, batch)
    losses = {round(float(step(*batch)), 7) for _ in range(4)}
    assert len(losses) > 1, losses


def test_shape_mismatch_is_rejected():
    _, b = _pair(0.0)
    ob = eoptim.Adam(b.parameters(), lr=1e-3)
    batch = _batches(1)[0]
    step = egraph.GraphedTrainStep(b, enn.CrossEntropyLoss(), ob, batch)
    bad = _batches(1, B=2)[0]
    with pytest.raises(lib.EcgmmError):
        step(*bad)
    with pytest.raises(lib.EcgmmError):
        egraph.GraphedTrainStep(b, enn.CrossEntropyLoss(), torch.optim.Adam(b.parameters()), batch)
