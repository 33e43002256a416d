"""CPU, world_size 2, gloo: the bucketed gradient all-reduce logic of ecgmm.parallel.DataParallel
(bucket boundaries, averaging, completion before optimizer.step) on stand-in stages."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _FakeArena:
    def __init__(self, n, fill):
        self.flat = torch.full((n,), float(fill))
        self.total = n


class _FakeStage:
    pass


class _FakeModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(3))
        self._stages = [_FakeStage(), _FakeStage()]

    def stages(self):
        return self._stages

    def forward(self, x):
        return x


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ecgmm.parallel import DataParallel

        m = _FakeModel()
        with torch.no_grad():
            m.w.fill_(float(rank + 1))
        dp = DataParallel(m, min_bucket_elems=4)
        assert torch.equal(m.w.detach(), torch.ones(3)), "parameters must be broadcast from rank 0"
        dp(torch.zeros(1))
        a = _FakeArena(20, fill=rank + 1)     # rank0: 1, rank1: 2 -> mean 1.5
        b = _FakeArena(6, fill=10 * (rank + 1))
        cb = m.stages()[0]._grad_ready_cb
        cb(a, 2)      # below min bucket: deferred
        assert dp.buckets_last_step == 0
        cb(a, 8)      # [0, 8)
        cb(a, 8)      # nothing new
        cb(a, 20)     # [8, 20), final
        m.stages()[1]._grad_ready_cb(b, 6)
        assert dp.buckets_last_step == 3 and dp.bytes_last_step == 4 * 26
        dp.finish()
        ok = bool(torch.allclose(a.flat, torch.full((20,), 1.5)) and torch.allclose(b.flat, torch.full((6,), 15.0)))

        # the same through a real backward pass: the collectives must be complete when backward() returns
        c = _FakeArena(12, fill=4 * (rank + 1))  # mean 6

        class Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x):
                return x * 2

            @staticmethod
            def backward(ctx, g):
                cb(c, 5)
                cb(c, 12)
                return g * 2

        dp(torch.zeros(1))
        x = torch.ones(2, requires_grad=True)
        Fn.apply(x).sum().backward()
        ok = ok and bool(torch.allclose(c.flat, torch.full((12,), 6.0))) and not dp._pending
        # no_sync(): gradients stay local
        d = _FakeArena(4, fill=rank + 1)
        with dp.no_sync():
            cb(d, 4)
        ok = ok and bool(torch.allclose(d.flat, torch.full((4,), float(rank + 1))))
        q.put((rank, ok, dp.buckets_last_step))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res)


def test_param_grads_alias_the_arena():
    """The contract the overlapped all-reduce relies on: a gradient handed to autograd as a GradArena slice becomes
    p.grad WITHOUT a copy (AccumulateGrad adopts a tensor nobody else references), so values that the asynchronous
    all-reduce writes into arena.flat after backward() returned its tensors are what optimizer.step() reads.
    (A cache of the slice views inside the arena broke exactly this: every p.grad became a pre-reduction clone.)"""
    from ecgmm.model import GradArena

    w = torch.nn.Parameter(torch.zeros(5, 3))
    b = torch.nn.Parameter(torch.zeros(7))
    holder = {}

    class Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w_, b_):
            return x.clone()

        @staticmethod
        def backward(ctx, g):
            G = GradArena([w, b], torch.device("cpu"))
            G(w).fill_(1.0)  # a kernel writing the gradient; the same accessor is used again for the return value
            G(b).fill_(2.0)
            holder["arena"] = G
            return g, G(w), G(b)

    x = torch.ones(2, requires_grad=True)
    Fn.apply(x, w, b).sum().backward()
    G = holder["arena"]
    lo, hi = G.flat.data_ptr(), G.flat.data_ptr() + G.flat.numel() * 4
    assert lo <= w.grad.data_ptr() < hi and lo <= b.grad.data_ptr() < hi, "p.grad must alias the arena"
    G.flat.mul_(0.5)  # what the all-reduce does after the fact
    assert torch.equal(w.grad, torch.full((5, 3), 0.5)) and torch.equal(b.grad, torch.full((7,), 1.0))


def _accum_worker(rank, world, port, q):
    """Gradient accumulation: a micro-batch under no_sync(), then a synchronising one; and zero_grad(set_to_none=False).
    In both cases p.grad exists when backward runs, so autograd ADDS the arena into it: the wrapper must reduce the
    accumulated p.grad (torch DDP semantics), not the arena."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ecgmm.model import GradArena
        from ecgmm.parallel import DataParallel

        class Stage:
            def __init__(self):
                self.w = torch.nn.Parameter(torch.zeros(5, 3))
                self.b = torch.nn.Parameter(torch.zeros(7))

            def cached_params(self):
                return [self.w, self.b]

        class Model(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.st = Stage()
                self.w, self.b = self.st.w, self.st.b

            def stages(self):
                return [self.st]

            def forward(self, x, scale):
                st = self.st

                class Fn(torch.autograd.Function):
                    @staticmethod
                    def forward(ctx, x_, w_, b_):
                        return x_.clone()

                    @staticmethod
                    def backward(ctx, g):
                        G = GradArena([st.w, st.b], torch.device("cpu"))
                        G(st.w).fill_(scale * (rank + 1))       # rank0: s, rank1: 2s -> mean 1.5 s
                        G(st.b).fill_(10 * scale * (rank + 1))  # mean 15 s
                        st._grad_ready_cb(G, G.end_of(st.b))    # a prefix bucket ...
                        st._grad_ready_cb(G, G.total)           # ... and the rest
                        return g, G(st.w), G(st.b)

                return Fn.apply(x, st.w, st.b)

        m = Model()
        dp = DataParallel(m)
        x = torch.ones(2, requires_grad=True)
        # (1) plain step: arena adopted, reduced in place, two buckets
        dp(x, 1.0).sum().backward()
        ok = bool(torch.allclose(m.w.grad, torch.full((5, 3), 1.5)) and torch.allclose(m.b.grad, torch.full((7,), 15.0)))
        ok = ok and dp.buckets_last_step == 2
        # (2) accumulate: micro-batch 1 local (scale 1), micro-batch 2 synchronising (scale 3) -> mean of (1+3)*(rank+1)
        m.w.grad = m.b.grad = None
        with dp.no_sync():
            dp(x, 1.0).sum().backward()
        ok = ok and bool(torch.allclose(m.w.grad, torch.full((5, 3), float(rank + 1))))
        dp(x, 3.0).sum().backward()
        ok = ok and bool(torch.allclose(m.w.grad, torch.full((5, 3), 6.0)) and torch.allclose(m.b.grad, torch.full((7,), 60.0)))
        ok = ok and dp.buckets_last_step == 1 and not dp._pending and not dp._deferred
        # (3) zero_grad(set_to_none=False): p.grad is a zero tensor that autograd accumulates into
        m.w.grad.zero_()
        m.b.grad.zero_()
        dp(x, 2.0).sum().backward()
        ok = ok and bool(torch.allclose(m.w.grad, torch.full((5, 3), 3.0)) and torch.allclose(m.b.grad, torch.full((7,), 30.0)))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_accumulated_grads_are_reduced_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_accum_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res


def test_folds_are_partitioned_over_ranks():
    from ecgmm.parallel import folds_for_rank

    for n_folds, world in ((5, 8), (5, 2), (15, 4), (3, 1)):
        seen = sorted(k for r in range(world) for k in folds_for_rank(n_folds, r, world))
        assert seen == list(range(n_folds))
        sizes = [len(folds_for_rank(n_folds, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    import pytest

    with pytest.raises(ValueError):
        folds_for_rank(5, 3, 2)


def _full_model_worker(rank, world, port, q):
    """The REAL model under DataParallel over gloo, kernel launches stubbed (see test_control_flow_cpu.py)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ecgmm
        from ecgmm import lib, ops
        from ecgmm import nn as enn
        from ecgmm import optim as eoptim
        from ecgmm.parallel import DataParallel

        from ecgmm import model as M

        lib.call = lambda name, *a: None
        ops._s = lambda: 0
        ops._chk = lambda *a: None            # backward runs with honest CPU tensors (gloo must see them as such)
        M._require_cuda_f32 = lambda *a: None
        torch.Tensor.pin_memory = lambda self, *a, **k: self

        class Cfg:
            num_classes = 2
            device = "cpu"

        torch.manual_seed(100 + rank)  # different weights per rank: the wrapper must broadcast rank 0's
        m = ecgmm.ECGMultimodalModel(Cfg)
        m.overlap_branches = False
        m.train()
        dp = DataParallel(m)
        w0 = m.fusion_classifier.lin1.weight.detach().clone()
        gathered = [torch.empty_like(w0) for _ in range(world)]
        dist.all_gather(gathered, w0)
        same_init = all(torch.equal(gathered[0], t) for t in gathered)
        torch.Tensor.is_cuda = property(lambda self: True)  # only now: gloo itself must see CPU tensors above
        g = torch.Generator().manual_seed(rank)
        image, ecg, clin = torch.randn(2, 3, 64, 160, generator=g), torch.randn(2, 600, generator=g), torch.randn(2, 24, generator=g)
        opt = eoptim.Adam(m.parameters(), lr=1e-3)
        opt.zero_grad()
        out = dp(image, ecg, clin)
        loss = enn.CrossEntropyLoss()(out[3], torch.tensor([0, 1])) + 0.1 * out[4]
        del torch.Tensor.is_cuda  # the collectives launched from backward run on gloo
        loss.backward()
        ok = same_init and not dp._pending
        q.put((rank, ok, dp.buckets_last_step, dp.bytes_last_step))
    finally:
        dist.destroy_process_group()


def test_full_model_buckets_gloo_world2():
    """Bucket schedule of the real model: 7 all-reduces per step (image encoder 4, signal, clinical, head) covering
    every trainable parameter once -- the numbers the 8-GPU bench line reports (7 buckets, 47 649 984 bytes)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_full_model_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert all(r[2] == 7 and r[3] == 47649984 for r in res), res
