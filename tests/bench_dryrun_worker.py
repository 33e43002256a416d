"""Worker of tests/test_bench_dryrun_cpu.py::test_bench_two_ranks_dry_run (launched under torch.distributed.run with two
processes): bench.run_ours with CUDA faked, kernel launches stubbed and the process group on gloo, so that the
multi-rank control flow of the bench -- DataParallel, the ranks' agreement on the launch mode, max-over-ranks timing,
rank-0-only output and the teardown that must never hang or fail a finished run -- executes on CPUs."""
import contextlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import ecgmm  # noqa: E402,F401
from ecgmm import lib, ops  # noqa: E402
from test_bench_dryrun_cpu import FakeEvent, FakeGraph, FakeStream  # noqa: E402

calls = []


def fake_call(name, *args):
    assert len(args) == len(lib.SIGNATURES[name]), name
    calls.append(name)


lib.call = fake_call
lib.require_device = lambda: None
lib.launch_count = lambda: len(calls)
ops._s = lambda: 0
torch.Tensor.is_cuda = property(lambda self: True)
torch.Tensor.pin_memory = lambda self, *a, **k: self
torch.Tensor.record_stream = lambda self, s: None
_real_to = torch.Tensor.to


def _to(self, *a, **k):
    a = tuple(x for x in a if not (isinstance(x, torch.device) and x.type == "cuda"))
    if isinstance(k.get("device"), torch.device) and k["device"].type == "cuda":
        k.pop("device")
    return _real_to(self, *a, **k) if (a or k) else self


torch.Tensor.to = _to
for _name in ("empty_like", "tensor", "zeros", "empty"):
    _real = getattr(torch, _name)

    def _wrap(*a, _real=_real, **k):
        if isinstance(k.get("device"), torch.device) and k["device"].type == "cuda":
            k.pop("device")
        return _real(*a, **k)

    setattr(torch, _name, _wrap)
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a: None
torch.cuda.Event = FakeEvent
torch.cuda.Stream = FakeStream
torch.cuda.stream = lambda s: contextlib.nullcontext()
torch.cuda.current_stream = lambda *a: FakeStream()
torch.cuda.current_device = lambda: 0
torch.cuda.CUDAGraph = FakeGraph
torch.cuda.graph = lambda g, *a, **k: contextlib.nullcontext()
torch.cuda.empty_cache = lambda: None
torch.cuda.cudart = lambda: types.SimpleNamespace(cudaHostRegister=lambda *a: 0)
os.environ["ECGMM_SIDE_STREAM"] = "0"
_real_init = dist.init_process_group
dist.init_process_group = lambda backend=None, **k: _real_init("gloo")

bench.H, bench.W, bench.L = 64, 160, 600
args = types.SimpleNamespace(gpus=int(os.environ["WORLD_SIZE"]), steps=2, warmup=3, impl="ours", global_batch=4,
                             no_cpu_baseline=True, detail=False, launch="graph", image_dtype="uint8")
bench.run_ours(args)
print("WORKER RETURNED WITHOUT THE HARD EXIT", flush=True)  # shutdown() ends every rank of a multi-rank run itself
