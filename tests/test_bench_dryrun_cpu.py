"""CPU: bench.py's own arm executed end to end with CUDA faked and the kernel launches stubbed (values are garbage,
control flow is real): warm-up, the timed loop, the double-buffered end-to-end loop with the pipelined loss read, the
per-kernel-class table, the roofline object and the JSON line.  The bench is what the driver runs at the end of a
round; a NameError in a path that only a GPU box reaches would cost the round its numbers."""
import contextlib
import ctypes
import json
import types

import pytest
import torch

import bench
from ecgmm import lib, ops


class FakeEvent:
    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return 2.0


class FakeGraph:
    def __init__(self, *a, **k):
        self.replays = 0

    def replay(self):
        self.replays += 1

    def reset(self):
        pass


class FakeStream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        self.device = types.SimpleNamespace(index=0)

    def wait_event(self, e):
        pass

    def wait_stream(self, s):
        pass


@pytest.fixture
def fake_cuda(monkeypatch):
    calls = []

    def fake_call(name, *args):
        assert len(args) == len(lib.SIGNATURES[name]), name
        calls.append(name)

    monkeypatch.setattr(lib, "call", fake_call)
    monkeypatch.setattr(lib, "require_device", lambda: None)
    monkeypatch.setattr(lib, "launch_count", lambda: len(calls))
    monkeypatch.setattr(ops, "_s", lambda: 0)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True), raising=False)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self, raising=False)
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None, raising=False)
    real_to = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple(x for x in a if not (isinstance(x, torch.device) and x.type == "cuda"))
        if isinstance(k.get("device"), torch.device) and k["device"].type == "cuda":
            k.pop("device")
        return real_to(self, *a, **k) if (a or k) else self

    monkeypatch.setattr(torch.Tensor, "to", to)
    for name in ("empty_like", "tensor", "zeros", "empty"):
        real = getattr(torch, name)

        def wrap(*a, _real=real, **k):
            if isinstance(k.get("device"), torch.device) and k["device"].type == "cuda":
                k.pop("device")
            return _real(*a, **k)

        monkeypatch.setattr(torch, name, wrap)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: FakeStream())
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.cuda, "CUDAGraph", FakeGraph)
    monkeypatch.setattr(torch.cuda, "graph", lambda g, *a, **k: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.cuda, "cudart", lambda: types.SimpleNamespace(cudaHostRegister=lambda *a: 0))
    monkeypatch.setenv("ECGMM_SIDE_STREAM", "0")
    return calls


@pytest.mark.parametrize("sync_loss,launch", [("0", "eager"), ("1", "eager"), ("0", "graph")])
def test_bench_own_arm_dry_run(fake_cuda, monkeypatch, capsys, sync_loss, launch):
    monkeypatch.setenv("ECGMM_BENCH_SYNC_LOSS", sync_loss)
    monkeypatch.setattr(bench, "H", 64)
    monkeypatch.setattr(bench, "W", 160)
    monkeypatch.setattr(bench, "L", 600)
    args = types.SimpleNamespace(gpus=1, steps=3, warmup=3, impl="ours", global_batch=2, no_cpu_baseline=True,
                                 detail=True, launch=launch, image_dtype="uint8")
    bench.run_ours(args)
    out = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(out) == 1
    line = json.loads(out[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "kernels"):
        assert key in line, key
    assert line["config"]["workload"].startswith("configs[2]") and line["graph_note"] is None
    assert line["config"]["launch"] == ("cuda_graph" if launch == "graph" else "eager")
    assert line["e2e"]["loss_read"].startswith("blocking" if sync_loss == "1" else "pipelined")
    assert line["e2e"]["note"] is None and line["e2e"]["h2d_bytes_per_step"] > 0 and line["gpu_launches"] > 300
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert "hbm" in line["kernels"]["stem_fwd"] and "hbm" in line["kernels"]["stem_wgrad"]
    assert line["cpu_baseline"] is None  # --no-cpu-baseline in this dry run


def test_bench_two_ranks_dry_run():
    """Two processes under torch.distributed.run (gloo, CUDA faked): one JSON line from rank 0, exit code 0, no hang in
    the teardown."""
    import os
    import socket
    import subprocess
    import sys

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(root, "tests", "bench_dryrun_worker.py")],
                       capture_output=True, text=True, timeout=300, cwd=root)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and line["config"]["parallelism"] == "dp2" and line["config"]["per_gpu_batch"] == 2
    assert line["config"]["launch"] == "cuda_graph" and line["graph_note"] is None
    assert line["allreduce"]["buckets_per_step"] >= 4 and line["cpu_baseline"] is None
    assert "WORKER RETURNED" not in p.stdout


def test_endpoint_and_eval_graph_plumbing(fake_cuda):
    """ecgmm.serve.ImageEndpoint(graph=True) and ecgmm.graph.GraphedEvalStep with a fake CUDAGraph: warm-up + capture
    once per request kind, replays issue no library call, a changed weight re-captures, shape / mode errors raise."""
    import ecgmm
    from ecgmm import graph as eg
    from ecgmm import serve

    class Cfg:
        num_classes = 2
        device = "cpu"

    m = ecgmm.ECGMultimodalModel(Cfg).eval()
    m.overlap_branches = False
    u8 = (torch.rand(2, 3, 64, 160) * 255).to(torch.uint8)
    ep = serve.ImageEndpoint(m, example_image=u8, graph=True)
    probs, classes = ep(u8)
    assert probs.shape == (2, 2) and classes.shape == (2,)
    n0 = len(fake_cuda)
    ep(u8 + 1)
    assert len(fake_cuda) == n0 and ep._graphs["classify"][0].replays == 2
    p2, c2, cam = ep.gradcam(u8)
    assert cam.shape == (2, 2, 5) and set(ep._graphs) == {"classify", "cam"}
    with torch.no_grad():
        m.image_classifier.bias.add_(1.0)
    g_old = ep._graphs["classify"][0]
    ep(u8)
    assert ep._graphs["classify"][0] is not g_old  # new weights: captured again
    with pytest.raises(lib.EcgmmError):
        ep(u8[:1])
    batch = (torch.randn(2, 3, 64, 160), torch.randn(2, 600), torch.randn(2, 24))
    infer = eg.GraphedEvalStep(m, batch)
    out = infer(*batch)
    assert len(out) == 6 and out[3].shape == (2, 2)
    n0 = len(fake_cuda)
    infer(*batch)
    assert len(fake_cuda) == n0 and infer.graph.replays == 2
    m.train()
    with torch.no_grad():
        m.image_classifier.bias.add_(1.0)
    with pytest.raises(lib.EcgmmError):
        infer(*batch)


def test_serve_bench_dry_run(fake_cuda, monkeypatch, capsys):
    """tools/serve_bench.py (written without hardware) end to end with CUDA faked."""
    import importlib.util
    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("serve_bench", os.path.join(root, "tools", "serve_bench.py"))
    sb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sb)
    monkeypatch.setattr(sb, "H", 64)
    monkeypatch.setattr(sb, "W", 160)
    monkeypatch.setattr(sys, "argv", ["serve_bench.py", "--batch", "2", "--iters", "2", "--cpu-images", "1"])
    sb.main()
    line = json.loads([ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")][0])
    assert line["metric"] == "image-endpoint images/sec" and line["config"]["batch"] == 2
    assert line["eager"]["launches_per_request"] > 40 and line["e2e"]["h2d_bytes_per_step"] == 2 * 3 * 64 * 160
    assert line["cpu_baseline"]["kind"] == "port" and line["with_gradcam"]["images_per_s"] > 0
